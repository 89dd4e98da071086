// FP64 pipe of the B200 SM: dependent-issue latency and throughput of DADD / DSETP+FSEL, measured
// with clock64 (development aid; explains what paces the float64 DP and Gibbs kernels).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dadd_chain(double *out, long long *cyc, int iters) {
    double a[ILP];
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-9 + i;
    const double b = 1.0000001;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = __dadd_rn(a[i], b);
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
__global__ void dmax_chain(double *out, long long *cyc, int iters) {
    double a[ILP], m[ILP];
    for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x * 1e-9 + i; m[i] = -1e300; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) { const double c = __dadd_rn(a[i], m[i] * 1e-300); m[i] = (m[i] > c) ? m[i] : c; a[i] = c; }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    const int iters = 4096;
#define RUN(K, ILP, W) { K<ILP><<<1, 32 * W>>>(out, cyc, iters); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf(#K " ILP=%d warps/SM=%d: %.2f clk per op per warp (%.2f clk per dependent step)\n", ILP, W, (double)h / iters / ILP, (double)h / iters); }
    RUN(dadd_chain, 1, 1) RUN(dadd_chain, 2, 1) RUN(dadd_chain, 4, 1) RUN(dadd_chain, 8, 1) RUN(dadd_chain, 16, 1)
    RUN(dadd_chain, 8, 4) RUN(dadd_chain, 8, 8) RUN(dadd_chain, 8, 16) RUN(dadd_chain, 8, 32)
    RUN(dmax_chain, 1, 1) RUN(dmax_chain, 6, 1) RUN(dmax_chain, 6, 8)
    return 0;
}
