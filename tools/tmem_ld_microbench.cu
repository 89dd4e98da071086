// TMEM read-out rate of one B200 SM (development aid): how fast can the epilogue warps of the
// filter GEMM drain accumulators with tcgen05.ld.32x32b?  Decides whether kmeans_filter_kernel sits
// on a hardware ceiling (TMEM port) or on its own instruction stream.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_ld_microbench tools/tmem_ld_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD_REGS16(r, o) "=r"(r[o+0]), "=r"(r[o+1]), "=r"(r[o+2]), "=r"(r[o+3]), "=r"(r[o+4]), "=r"(r[o+5]), "=r"(r[o+6]), "=r"(r[o+7]), "=r"(r[o+8]), "=r"(r[o+9]), "=r"(r[o+10]), "=r"(r[o+11]), "=r"(r[o+12]), "=r"(r[o+13]), "=r"(r[o+14]), "=r"(r[o+15])

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : LD_REGS16(r, 0), LD_REGS16(r, 16) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld64(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : LD_REGS16(r, 0), LD_REGS16(r, 16), LD_REGS16(r, 32), LD_REGS16(r, 48) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: x64 load, wait, xor-consume.  MODE 1: two x32 loads in flight (software pipelined), xor-consume.
// MODE 2: x64 load, wait, 64-wide fmax tree (the filter epilogue's arithmetic floor).
template <int MODE>
__global__ void __launch_bounds__(512, 1) tmem_read(int iters, long long *cyc, uint32_t *sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    const int part = warp >> 2;                    // warps sharing a lane quadrant read different columns
    uint32_t acc = 0;
    float fm = -1e30f;
    __syncthreads();
    const long long t0 = clock64();
    if (MODE == 0 || MODE == 2) {
        for (int it = 0; it < iters; ++it) {
            uint32_t r[64];
            ld64(base + ((part * 64 + it * 64) & 511 & ~63), r);
            ld_wait();
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < 64; j += 2) acc ^= r[j] ^ r[j + 1];
            } else {
                float m = __uint_as_float(r[0]);
#pragma unroll
                for (int j = 1; j < 64; ++j) m = fmaxf(m, __uint_as_float(r[j]));
                fm = fmaxf(fm, m);
            }
        }
    } else {
        uint32_t a[32], b[32];
        ld32(base + ((part * 64) & 511), a);
        for (int it = 0; it < iters; ++it) {
            ld32(base + ((part * 64 + it * 64 + 32) & 511), b);
            // wait::ld waits for ALL outstanding loads, so pipelining is limited to issue order
            ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 2) acc ^= a[j] ^ a[j + 1];
            ld32(base + ((part * 64 + it * 64 + 64) & 511), a);
#pragma unroll
            for (int j = 0; j < 32; j += 2) acc ^= b[j] ^ b[j + 1];
            ld_wait();
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(fm);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
    }
}

template <int MODE>
static void run(const char *name, int warps, int iters, long long *d_cyc, uint32_t *d_sink) {
    tmem_read<MODE><<<148, warps * 32>>>(iters, d_cyc, d_sink);
    cudaDeviceSynchronize();
    tmem_read<MODE><<<148, warps * 32>>>(iters, d_cyc, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    const double bytes = (double)warps * iters * 64 * 32 * 4;   // fp32 cells delivered per SM
    printf("%-28s warps=%2d  %8lld clk  %7.1f B/clk/SM  (%.0f clk per 256x128 fp32 tile)  %s\n", name, warps, cyc,
           bytes / cyc, 131072.0 / (bytes / cyc), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long *d_cyc;
    uint32_t *d_sink;
    cudaMalloc(&d_cyc, 8);
    cudaMalloc(&d_sink, 148 * 512 * 4);
    const int iters = 4096;
    for (int warps : {4, 8, 16}) {
        run<0>("x64 + wait + xor", warps, iters, d_cyc, d_sink);
        run<1>("2 x x32 pipelined + xor", warps, iters, d_cyc, d_sink);
        run<2>("x64 + wait + fmax tree", warps, iters, d_cyc, d_sink);
    }
    return 0;
}
