// tcgen05.mma issue-rate microbenchmark (development aid): clocks per kind::f16 MMA, cta_group::1,
// A and B from shared memory (K-major, no swizzle, the layout of the libsegb200 tile images), for
// several shapes and accumulator patterns.  Decides the tile shape of the scoring GEMMs:
//   - does a chain of MMAs on ONE accumulator run at the 64-clock floor (M=128, N=128)?
//   - is the floor set by the tensor pipe or by the shared-memory operand reads (A re-read per MMA)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_microbench tools/mma_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, int R) {
    return ((saddr >> 4) & 0x3FFFu) | ((uint32_t)((R / 8) * 128 >> 4) << 16);
}
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
        "setp.ne.b32 p, %5, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(DESC_HI), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// PATTERN 0: all MMAs accumulate into one tile.  1: alternate between two accumulator tiles.
// 2: chains of 9 on one tile, then 9 on the other (the filter kernel's order).
template <int N, int PATTERN>
__global__ void __launch_bounds__(128, 1) mma_rate(int iters, long long *cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (128 + N) * 144 * 2 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp == 1 && elect_one()) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = desc_lo(smem_u32(smem), 128), b0 = desc_lo(smem_u32(smem + 128 * 144 * 2), N);
        constexpr uint32_t KA = (2 * (128 / 8) * 128) >> 4, KB = (2 * (N / 8) * 128) >> 4;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                if (PATTERN == 0) { mma(tmem, a0 + k * KA, b0 + k * KB, idesc, 1); mma(tmem, a0 + k * KA, b0 + k * KB, idesc, 1); }
                if (PATTERN == 1) { mma(tmem, a0 + k * KA, b0 + k * KB, idesc, 1); mma(tmem + (N == 256 ? 256 : 128), a0 + k * KA, b0 + k * KB, idesc, 1); }
            }
            if (PATTERN == 2) {
#pragma unroll
                for (int k = 0; k < 9; ++k) mma(tmem, a0 + k * KA, b0 + k * KB, idesc, 1);
#pragma unroll
                for (int k = 0; k < 9; ++k) mma(tmem + (N == 256 ? 256 : 128), a0 + k * KA, b0 + k * KB, idesc, 1);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(smem_u32(&bar), 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) *cyc = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int N, int PATTERN>
static void run(const char *name, long long *d_cyc) {
    const int iters = 2000;
    const size_t smem = (128 + N) * 144 * 2 + 1024;
    cudaFuncSetAttribute(mma_rate<N, PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) { mma_rate<N, PATTERN><<<148, 128, smem>>>(iters, d_cyc); cudaDeviceSynchronize(); }
    cudaError_t e = cudaGetLastError();
    long long cyc = 0;
    cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)cyc / (iters * 18.0);
    printf("M=128 N=%3d %-34s %7.1f clk/MMA  = %5.1f %% of the %d-clk floor  %s\n", N, name, per, 100.0 * (N / 2) / per, N / 2,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long *d_cyc;
    cudaMalloc(&d_cyc, 8);
    run<128, 0>("one accumulator", d_cyc);
    run<128, 1>("alternating two accumulators", d_cyc);
    run<128, 2>("9 + 9 (filter kernel order)", d_cyc);
    run<256, 0>("one accumulator", d_cyc);
    run<256, 1>("alternating two accumulators", d_cyc);
    run<256, 2>("9 + 9", d_cyc);
    run<64, 0>("one accumulator", d_cyc);
    return 0;
}
