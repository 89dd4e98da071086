// tcgen05 / TMA / mbarrier PTX wrappers and the operand "tile image" layout shared by the
// tensor-core kernels of libsegb200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace segb {
namespace mma {

// Operands live in HBM as fp16 "tile images": consecutive R-row tiles, each already in the UMMA
// canonical K-major no-swizzle shared-memory layout (8x8 core matrices of 128 contiguous bytes).
// A tile is ONE contiguous cp.async.bulk into shared memory.
// byte offset of element (r, c) inside an R-row tile image
__host__ __device__ inline int tile_off_r(int r, int c, int R) {
    return ((c >> 3) * (R / 8) + (r >> 3)) * 128 + (r & 7) * 16 + (c & 7) * 2;
}
// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// SBO = distance between 8-row core matrices (128 B), LBO = distance between the two
// 8-element K chunks of one K=16 instruction ((R/8)*128 B for an R-row tile).
__device__ __forceinline__ uint64_t make_desc_r(uint32_t saddr, int R) {
    const uint64_t lbo = (uint64_t)(R / 8) * 128, sbo = 128;
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: A,B = fp16 K-major, D = fp32, M x N
__device__ __forceinline__ uint32_t make_idesc_mn(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- PTX wrappers

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps (context error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    printf("segb mma: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
    __trap();
}
// Wait for TWO barriers with one poll loop: both try_wait probes are in flight together, so the
// single issuing thread pays one shared-memory round trip per tile instead of two.
__device__ __forceinline__ void mbar_wait2(uint32_t bar_a, uint32_t par_a, uint32_t bar_b, uint32_t par_b) {
    uint32_t da = 0, db = 0;
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%2], %3;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 q, [%4], %5;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "selp.u32 %1, 1, 0, q;\n\t}"
            : "=r"(da), "=r"(db) : "r"(bar_a), "r"(par_a), "r"(bar_b), "r"(par_b) : "memory");
        if (da && db) return;
    }
    printf("segb mma: mbarrier pair wait timed out (block %d thread %d bars %u %u)\n", blockIdx.x, threadIdx.x, bar_a, bar_b);
    __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum) : "memory");
}
// One lane of a converged warp.  The compiler knows elect.sync yields exactly one thread, so the
// single-thread tcgen05 / bulk-copy instructions issued under it need no per-active-lane loop
// (with `lane == 0` every tcgen05.mma was wrapped in an ELECT / BRA.U.ANY waterfall).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// Shared-memory descriptors split into words: the high word (SBO = 128 B, descriptor version 1) is the
// same for every K-major no-swizzle operand; the low word is (addr >> 4) | (LBO >> 4) << 16, and a K
// step only adds a constant to it (shared memory is < 256 KB, so the 14-bit address field cannot carry).
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t make_desc_lo(uint32_t saddr, int R) {
    return ((saddr >> 4) & 0x3FFFu) | ((uint32_t)((R / 8) * 128 >> 4) << 16);
}
// kind::f8f6f4 with e4m3 operands: the same 32 bytes of K per row and instruction (K = 32 elements), the same
// K-major no-swizzle descriptors and -- e4m3 and f16 both being format code 0 -- the same instruction descriptor
template <bool ACCUM>
__device__ __forceinline__ void tc_mma_f8_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %4};\n\t"
        "mov.b64 db, {%2, %4};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(DESC_HI), "n"(ACCUM ? 1 : 0) : "memory");
}
template <bool ACCUM>
__device__ __forceinline__ void tc_mma_f16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %4};\n\t"
        "mov.b64 db, {%2, %4};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(DESC_HI), "n"(ACCUM ? 1 : 0) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 64 consecutive fp32 columns -> 64 registers per thread (thread = TMEM lane = row).
// Load and wait live in ONE asm statement so no use of the outputs can be scheduled before
// tcgen05.wait::ld.
__device__ __forceinline__ void tc_ld64_wait(uint32_t taddr, float *v) {
    uint32_t *r = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr) : "memory");
}


// ---------------------------------------------------------------- filter GEMM: shared definitions
//
// The filter GEMM (kmeans_mma.cu: kmeans_filter_kernel) serves two scorers: the k-means
// max / argmax (kmeans_mma.cu) and the FBGMM log_marg_i logsumexp (fixedvar_filter.cu).  Both hand
// it fp16 tile images whose inner dimension already carries every additive constant, and both read
// back the same per-row record: the best three 16-column chunks and member masks.

constexpr int TILE_ROWS = 128;     // rows per operand tile image
constexpr int MT_ROWS = 256;       // embeddings per CTA work item (two A tiles)
constexpr int NT_COLS = 128;       // components per accumulator tile
constexpr int CHUNK = 16;          // components per candidate chunk

struct __align__(32) Cand {        // per-embedding filter record
    float m1, m2, m3;              // best / second / third chunk maximum of the filter score
    int32_t i1, i2;                // chunk ids of m1, m2
    uint32_t masks;                // bits 0-15: members of chunk i1 within tau of its max; 16-31: chunk i2
    int32_t pad[2];
};

// byte offset of element (r, c) inside a 128-row tile image
__host__ __device__ inline int tile_off(int r, int c) {
    return ((c >> 3) * (TILE_ROWS / 8) + (r >> 3)) * 128 + (r & 7) * 16 + (c & 7) * 2;
}

// byte offset of element (r, c) inside a 128-row tile image of 1-byte (fp8) elements: the same 8-row x 16-byte core
// matrices, 16 elements per core-matrix row instead of 8
__host__ __device__ inline int tile_off8(int r, int c) {
    return ((c >> 4) * (TILE_ROWS / 8) + (r >> 3)) * 128 + (r & 7) * 16 + (c & 15);
}

// Rigorous bound on |t^ - t| for t = x.mu - |mu|^2/2 as computed by the fp16 filter GEMM:
//   ex*|mu^| + |x|*e_mu                      fp16 rounding of the two operands (Cauchy-Schwarz)
//   c_acc*((|x|+ex)*|mu^| + |mu|^2/2)        fp32 accumulation in the tensor core, bias split
//   eta = (D+3)*2^-24*(|x|+|mu|)^2/2         the reference's own float32 rounding of the score
// ex = |x - fp16(x)|, nx = |x|, e_mu = max_k |mu_k - fp16(mu_k)|, n_mu = max_k |fp16(mu_k)|.
// A component whose t^ is more than tau = 2*bound below the best t^ cannot be the reference's argmax.
__host__ __device__ inline float filter_tau(float ex, float nx, float e_mu, float n_mu, int D) {
    const int kp = (D + 3 + 15) / 16 * 16;
    const float c_acc = ldexpf((float)kp, -21) + ldexpf(1.f, -19);
    const float eta = 0.5f * ldexpf((float)(D + 3), -24) * (nx + n_mu + e_mu) * (nx + n_mu + e_mu);
    const float bound = ex * n_mu + nx * e_mu + c_acc * ((nx + ex) * n_mu + 0.5f * n_mu * n_mu + 1e-30f) + eta;
    return 2.0f * bound;
}

// The same bound for the fp8 (e4m3) first-level filter.  Operands are scaled by a common power of two s (exact), so
// every quantity below lives in the scaled space (scores x s^2); ex / e_mu are the actual e4m3 rounding-error norms
// measured at packing time, e_bias the largest error of the three-term e4m3 representation of -s^2|mu|^2/2.
// Products of two e4m3 numbers are exact in fp32; the accumulation term C_ACC8 * (sum of |products|) was
// calibrated against float64 dot products of the quantised operands (tools/fp8_acc_microbench.py: the largest
// observed |error| / (|x^||mu^| + |bias|) is 1.0e-7 x KP/32; C_ACC8 leaves a factor > 8).
constexpr float C_ACC8_PER_STEP = 1.0f / 1048576.0f;          // 2^-20 per K = 32 step
// The e4m3 pass keeps its running top-3 as PACKED KEYS (top3_insert_key): the chunk id replaces the low KEY_ID_BITS
// mantissa bits of the chunk maximum, so the record's m1 / m2 / m3 are within 2^(KEY_ID_BITS - 23) |score| of the
// chunk maxima.  |score| <= the magnitude the accumulation term already multiplies, so the keys cost one more
// relative term there (1.05: the magnitude expression bounds the computed score up to its own small error terms).
constexpr int KEY_ID_BITS = 12;                                // chunk ids < 4096: K_max <= 65536 for the e4m3 pass
constexpr uint32_t KEY_ID_MASK = (1u << KEY_ID_BITS) - 1u;
constexpr float C_KEY8 = 1.05f / (float)(1u << (23 - KEY_ID_BITS));
constexpr float KEY_FLOOR = -3.0e38f;                          // NaN chunk maxima enter as this (a key must not be NaN)
constexpr float KEY_FLOOR_TEST = -2.9e38f;                     // keys above this belong to a real chunk
__host__ __device__ inline int kp8_of(int D) { return (D + 3 + 31) / 32 * 32; }
__host__ __device__ inline float filter_tau8(float ex, float nx, float e_mu, float n_mu, float e_bias, float bias_max, int D) {
    const float c_acc = C_ACC8_PER_STEP * (float)(kp8_of(D) / 32) + ldexpf(1.f, -19) + C_KEY8;
    const float eta = 0.5f * ldexpf((float)(D + 3), -24) * (nx + n_mu + e_mu) * (nx + n_mu + e_mu);
    const float bound = ex * n_mu + nx * e_mu + e_bias + c_acc * ((nx + ex) * n_mu + 1.1f * bias_max + 1e-30f) + eta;
    return 2.0f * bound;
}

// Logsumexp filter (FBGMM log_marg_i).  The GEMM computes s^_k ~ s_k = A_k + F(x).W_k, where F are
// the row's features (x, or [x, x*x] for anisotropic variances) and A_k plus, in the isotropic case,
// -p_k/2 |x|^2 ride in split constant columns.  Bound on |s^ - s|:
//   eF*nW + nF*eW                      fp16 rounding of features and weights (Cauchy-Schwarz)
//   c_acc*((nF+eF)*nW + cmag)          fp32 accumulation over KP products
//   2^-20*cmag                         the hi/lo splits of the constants, cmag = |A| + (p/2) nF^2
//   2^-22*nF*nW                        float32 rounding of x*x / P*mu before the fp16 conversion
// w4 = model-wide maxima (eW, nW, |A|, p/2).  Every component with s_k >= max_k s_k - T has
// s^_k >= max s^ - (T + 2*bound) = max s^ - tau; the components below carry at most K*exp(-T) of
// the sum, so a logsumexp over the exact scores of the survivors is exact to K*exp(-T) absolute.
struct W4 { float eW, nW, Aabs, ph; };
__host__ __device__ inline float lse_bound(float eF, float nF, W4 w, int KP) {
    const float c_acc = ldexpf((float)KP, -21) + ldexpf(1.f, -19);
    const float cmag = w.Aabs + w.ph * nF * nF;
    return eF * w.nW + nF * w.eW + c_acc * ((nF + eF) * w.nW + cmag) + ldexpf(cmag, -20) + ldexpf(nF * w.nW, -22) + 1e-30f;
}
__host__ __device__ inline float lse_tau(float eF, float nF, W4 w, int KP, float T) {
    return T + 2.0f * lse_bound(eF, nF, w, KP);
}

// The logsumexp filter with e4m3 operands (isotropic variances).  Scaled space: features sx * x, weights sw * p_k mu_k,
// scores S = sx * sw times the true ones.  Constant columns: A_k as 256 * (a0 + a1), and -p_k/2 |x|^2 as
// (u0 + u1)(v0 + v1) without the u1 v1 product, u = alpha |x|^2, v = -S (p_k / 2) / alpha (two-term e4m3 splits).
// w8 = model-wide maxima measured at packing time:
//   eW |sw W - e4m3|_2, nW |e4m3(sw W)|_2, Aabs |S A|, dA = 256 |a - a0 - a1|, vabs |v|, v1abs |v1|, dv |v - v0 - v1|, sw.
// The row's u-side errors follow from e4m3's relative rounding error 2^-4 (normal range; 2^-10 absolute below it):
//   |u1| <= 2^-4 u (1 + 2^-4) + 2^-10,  |u - u0 - u1| <= 2^-8 u + 2^-9.
struct W8 { float eW, nW, Aabs, dA, vabs, v1abs, dv, sw; };
__host__ __device__ inline float lse_bound8(float eF, float nF, const W8 &w, float sx, float alpha, int D) {
    const int kp = (D + 5 + 16 + 31) / 32 * 32;
    const float c_acc = C_ACC8_PER_STEP * (float)(kp / 32) + ldexpf(1.f, -19) + C_KEY8;     // C_KEY8: packed top-3 keys
    const float u = alpha * (nF / sx) * (nF / sx);                 // nF already carries a 1.0001 factor
    const float u1 = 0.0665f * u + 0.001f, du = 0.0040f * u + 0.002f;
    return eF * w.nW + nF * w.eW + w.dA + u1 * w.v1abs + du * w.vabs + u * w.dv
         + c_acc * ((nF + eF) * w.nW + w.Aabs + (u + u1) * (w.vabs + w.v1abs)) + 1e-30f;
}
// tau in the scaled space: S * T + 2 * bound
__host__ __device__ inline float lse_tau8(float eF, float nF, const W8 &w, float sx, float alpha, int D, float T) {
    return sx * w.sw * T + 2.0f * lse_bound8(eF, nF, w, sx, alpha, D);
}
__host__ __device__ inline int kp8_fv_of(int D) { return (D + 5 + 16 + 31) / 32 * 32; }   // + 5 constants + 16 "dead row" columns

// L2 prefetch of one float32 row by the 8 lanes of a group: lane j touches byte 64 j (+ 512 j' for longer rows).
// fv_refine_kernel pulls the NEXT row's embedding into L2 this way (no registers): 4.7 -> 3.9 ms at 21M rows; the
// k-means refine, which already requests its row before looking at the record, did not gain (3.06 -> 3.19 ms).
__device__ __forceinline__ void prefetch_row_l2(const float *xr, int D, int j) {
    const char *p = reinterpret_cast<const char *>(xr);
    for (int off = 64 * j; off < 4 * D; off += 512) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
}

// What the refine does with a row: -2 exhaustive exact scan needed (the third-best chunk is still
// inside the bound, or the record is unusable), -1 the best chunk suffices, >= 0 also visit that chunk.
__device__ __forceinline__ int refine_decide(const Cand &c, float tau, int n_chunks) {
    if (c.i1 < 0 || c.i1 >= n_chunks || !(tau < CUDART_INF_F) ||
        (!(c.m1 - c.m3 > tau) && (c.m3 > -CUDART_INF_F))) return -2;
    if ((c.i2 >= 0) && !(c.m1 - c.m2 > tau)) return c.i2;
    return -1;
}

// Maximum of a 16-score chunk as a depth-3 tree of 3-input maxima (8 FMNMX3/FMNMX like the linear chain, but the
// epilogue is latency-sensitive: four chunks x a chain of eight dependent instructions leave the schedulers idle).
__device__ __forceinline__ float chunk_max16(const float *v) {
    const float a0 = fmaxf(fmaxf(v[0], v[1]), v[2]), a1 = fmaxf(fmaxf(v[3], v[4]), v[5]);
    const float a2 = fmaxf(fmaxf(v[6], v[7]), v[8]), a3 = fmaxf(fmaxf(v[9], v[10]), v[11]);
    const float a4 = fmaxf(fmaxf(v[12], v[13]), v[14]);
    const float b0 = fmaxf(fmaxf(a0, a1), a2), b1 = fmaxf(fmaxf(a3, a4), v[15]);
    return fmaxf(b0, b1);
}

// the same tree ending in a 3-input maximum with KEY_FLOOR: never NaN (fmaxf drops NaN operands), same instruction count
__device__ __forceinline__ float chunk_max16_floor(const float *v) {
    const float a0 = fmaxf(fmaxf(v[0], v[1]), v[2]), a1 = fmaxf(fmaxf(v[3], v[4]), v[5]);
    const float a2 = fmaxf(fmaxf(v[6], v[7]), v[8]), a3 = fmaxf(fmaxf(v[9], v[10]), v[11]);
    const float a4 = fmaxf(fmaxf(v[12], v[13]), v[14]);
    const float b0 = fmaxf(fmaxf(a0, a1), a2), b1 = fmaxf(fmaxf(a3, a4), v[15]);
    return fmaxf(fmaxf(b0, b1), KEY_FLOOR);
}

// Insert a chunk maximum into the running top-3.  A chunk that becomes the new BEST also records
// which of its 16 members lie within tau_c of the chunk maximum (the only ones the refine has to
// score); a chunk entering as runner-up keeps all 16 (it is only visited for the rare rows whose
// runner-up chunk is inside the bound).  Only the new-best case is a (divergent) branch; the
// runner-up / third-place updates are selects, so the common path is straight-line code.
__device__ __forceinline__ void top3_insert(const float *vv, float cm, int cid, float tau_c, float &m1, float &m2,
                                            float &m3, int &i1, int &i2, uint32_t &k1, uint32_t &k2) {
    const bool p2 = cm > m2;
    m3 = p2 ? m2 : fmaxf(m3, cm);
    if (cm > m1) {
        // member j is kept iff vv[j] >= cm - tau_c, i.e. the sign bit of (vv[j] - thr) is clear; the
        // sign bits are funnel-shifted into two 8-bit chains (member 15 first, so bit j = member j)
        const float thr = cm - tau_c;
        uint32_t hi = 0, lo = 0;
#pragma unroll
        for (int j = 7; j >= 0; --j) {
            hi = __funnelshift_l(__float_as_uint(vv[8 + j] - thr), hi, 1);
            lo = __funnelshift_l(__float_as_uint(vv[j] - thr), lo, 1);
        }
        m2 = m1; i2 = i1; k2 = k1;
        m1 = cm; i1 = cid; k1 = ~((hi << 8) | lo) & 0xffffu;
    } else {
        m2 = p2 ? cm : m2; i2 = p2 ? cid : i2; k2 = p2 ? 0xffffu : k2;
    }
}

// same insertion with the member mask already known (merging two partial top-3 lists)
__device__ __forceinline__ void top3_merge(float cm, int cid, uint32_t mk, float &m1, float &m2, float &m3, int &i1,
                                           int &i2, uint32_t &k1, uint32_t &k2) {
    if (cm > m3) {
        if (cm > m2) {
            m3 = m2;
            if (cm > m1) { m2 = m1; i2 = i1; k2 = k1; m1 = cm; i1 = cid; k1 = mk; }
            else { m2 = cm; i2 = cid; k2 = mk; }
        } else m3 = cm;
    }
}

// The e4m3 pass defers the member mask: a chunk that becomes the new best only parks its 16 scores in a per-thread
// shared-memory slot (four 16-byte stores; snap[q * SNAP_STRIDE] with consecutive threads 16 bytes apart:
// conflict-free), and the mask is formed once per row and work item by top3_snapshot_mask.  A new-best event costs
// the WHOLE warp ~8 instructions instead of ~38 (16 x (FADD + SHF)) -- and with 32 independent rows per warp such
// events hit 45 % of all chunk visits.  A best chunk that is demoted to runner-up keeps all 16 members.
constexpr int SNAP_STRIDE = 512;      // float4 slots per quarter: one per epilogue thread
// The e4m3 pass's insertion: the running top-3 are three PACKED KEYS a1 >= a2 >= a3 -- the chunk maximum with the
// chunk id in its low KEY_ID_BITS mantissa bits (filter_tau8 / lse_bound8 carry the 2^-11 relative term this costs).
// Keys of different chunks differ, a key orders like its maximum, and the id travels with the value, so the update
// is a five-instruction min / max network instead of compares and selects on (value, id) pairs; a new best parks its
// scores in the snapshot slot (four predicated stores).  cm must not be NaN (chunk_max16_floor).
__device__ __forceinline__ void top3_insert_key(const float *vv, float cm, uint32_t cid, float4 *snap, float &a1, float &a2,
                                                float &a3) {
    const float key = __uint_as_float((__float_as_uint(cm) & ~KEY_ID_MASK) | cid);
    if (key > a1) {
        snap[0 * SNAP_STRIDE] = make_float4(vv[0], vv[1], vv[2], vv[3]);
        snap[1 * SNAP_STRIDE] = make_float4(vv[4], vv[5], vv[6], vv[7]);
        snap[2 * SNAP_STRIDE] = make_float4(vv[8], vv[9], vv[10], vv[11]);
        snap[3 * SNAP_STRIDE] = make_float4(vv[12], vv[13], vv[14], vv[15]);
    }
    const float t1 = fminf(a1, key);
    a1 = fmaxf(a1, key);
    const float t2 = fminf(a2, t1);
    a2 = fmaxf(a2, t1);
    a3 = fmaxf(a3, t2);
}
// (value, chunk id) of a key; a slot that never saw a real chunk (-inf, or the floor of an all-NaN chunk) -> (-inf, -1)
__device__ __forceinline__ void key_unpack(float a, float &m, int &i) {
    const bool real = a > KEY_FLOOR_TEST;
    m = real ? a : -CUDART_INF_F;
    i = real ? (int)(__float_as_uint(a) & KEY_ID_MASK) : -1;
}
// members of the parked best chunk within tau_c of its maximum m1 (bit j = member j)
__device__ __forceinline__ uint32_t top3_snapshot_mask(const float4 *snap, float m1, float tau_c) {
    const float thr = m1 - tau_c;
    uint32_t bits = 0;
#pragma unroll
    for (int qd = 3; qd >= 0; --qd) {
        const float4 f = snap[qd * SNAP_STRIDE];
        bits = __funnelshift_l(__float_as_uint(f.w - thr), bits, 1);
        bits = __funnelshift_l(__float_as_uint(f.z - thr), bits, 1);
        bits = __funnelshift_l(__float_as_uint(f.y - thr), bits, 1);
        bits = __funnelshift_l(__float_as_uint(f.x - thr), bits, 1);
    }
    return ~bits & 0xffffu;
}

// What survives of a filter record once the row's threshold has been applied: the chunks to visit, the
// member masks and refine_decide's code.  16 bytes.
struct __align__(16) RowRec { int32_t i1, i2; uint32_t masks; int32_t code; };

constexpr int TAU_KMEANS = 0, TAU_LSE = 1, TAU_KMEANS_FP8 = 2, TAU_LSE_FP8 = 3;   // KMEANS_FP8: w_max = (e_mu, n_mu, e_bias, bias_max); LSE_FP8: w_max = W8

// Launch description of the filter GEMM over pre-packed tile images (host side).
struct FilterLaunch {
    const void *x_tiles, *w_tiles;     // [rows_pad / 128] and [w_rows_pad / 128] tile images of KP columns
    void *cand;                        // [n_emb] Cand records out
    int64_t n_emb;
    int32_t w_rows_pad;                // multiple of NT_COLS
    int32_t w_rows = 0;                // rows that are not padding (0: unknown, all of w_rows_pad): the last
                                       // accumulator tile is computed only as wide as they reach
    int32_t KP;                        // padded inner dimension of ONE chunk (multiple of 16)
    int32_t n_chunks;                  // 1, or 2: every tile is two consecutive chunk images of KP columns
    int32_t D;
    const float *x_max, *w_max;
    int32_t tau_kind;
    float tau_T;
    int32_t fp8 = 0;                   // operands are e4m3 tile images (KP bytes per row), kind::f8f6f4
    float sx = 1.f, alpha = 1.f;       // TAU_LSE_FP8: feature scale and |x|^2 scale
};
int launch_filter(const FilterLaunch &f, cudaStream_t stream);

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (load + wait in one statement)
__device__ __forceinline__ void tc_ld32_wait(uint32_t taddr, float *v) {
    uint32_t *r = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}

}  // namespace mma
}  // namespace segb
