// Tensor-core scoring for the frozen k-means sweep: filter GEMM + exact refine.
//
// Replaces, for ALL embeddings at once, the reference's per-item
//   max_k / argmax_k  -sum_d (means[k,d] - x[d])^2
// (KMeansComponents.neg_sqrd_norm / max_ / argmax_neg_sqrd_norm_i,
//  segmentalist/kmeans_components.py:225-232, called from
//  kmeans_acoustic_wordseg.py:334-351 and kmeans.py:141-143).
//
// Exactness strategy (SURVEY.md 7.3 "filter-and-refine"): the reference decides in
// float32 with NumPy's summation order, which no tensor-core pass reproduces.
// So the GEMM only FILTERS: one fp16 tcgen05 pass computes t^[m,k] ~ x.mu_k - |mu_k|^2/2
// (argmax_k t = argmax_k of the reference score) and keeps, per embedding, the
// best three 16-component chunks.  A rigorous per-row error bound then decides
// which chunks can contain the true winner; those (normally 16 components out
// of K_max) are re-scored by the exact float32 routine the SIMT path uses, so
// max and first-argmax come out bit-identical to the reference.  Rows whose
// third-best chunk is still inside the bound go to a SECOND-LEVEL tensor pass (compact
// image of those rows, bitmap epilogue, exact re-score of the flagged components), and only
// rows that even that cannot resolve (NaN data) to a full exact scan.  The default first
// level runs in e4m3 (segb_mma8_*: scaled operands, its own rigorous bound), with the fp16
// pass as its second level: a cascade e4m3 -> fp16 -> exact.
//
// Memory layout (designed for the copy engine, not for humans): X and the means
// are stored in HBM as fp16 "tile images": consecutive 128-row tiles, each
// already in the UMMA canonical K-major no-swizzle shared-memory layout (8x8
// core matrices of 128 contiguous bytes).  A tile is therefore ONE contiguous
// cp.async.bulk (TMA bulk copy, SASS UBLKCP) into shared memory, with no tensor
// map and no swizzle agreement to get wrong.  The inner dimension is padded from
// D to KP = roundup(D + 3, 16); three padding columns carry a 3-way fp16 split
// of -|mu_k|^2/2 (x side: 1.0), so the norm rides along in the GEMM for free.
//
// Kernel: persistent, one CTA per SM, 640 threads, warp-specialised:
//   warp 0  TMA producer  (A = 256 embeddings per work item, double-buffered; B = 128-component
//                          tiles, L2-resident, a ring of 2 (fp16) or 4 (e4m3) stages)
//   warp 1  MMA issuer    (one elected thread; tcgen05.mma kind::f16 or kind::f8f6f4, M=128 N=128, 32 bytes of
//                          K per row and instruction, fp32 accumulate in TMEM, 2 row-halves x double-buffered
//                          accumulators = 512 TMEM columns; unrolled issue loop, ONE commit per tile; the last
//                          tile only as wide as the components reach)
//   warp 2  TMEM allocator
//   warps 4-19 epilogue   (two SETS of eight warps drain alternate tiles: set s owns accumulator pair s and takes
//                          its tile in two tcgen05.ld 32x32b.x64 -- one thread owns one embedding row; running
//                          top-3 of chunk maxima in registers; nothing but 32 B per embedding ever goes back to
//                          HBM.  Variants: EPI = 1 writes a per-row bitmap of the components above the row's
//                          threshold instead (second-level pass over undecided rows, row count on the device);
//                          F8 defers the member mask of the best chunk through a shared-memory snapshot.)
// The same kernel is the first level of the FBGMM log_marg_i filter (fixedvar_filter.cu: logsumexp thresholds,
// NCH = 2 inner-dimension chunks for anisotropic variances).
#include <cuda_fp8.h>
#include "mma_common.cuh"
#include "refine_rows.cuh"

namespace segb {
namespace mma {

constexpr uint32_t KSTEP_BYTES = 2 * (TILE_ROWS / 8) * 128;        // one K=16 step of a 128-row tile image
constexpr int EPI_PARTS = 2;         // column halves of an accumulator tile handled by separate warp sets
constexpr int N_EPI_WARPS = 8 * EPI_PARTS;
constexpr int N_THREADS = 128 + 32 * N_EPI_WARPS;
constexpr uint32_t TMEM_COLS = 512;
constexpr float PAD_BIAS = -30000.0f;   // -|mu|^2/2 stand-in for padded components: never wins

__host__ __device__ inline int kp_of(int D) { return (D + 3 + 15) / 16 * 16; }
__host__ __device__ inline int64_t tile_bytes_of(int D) { return (int64_t)TILE_ROWS * kp_of(D) * 2; }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// SBO = distance between 8-row core matrices (128 B), LBO = distance between the two
// 8-element K chunks of one K=16 instruction ((TILE_ROWS/8)*128 B).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    const uint64_t lbo = (TILE_ROWS / 8) * 128, sbo = 128;
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: A,B = fp16 K-major, D = fp32, M = 128, N = 128
__device__ __forceinline__ uint32_t make_idesc(int n_cols = NT_COLS) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(TILE_ROWS >> 4) << 24);
}

// ---------------------------------------------------------------- filter GEMM

struct FilterParams {
    const uint8_t *x_tiles, *w_tiles;
    Cand *cand;
    int64_t n_emb;
    int32_t n_mtiles, n_ntiles, n_ksteps, n_abuf;     // n_ksteps: K=16 steps per inner-dimension CHUNK
    int32_t n_bstage;              // NCH = 1: B ring depth (2, 4 or 8 tiles; a power of two).  NCH = 2: two stages = the two chunks
    int32_t n_last_cols;           // MMA width of the LAST accumulator tile (multiple of 16): padding components beyond it
                                   // are never multiplied, and the epilogue skips their (stale) TMEM columns
    int32_t n_chunks_valid;        // 16-component chunks that were computed = (n_ntiles - 1) * 8 + n_last_cols / 16
    int32_t n_valid_last;          // real components in the last computed chunk (1..16); F8 masks the padding ones
    uint32_t tile_bytes;                              // bytes of one 128-row tile image of one chunk
    const float *x_max, *w_max;    // corpus-wide (max ex, max nx); model-wide maxima (k-means: e_mu, n_mu; FBGMM: eB, nB, |A|, p/2)
    int32_t D;
    int32_t tau_kind;              // TAU_KMEANS: argmax filter (2 * bound); TAU_LSE: logsumexp filter (tau_T + 2 * bound)
    float tau_T;
    float sx, alpha;               // TAU_LSE_FP8: feature scale and |x|^2 scale of the e4m3 images
    // second-level pass over the rows the top-3 records could not decide (EPI = 1): the row count lives on the
    // device, every row has its own threshold and gets a bitmap of ALL the components at or above it
    const unsigned long long *n_rows_dev;   // rows = clamp(*n_rows_dev - rows_first, 0, rows_cap): one round of the undecided list
    int64_t rows_cap, rows_first;
    const float *thr;              // [rows_cap] per-row threshold (best filter score - tau of the first pass)
    uint32_t *bitmap;              // [rows_cap][n_ntiles * 4] bit k of a row: filter score of component k >= thr
};

// KS = number of K=16 steps per chunk when known at compile time (fully unrolled issue loop), 0 = runtime.
// NCH = chunks of the inner dimension (1: k-means and isotropic FBGMM; 2: anisotropic FBGMM, whose
// inner dimension [x | x*x] is twice as long).  With NCH = 2 an operand tile is two consecutive
// chunk images; the B ring holds one chunk per stage (stage = chunk index), the accumulators stay
// double-buffered by tile parity, and the issuer commits twice per tile: "chunk-0 stage free" and
// "tile done" (= chunk-1 stage free + accumulators ready).
// EPI = 0: running top-3 chunk maxima per row (Cand records); EPI = 1: per-row bitmap of the components whose
// score reaches the row's threshold (second-level pass, device-side row count).
// F8: e4m3 operands (kind::f8f6f4; a K step is 32 elements = the same 32 bytes per row), NCH = 1 only.
template <int KS, int NCH, int EPI = 0, bool F8 = false>
__global__ void __launch_bounds__(N_THREADS, 1) kmeans_filter_kernel(FilterParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (EPI == 1) {
        const int64_t left = (int64_t)*p.n_rows_dev - p.rows_first;
        p.n_emb = left < 0 ? 0 : (left < p.rows_cap ? left : p.rows_cap);
        p.n_mtiles = (int32_t)((p.n_emb + MT_ROWS - 1) / MT_ROWS);
    }
    const uint32_t tb = p.tile_bytes;
    const int n_ks = KS > 0 ? KS : p.n_ksteps;
    uint8_t *sA = smem;                                        // n_abuf x 2 tiles x NCH chunks
    const uint32_t NB = NCH == 1 ? (uint32_t)p.n_bstage : 2u;  // B ring: tile n_use lives in stage n_use % NB, its accumulators in pair n_use % 2
    const uint32_t nb_sh = NB == 8 ? 3u : (NB == 4 ? 2u : 1u);
    uint8_t *sB = smem + (size_t)p.n_abuf * 2 * NCH * tb;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + NB * (size_t)tb);
    // MMA_DONE[s] (one tcgen05.commit per tile, s = the tile's B stage) means BOTH "B stage s may be refilled" (producer) and
    // "accumulator pair b is complete" (epilogue): the issuing thread pays for one commit per tile.
    // NCH = 2 adds B_EMPTY0: the chunk-0 stage is free as soon as the tile's chunk-0 MMAs are done.
    // With the e4m3 operands a tile's MMAs take 640 clocks, less than the L2 round trip of the next B tile: the ring is
    // then deeper than the two accumulator pairs (measured: 2 stages 18.5 ms, see profiles/r2_experiments.md).
    constexpr int A_FULL = 0, A_EMPTY = 2, B_FULL = 4, MMA_DONE = 12, ACC_EMPTY = 20, B_EMPTY0 = 22, N_BARS = 23;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + N_BARS);
    float4 *merge = reinterpret_cast<float4 *>(reinterpret_cast<uint8_t *>(bars) + 256);   // [MT_ROWS][2] partial top-3
    // F8: parked scores of every epilogue thread's current best chunk (top3_insert_key), 4 x 512 float4 = 32 KB
    float4 *snap_all = reinterpret_cast<float4 *>(reinterpret_cast<uint8_t *>(bars) + 256 + (size_t)MT_ROWS * 32);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < ACC_EMPTY; ++i) mbar_init(BAR(i), 1);
        // top-3 epilogue: accumulator pair b is drained by the 8 warps of set b; bitmap epilogue: by all 16 warps
        for (int b = 0; b < 2; ++b) mbar_init(BAR(ACC_EMPTY + b), EPI == 1 ? N_EPI_WARPS : N_EPI_WARPS / 2);
        mbar_init(BAR(B_EMPTY0), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // A (256 embeddings) is double-buffered: the next work item's tiles are requested while the
        // current one is still being multiplied, so the tensor pipe does not drain at work-item
        // boundaries.  B (one 128-component tile, L2-resident) alternates between two stages.
        if (elect_one()) {
            auto load_a = [&](int it, int mt) {
                const int ab = p.n_abuf == 2 ? (it & 1) : 0;
                const uint32_t use = p.n_abuf == 2 ? (uint32_t)(it >> 1) : (uint32_t)it;
                mbar_wait(BAR(A_EMPTY + ab), (use & 1) ^ 1);
                mbar_expect_tx(BAR(A_FULL + ab), 2 * NCH * tb);
#pragma unroll
                for (int j = 0; j < 2 * NCH; ++j)          // tiles 2mt, 2mt+1, each NCH consecutive chunk images
                    bulk_g2s(smem_u32(sA + (size_t)(2 * NCH * ab + j) * tb),
                             p.x_tiles + ((size_t)(2 * mt) * NCH + j) * tb, tb, BAR(A_FULL + ab));
            };
            uint32_t n_use = 0;
            int it = 0;
            if ((int)blockIdx.x < p.n_mtiles) load_a(0, blockIdx.x);
            for (int mt = blockIdx.x; mt < p.n_mtiles; mt += gridDim.x, ++it) {
                const int mt_next = mt + gridDim.x;
                const int nt_pref = p.n_abuf == 2 ? (p.n_ntiles > 4 ? 4 : p.n_ntiles - 1) : -1;
                for (int nt = 0; nt < p.n_ntiles; ++nt, ++n_use) {
                    if (nt == nt_pref && mt_next < p.n_mtiles) load_a(it + 1, mt_next);
                    if (NCH == 1) {
                        const uint32_t s = n_use & (NB - 1), use = n_use >> nb_sh;
                        mbar_wait(BAR(MMA_DONE + s), (use & 1) ^ 1);       // the MMAs of the stage's previous tile are done
                        mbar_expect_tx(BAR(B_FULL + s), tb);
                        bulk_g2s(smem_u32(sB + (size_t)s * tb), p.w_tiles + (size_t)nt * tb, tb, BAR(B_FULL + s));
                    } else {
                        // chunk 0 -> stage 0: free when the previous tile's chunk-0 MMAs are done
                        mbar_wait(BAR(B_EMPTY0), (n_use & 1) ^ 1);
                        mbar_expect_tx(BAR(B_FULL + 0), tb);
                        bulk_g2s(smem_u32(sB), p.w_tiles + (size_t)(2 * nt) * tb, tb, BAR(B_FULL + 0));
                        // chunk 1 -> stage 1: free when the previous tile is done (its MMA_DONE completion)
                        if (n_use > 0) mbar_wait(BAR(MMA_DONE + ((n_use - 1) & 1)), ((n_use - 1) >> 1) & 1);
                        mbar_expect_tx(BAR(B_FULL + 1), tb);
                        bulk_g2s(smem_u32(sB + tb), p.w_tiles + (size_t)(2 * nt + 1) * tb, tb, BAR(B_FULL + 1));
                    }
                }
                if (p.n_abuf == 1 && mt_next < p.n_mtiles) load_a(it + 1, mt_next);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The issue loop is the kernel's pacemaker: one M=128 N=128 K=16 MMA occupies the tensor pipe
        // for 64 clocks, and everything the elected thread does between two tiles (barrier polls, the
        // commit) has to fit into the few MMAs still queued.  So: descriptors are 32-bit low words
        // advanced by a constant per K step, the K loop is unrolled at compile time, the issuer is
        // chosen with elect.sync (see elect_one()), both barriers of a tile are polled together, and
        // there is ONE commit per tile.
        if (elect_one()) {
            const uint32_t idesc_full = make_idesc(), idesc_last = make_idesc(p.n_last_cols);
            constexpr uint32_t KSTEP = KSTEP_BYTES >> 4;              // descriptor units per K=16 step
            const uint32_t a_lo_base = make_desc_lo(smem_u32(sA), TILE_ROWS);
            const uint32_t b_lo_base = make_desc_lo(smem_u32(sB), TILE_ROWS);
            uint32_t n_use = 0;
            int it = 0;
            for (int mt = blockIdx.x; mt < p.n_mtiles; mt += gridDim.x, ++it) {
                const int ab = p.n_abuf == 2 ? (it & 1) : 0;
                const uint32_t ause = p.n_abuf == 2 ? (uint32_t)(it >> 1) : (uint32_t)it;
                mbar_wait(BAR(A_FULL + ab), ause & 1);
                const uint32_t a_lo0 = a_lo_base + (uint32_t)(2 * NCH * ab) * (tb >> 4), a_lo1 = a_lo0 + NCH * (tb >> 4);
                for (int nt = 0; nt < p.n_ntiles; ++nt, ++n_use) {
                    const uint32_t buf = n_use & 1, use = n_use >> 1;
                    const uint32_t d0 = tmem_base + (buf * 2) * NT_COLS, d1 = d0 + NT_COLS;
                    const uint32_t idesc = (nt == p.n_ntiles - 1) ? idesc_last : idesc_full;
                    if (NCH == 1) {
                        const uint32_t s = n_use & (NB - 1);
                        mbar_wait2(BAR(B_FULL + s), (n_use >> nb_sh) & 1, BAR(ACC_EMPTY + buf), (use & 1) ^ 1);
                        tc_fence_after();
                        const uint32_t b_lo = b_lo_base + s * (tb >> 4);
                        if (F8 && KS > 0) {
                            tc_mma_f8_lo<false>(d0, a_lo0, b_lo, idesc);
                            tc_mma_f8_lo<false>(d1, a_lo1, b_lo, idesc);
#pragma unroll
                            for (int k = 1; k < KS; ++k) {
                                tc_mma_f8_lo<true>(d0, a_lo0 + k * KSTEP, b_lo + k * KSTEP, idesc);
                                tc_mma_f8_lo<true>(d1, a_lo1 + k * KSTEP, b_lo + k * KSTEP, idesc);
                            }
                        } else if (F8) {
                            for (int k = 0; k < n_ks; ++k) {
                                tc_mma_f8(d0, ((uint64_t)DESC_HI << 32) | (a_lo0 + k * KSTEP),
                                          ((uint64_t)DESC_HI << 32) | (b_lo + k * KSTEP), idesc, k > 0 ? 1u : 0u);
                                tc_mma_f8(d1, ((uint64_t)DESC_HI << 32) | (a_lo1 + k * KSTEP),
                                          ((uint64_t)DESC_HI << 32) | (b_lo + k * KSTEP), idesc, k > 0 ? 1u : 0u);
                            }
                        } else if (KS > 0) {
                            tc_mma_f16_lo<false>(d0, a_lo0, b_lo, idesc);
                            tc_mma_f16_lo<false>(d1, a_lo1, b_lo, idesc);
#pragma unroll
                            for (int k = 1; k < KS; ++k) {
                                tc_mma_f16_lo<true>(d0, a_lo0 + k * KSTEP, b_lo + k * KSTEP, idesc);
                                tc_mma_f16_lo<true>(d1, a_lo1 + k * KSTEP, b_lo + k * KSTEP, idesc);
                            }
                        } else {
                            for (int k = 0; k < n_ks; ++k) {
                                tc_mma_f16(d0, ((uint64_t)DESC_HI << 32) | (a_lo0 + k * KSTEP),
                                           ((uint64_t)DESC_HI << 32) | (b_lo + k * KSTEP), idesc, k > 0 ? 1u : 0u);
                                tc_mma_f16(d1, ((uint64_t)DESC_HI << 32) | (a_lo1 + k * KSTEP),
                                           ((uint64_t)DESC_HI << 32) | (b_lo + k * KSTEP), idesc, k > 0 ? 1u : 0u);
                            }
                        }
                        tc_commit(BAR(MMA_DONE + s));         // B stage free + accumulators ready
                    } else {
                        // chunk 0 (stage 0) opens the accumulators, chunk 1 (stage 1) completes them
                        mbar_wait2(BAR(B_FULL + 0), n_use & 1, BAR(ACC_EMPTY + buf), (use & 1) ^ 1);
                        tc_fence_after();
                        if (KS > 0) {
                            tc_mma_f16_lo<false>(d0, a_lo0, b_lo_base, idesc);
                            tc_mma_f16_lo<false>(d1, a_lo1, b_lo_base, idesc);
#pragma unroll
                            for (int k = 1; k < KS; ++k) {
                                tc_mma_f16_lo<true>(d0, a_lo0 + k * KSTEP, b_lo_base + k * KSTEP, idesc);
                                tc_mma_f16_lo<true>(d1, a_lo1 + k * KSTEP, b_lo_base + k * KSTEP, idesc);
                            }
                        } else {
                            for (int k = 0; k < n_ks; ++k) {
                                tc_mma_f16(d0, ((uint64_t)DESC_HI << 32) | (a_lo0 + k * KSTEP),
                                           ((uint64_t)DESC_HI << 32) | (b_lo_base + k * KSTEP), idesc, k > 0 ? 1u : 0u);
                                tc_mma_f16(d1, ((uint64_t)DESC_HI << 32) | (a_lo1 + k * KSTEP),
                                           ((uint64_t)DESC_HI << 32) | (b_lo_base + k * KSTEP), idesc, k > 0 ? 1u : 0u);
                            }
                        }
                        tc_commit(BAR(B_EMPTY0));
                        mbar_wait(BAR(B_FULL + 1), n_use & 1);
                        tc_fence_after();
                        const uint32_t a_c0 = a_lo0 + (tb >> 4), a_c1 = a_lo1 + (tb >> 4), b_c = b_lo_base + (tb >> 4);
                        if (KS > 0) {
#pragma unroll
                            for (int k = 0; k < KS; ++k) {
                                tc_mma_f16_lo<true>(d0, a_c0 + k * KSTEP, b_c + k * KSTEP, idesc);
                                tc_mma_f16_lo<true>(d1, a_c1 + k * KSTEP, b_c + k * KSTEP, idesc);
                            }
                        } else {
                            for (int k = 0; k < n_ks; ++k) {
                                tc_mma_f16(d0, ((uint64_t)DESC_HI << 32) | (a_c0 + k * KSTEP),
                                           ((uint64_t)DESC_HI << 32) | (b_c + k * KSTEP), idesc, 1u);
                                tc_mma_f16(d1, ((uint64_t)DESC_HI << 32) | (a_c1 + k * KSTEP),
                                           ((uint64_t)DESC_HI << 32) | (b_c + k * KSTEP), idesc, 1u);
                            }
                        }
                        tc_commit(BAR(MMA_DONE + buf));       // chunk-1 stage free + accumulators ready
                    }
                }
                tc_commit(BAR(A_EMPTY + ab));             // A tiles free
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: running top-3 chunk maxima per embedding =====================
        // warp -> (TMEM lane quadrant q = warp % 4, row half h, column part): 16 warps, each draining a
        // 32-row x 64-column block of every accumulator tile with ONE tcgen05.ld; the accumulator is
        // handed back to the MMA warp as soon as the values are in registers, before they are reduced.
        const int e = warp - 4, q = warp & 3, h = (e >> 2) & 1, part = e >> 3;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        constexpr int COLS = NT_COLS / EPI_PARTS;          // columns of each tile this warp reduces
        static_assert(COLS == 64, "one x64 load per warp and tile");
        // one threshold for the whole launch: the loosest per-row tau (refine re-derives the exact per-row one)
        const float tau_c = p.tau_kind == TAU_KMEANS
            ? filter_tau(p.x_max[0], p.x_max[1], p.w_max[0], p.w_max[1], p.D)
            : p.tau_kind == TAU_KMEANS_FP8
            ? filter_tau8(p.x_max[0], p.x_max[1], p.w_max[0], p.w_max[1], p.w_max[2], p.w_max[3], p.D)
            : p.tau_kind == TAU_LSE_FP8
            ? lse_tau8(p.x_max[0], p.x_max[1], W8{p.w_max[0], p.w_max[1], p.w_max[2], p.w_max[3], p.w_max[4], p.w_max[5], p.w_max[6], p.w_max[7]},
                       p.sx, p.alpha, p.D, p.tau_T)
            : lse_tau(p.x_max[0], p.x_max[1], W4{p.w_max[0], p.w_max[1], p.w_max[2], p.w_max[3]}, 16 * n_ks * NCH, p.tau_T);
        uint32_t n_use = 0;
        if (EPI == 1) {
            // bit j of word w of a row = component 32 w + j reaches the row's threshold: the sign bits of
            // (score - thr) funnel-shifted into two words per tile and thread, 8 bytes per tile to HBM
            for (int mt = blockIdx.x; mt < p.n_mtiles; mt += gridDim.x) {
                const int64_t row = (int64_t)mt * MT_ROWS + h * TILE_ROWS + q * 32 + lane;
                const float thr = p.thr[row];
                uint2 *out = reinterpret_cast<uint2 *>(p.bitmap + row * (int64_t)(p.n_ntiles * 4)) + part;
                for (int nt = 0; nt < p.n_ntiles; ++nt, ++n_use) {
                    const uint32_t buf = n_use & 1, acc_phase = (n_use >> 1) & 1;
                    mbar_wait(BAR(MMA_DONE + (NCH == 1 ? (n_use & (NB - 1)) : buf)), NCH == 1 ? (n_use >> nb_sh) & 1 : acc_phase);
                    tc_fence_after();
                    float v[64];
                    tc_ld64_wait(tmem_base + lane_base + (buf * 2 + h) * NT_COLS + part * COLS, v);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(ACC_EMPTY + buf));
                    uint32_t w0 = 0, w1 = 0;
#pragma unroll
                    for (int j = 31; j >= 0; --j) {
                        w0 = __funnelshift_l(__float_as_uint(v[j] - thr), w0, 1);
                        w1 = __funnelshift_l(__float_as_uint(v[32 + j] - thr), w1, 1);
                    }
                    // columns past the last tile's MMA width hold an earlier tile's scores
                    const int n_cols = (p.n_chunks_valid - (nt * (NT_COLS / CHUNK) + (part * COLS) / CHUNK)) * CHUNK;
                    const uint32_t k0 = n_cols >= 32 ? 0xffffffffu : (n_cols <= 0 ? 0u : ((1u << n_cols) - 1u));
                    const uint32_t k1 = n_cols >= 64 ? 0xffffffffu : (n_cols <= 32 ? 0u : ((1u << (n_cols - 32)) - 1u));
                    out[2 * nt] = make_uint2(~w0 & k0, ~w1 & k1);
                }
            }
        } else {
            // Two SETS of eight warps drain alternate tiles: set s owns accumulator pair s (= tiles of parity s) and takes
            // all 128 columns of its tile in two 64-column loads.  While one set waits for its tcgen05.ld the other one
            // is reducing, so TMEM reads and issue slots overlap instead of alternating for all 16 warps at once (the
            // e4m3 pass is bound by this epilogue, not by the MMAs).  The sets' partial top-3 lists of a row meet through
            // shared memory at the end of every work item.
            const int set = part;                              // e >> 3
            float4 *snap = snap_all + (threadIdx.x - 128);     // this thread's slot (epilogue threads are 128 .. 639)
            for (int mt = blockIdx.x; mt < p.n_mtiles; mt += gridDim.x) {
                float m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F;
                int i1 = -1, i2 = -1;
                uint32_t k1 = 0, k2 = 0;
                for (int nt = 0; nt < p.n_ntiles; ++nt, ++n_use) {
                    if ((int)(n_use & 1) != set) continue;
                    const uint32_t buf = n_use & 1, acc_phase = (n_use >> 1) & 1;
                    mbar_wait(BAR(MMA_DONE + (NCH == 1 ? (n_use & (NB - 1)) : buf)), NCH == 1 ? (n_use >> nb_sh) & 1 : acc_phase);
                    tc_fence_after();
                    const bool last_tile = nt + 1 == p.n_ntiles;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        float v[64];
                        tc_ld64_wait(tmem_base + lane_base + (buf * 2 + h) * NT_COLS + half * COLS, v);
                        if (half == 1) {                       // both halves are in registers: hand the accumulator back
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(BAR(ACC_EMPTY + buf));
                        }
                        const int cid0 = nt * (NT_COLS / CHUNK) + (half * COLS) / CHUNK;
                        if (!last_tile) {
                            // every tile but the last: four full chunks, straight-line
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                // F8: m1 >= m2 >= m3 are packed keys (top3_insert_key), unpacked after the last tile
                                if (F8) top3_insert_key(&v[c * 16], chunk_max16_floor(&v[c * 16]), (uint32_t)(cid0 + c), snap, m1, m2, m3);
                                else top3_insert(&v[c * 16], chunk_max16(&v[c * 16]), cid0 + c, tau_c, m1, m2, m3, i1, i2, k1, k2);
                            }
                        } else {
                            const int n_c = p.n_chunks_valid - cid0;         // chunks of the last tile that were computed
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                if (c < n_c) {
                                    float cm = v[c * 16];
                                    if (F8 && cid0 + c == p.n_chunks_valid - 1 && p.n_valid_last < CHUNK) {
                                        // e4m3 cannot carry a "never wins" bias: the padding components of the last chunk
                                        // are left out of its maximum (their member bits may be set; the refine skips
                                        // k >= K_max)
#pragma unroll
                                        for (int j = 1; j < 16; ++j) cm = fmaxf(cm, j < p.n_valid_last ? v[c * 16 + j] : -CUDART_INF_F);
                                    } else {
#pragma unroll
                                        for (int j = 1; j < 16; ++j) cm = fmaxf(cm, v[c * 16 + j]);
                                    }
                                    if (F8) top3_insert_key(&v[c * 16], fmaxf(cm, KEY_FLOOR), (uint32_t)(cid0 + c), snap, m1, m2, m3);
                                    else top3_insert(&v[c * 16], cm, cid0 + c, tau_c, m1, m2, m3, i1, i2, k1, k2);
                                }
                            }
                        }
                    }
                }
                if (F8) {                                       // keys -> (maximum, chunk id); the deferred member mask of this set's best chunk
                    key_unpack(m1, m1, i1);
                    key_unpack(m2, m2, i2);
                    if (!(m3 > KEY_FLOOR_TEST)) m3 = -CUDART_INF_F;
                    k1 = i1 >= 0 ? top3_snapshot_mask(snap, m1, tau_c) : 0u;
                    k2 = 0xffffu;
                }
                const int r_local = h * TILE_ROWS + q * 32 + lane;
                // the two sets' lists of a row meet through shared memory
                if (set == 1) {
                    merge[2 * r_local] = make_float4(m1, m2, m3, __int_as_float(i1));
                    merge[2 * r_local + 1] = make_float4(__int_as_float(i2), __uint_as_float(k1), __uint_as_float(k2), 0.f);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * N_EPI_WARPS) : "memory");
                if (set == 0) {
                    const float4 a = merge[2 * r_local], b = merge[2 * r_local + 1];
                    top3_merge(a.x, __float_as_int(a.w), __float_as_uint(b.y), m1, m2, m3, i1, i2, k1, k2);
                    top3_merge(a.y, __float_as_int(b.x), __float_as_uint(b.z), m1, m2, m3, i1, i2, k1, k2);
                    if (a.z > m3) m3 = a.z;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * N_EPI_WARPS) : "memory");
                const int64_t row = (int64_t)mt * MT_ROWS + r_local;
                if (set == 0 && row < p.n_emb) {
                    Cand c;
                    c.m1 = m1; c.m2 = m2; c.m3 = m3; c.i1 = i1; c.i2 = i2;
                    c.masks = (k1 & 0xffffu) | (k2 << 16); c.pad[0] = c.pad[1] = 0;
                    p.cand[row] = c;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------- operand packing

// One warp per row: lane j converts the 8-element chunk j of the padded row to fp16 and
// writes its 16 bytes into the tile image.  err[2r] = |x - fp16(x)|_2, err[2r+1] = |x|_2.
__global__ void pack_x_kernel(const float *X, int64_t n_emb, int64_t n_rows_pad, int D, int KP, uint8_t *tiles,
                              float *err, float *x_max) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n_rows_pad) return;
    const int64_t tile = row / TILE_ROWS;
    const int r = (int)(row % TILE_ROWS);
    uint8_t *base = tiles + tile * ((int64_t)TILE_ROWS * KP * 2);
    float e2 = 0.f, n2 = 0.f;
    bool overflow = false;
    for (int ch = lane; ch < KP / 8; ch += 32) {
        __align__(16) __half hv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = ch * 8 + j;
            float x = 0.f;
            if (row < n_emb) {
                if (c < D) x = X[row * D + c];
                else if (c < D + 3) x = 1.0f;          // multiplies the three -|mu|^2/2 pieces
            }
            if (fabsf(x) > 60000.f) overflow = true;
            const __half hx = __float2half_rn(x);
            hv[j] = hx;
            if (c < D) { const float dl = x - __half2float(hx); e2 += dl * dl; n2 += x * x; }
        }
        *reinterpret_cast<uint4 *>(base + tile_off(r, ch * 8)) = *reinterpret_cast<const uint4 *>(hv);
    }
    for (int o = 16; o > 0; o >>= 1) { e2 += __shfl_xor_sync(FULL, e2, o); n2 += __shfl_xor_sync(FULL, n2, o); }
    overflow = __any_sync(FULL, overflow);
    if (lane == 0 && row < n_emb) {
        const float ex = overflow ? CUDART_INF_F : sqrtf(e2) * 1.0001f, nx = sqrtf(n2) * 1.0001f;
        err[2 * row] = ex;
        err[2 * row + 1] = nx;
        // non-negative floats order like their bit patterns
        atomicMax(reinterpret_cast<int *>(x_max), __float_as_int(ex));
        atomicMax(reinterpret_cast<int *>(x_max) + 1, __float_as_int(nx));
    }
}

// model-wide maxima of the per-component rounding error and norm -> w_max[0..1]
__global__ void wmax_kernel(const float *w_err, int K_max, float *w_max) {
    __shared__ float red[64];
    float e_mu = 0.f, n_mu = 0.f;
    for (int k = threadIdx.x; k < K_max; k += blockDim.x) { e_mu = fmaxf(e_mu, w_err[2 * k]); n_mu = fmaxf(n_mu, w_err[2 * k + 1]); }
    for (int o = 16; o > 0; o >>= 1) { e_mu = fmaxf(e_mu, __shfl_xor_sync(FULL, e_mu, o)); n_mu = fmaxf(n_mu, __shfl_xor_sync(FULL, n_mu, o)); }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = e_mu; red[32 + (threadIdx.x >> 5)] = n_mu; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) { e_mu = fmaxf(e_mu, red[i]); n_mu = fmaxf(n_mu, red[32 + i]); }
        w_max[0] = e_mu; w_max[1] = n_mu;
    }
}

// Means tile image + per-component (|mu - fp16(mu)|, |fp16(mu)|); padded components get PAD_BIAS.
__global__ void pack_w_kernel(const float *means, int K_max, int K_pad, int D, int KP, uint8_t *tiles, float *err) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= K_pad) return;
    const int tile = row / TILE_ROWS, r = row % TILE_ROWS;
    uint8_t *base = tiles + (int64_t)tile * ((int64_t)TILE_ROWS * KP * 2);
    // |mu|^2 in float64 over the float32 means
    double nrm = 0.0;
    if (row < K_max)
        for (int d = lane; d < D; d += 32) { const double v = means[(int64_t)row * D + d]; nrm += v * v; }
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(FULL, nrm, o);
    float bias = (row < K_max) ? (float)(-0.5 * nrm) : PAD_BIAS;
    bool overflow = !(fabsf(bias) < 60000.f);
    const __half b0 = __float2half_rn(bias);
    const float r1 = bias - __half2float(b0);
    const __half b1 = __float2half_rn(r1);
    const __half b2 = __float2half_rn(r1 - __half2float(b1));
    float e2 = 0.f, n2 = 0.f;
    for (int ch = lane; ch < KP / 8; ch += 32) {
        __align__(16) __half hv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = ch * 8 + j;
            __half hx = __float2half_rn(0.f);
            if (c < D) {
                const float x = (row < K_max) ? means[(int64_t)row * D + c] : 0.f;
                if (fabsf(x) > 60000.f) overflow = true;
                hx = __float2half_rn(x);
                const float dl = x - __half2float(hx);
                e2 += dl * dl;
                n2 += __half2float(hx) * __half2float(hx);
            } else if (c == D) hx = b0;
            else if (c == D + 1) hx = b1;
            else if (c == D + 2) hx = b2;
            hv[j] = hx;
        }
        *reinterpret_cast<uint4 *>(base + tile_off(r, ch * 8)) = *reinterpret_cast<const uint4 *>(hv);
    }
    for (int o = 16; o > 0; o >>= 1) { e2 += __shfl_xor_sync(FULL, e2, o); n2 += __shfl_xor_sync(FULL, n2, o); }
    overflow = __any_sync(FULL, overflow);
    if (lane == 0) {
        err[2 * row] = (row < K_max) ? (overflow ? CUDART_INF_F : sqrtf(e2) * 1.0001f) : 0.f;
        err[2 * row + 1] = (row < K_max) ? sqrtf(n2) * 1.0001f : 0.f;
    }
}

}  // namespace mma
}  // namespace segb

// ---------------------------------------------------------------- exact refine (float32, NumPy order)

namespace segb {
namespace mma {

// Exact float32 score of component k for the row staged in xs (same routine as kmeans.cu).
// Cold path (runner-up chunks, exhaustive scans): kept out of line so it does not inflate the
// register footprint of the hot loops.
__device__ __noinline__ float exact_neg_dist(const float *meansT, int KM_, int k, const float *xs, int D) {
    auto f = [&](int d) -> float {
        const float dl = __fsub_rn(meansT[(size_t)d * KM_ + k], xs[d]);
        return __fmul_rn(dl, dl);
    };
    float s;
    if (D <= 128) s = pairwise_block<float>(f, 0, D);
    else if (D <= 256) {
        int n2 = D / 2; n2 -= n2 % 8;
        s = __fadd_rn(pairwise_block<float>(f, 0, n2), pairwise_block<float>(f, n2, D - n2));
    } else s = pairwise_sum<float>(f, D);
    return -s;
}

// ---- refine ---------------------------------------------------------------------------
// The filter leaves, per embedding, the best two chunks and a 16-bit mask of the members of
// each that lie within tau of the chunk maximum -- normally ONE component.  Those are re-scored
// in the reference's exact arithmetic by a half-warp per embedding: NumPy's pairwise sum keeps 8
// running accumulators per block of <= 128 terms, so lane q of an 8-lane group owns accumulator
// q (terms d = q, q+8, ...), the butterfly (xor 1, 2, 4) reproduces the combination tree
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) bit for bit, and the second 8-lane group handles the second
// block when D > 128.  Loads are coalesced along d (row-major X and means).

constexpr int REFINE_THREADS = 256;
constexpr int REFINE_MAX_STEPS = 16;     // 128 terms / 8 accumulators

__global__ void __launch_bounds__(REFINE_THREADS) refine_rows_kernel(
    segb_kmeans m, const Cand *cand, const float *x_err, const float *w_max, int64_t n_emb, int n_chunks,
    float *best_val, int32_t *best_k, unsigned long long *n_fallback, int32_t *fb_list) {
    const int D = m.D, KM = m.K_max;
    const int lane = threadIdx.x & 31, j = lane & 15, g = j >> 3, q = j & 7;
    const unsigned hmask = 0xffffu << (lane & 16);
    const int hbase = lane & 16;
    const float *X = (const float *)m.X;
    const float *means = (const float *)m.means;
    const float e_mu = w_max[0], n_mu = w_max[1];
    // block structure of NumPy's pairwise sum for n = D (n <= 128: one block; else split once)
    int n2 = 0;
    if (D > 128) { n2 = D / 2; n2 -= n2 % 8; }
    const int lo_g = g == 0 ? 0 : n2;                         // first term of this group's block
    const int n_g = (D > 128) ? (g == 0 ? n2 : D - n2) : (g == 0 ? D : 0);
    const int n8_g = n_g >= 8 ? n_g - (n_g % 8) : 0;          // terms covered by the 8 accumulators
    const int steps = n8_g / 8;
    const bool two_blocks = D > 128;

    const int64_t hw_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int64_t hw_total = ((int64_t)gridDim.x * blockDim.x) >> 4;
    for (int64_t row = hw_global; row < n_emb; row += hw_total) {
        const Cand c = cand[row];
        const float tau = filter_tau(x_err[2 * row], x_err[2 * row + 1], e_mu, n_mu, D);
        const int code = refine_decide(c, tau, n_chunks);
        if (code == -2) {
            if (j == 0) fb_list[atomicAdd(n_fallback, 1ull)] = (int32_t)row;      // -> refine_full_kernel
            continue;
        }
        const float *xr = X + row * D;
        float xv[REFINE_MAX_STEPS];
#pragma unroll
        for (int i = 0; i < REFINE_MAX_STEPS; ++i) xv[i] = (i < steps) ? xr[lo_g + i * 8 + q] : 0.f;
        float bv = -CUDART_INF_F;
        int bk = 0x7fffffff;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            uint32_t mk = pass == 0 ? (c.masks & 0xffffu) : (code >= 0 ? (c.masks >> 16) : 0u);
            const int chunk = pass == 0 ? c.i1 : c.i2;
            while (mk) {
                const int bit = __ffs(mk) - 1;
                mk &= mk - 1;
                const int k = chunk * CHUNK + bit;
                if (k >= KM) continue;
                const float *mu = means + (size_t)k * D;
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < REFINE_MAX_STEPS; ++i) {
                    if (i < steps) {
                        const float dl = __fsub_rn(mu[lo_g + i * 8 + q], xv[i]);
                        const float pr = __fmul_rn(dl, dl);
                        acc = (i == 0) ? pr : __fadd_rn(acc, pr);
                    }
                }
                // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) inside each 8-lane group
                acc = __fadd_rn(acc, __shfl_xor_sync(hmask, acc, 1));
                acc = __fadd_rn(acc, __shfl_xor_sync(hmask, acc, 2));
                acc = __fadd_rn(acc, __shfl_xor_sync(hmask, acc, 4));
                // the remaining n % 8 terms of the block are added one by one (lane q == 0 of the group)
                if (n8_g == 0) acc = 0.f;
                for (int d = lo_g + n8_g; d < lo_g + n_g; ++d) {
                    const float dl = __fsub_rn(mu[d], xr[d]);
                    acc = __fadd_rn(acc, __fmul_rn(dl, dl));
                }
                float tot = __shfl_sync(hmask, acc, hbase);
                if (two_blocks) tot = __fadd_rn(tot, __shfl_sync(hmask, acc, hbase + 8));
                const float v = -tot;
                if (v > bv || (v == bv && k < bk)) { bv = v; bk = k; }
            }
        }
        if (j == 0) { best_val[row] = bv; best_k[row] = (bk == 0x7fffffff) ? -1 : bk; }
    }
}

// Same refine with EIGHT lanes per embedding (four embeddings per warp): lane (g, c) owns NumPy's
// accumulators 2c and 2c+1 of block g and moves them with 8-byte loads, so the per-row bookkeeping
// (filter record, bound, mask walk) is shared by 8 lanes instead of 16 and every load instruction
// carries two terms.  Needs an even D (8-byte aligned rows); the combination tree is unchanged:
// (r[2c] + r[2c+1]) locally, then xor 1 and xor 2 inside the block's four lanes.
// fp8: the records come from the e4m3 filter pass: x_err / w_max hold the e4m3 rounding-error norms of the scaled
// operands (w_max = (e_mu, n_mu, e_bias, bias_max)) and the threshold is filter_tau8's.
// DC: the embedding dimension when known at compile time (0 = runtime): the per-lane step counts, load offsets and the
// n % 8 tail become constants, which removes the predicates and address arithmetic of the generic unrolled loops.
template <int MAXS, int DC = 0>
__global__ void __launch_bounds__(REFINE_THREADS, 4) refine_rows8_kernel(
    segb_kmeans m, const Cand *cand, const float *x_err, const float *w_max, int64_t n_emb, int n_chunks,
    float *best_val, int32_t *best_k, unsigned long long *n_fallback, int32_t *fb_list, int fp8 = 0) {
    const int D = DC ? DC : m.D, KM = m.K_max;
    const int lane = threadIdx.x & 31, j = lane & 7;
    const Row8Geom geo(D, lane);
    const float *X = (const float *)m.X;
    const float *means = (const float *)m.means;
    const float e_mu = w_max[0], n_mu = w_max[1];
    const int64_t grp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int64_t grp_total = ((int64_t)gridDim.x * blockDim.x) >> 3;
    // The row's dependent chain is filter record -> mean row -> arithmetic; its embedding row does not
    // depend on the record, and neither does the NEXT row's record: both are requested before the current
    // record is examined, so three memory round trips overlap instead of following one another.
    Cand cd_next;
    float2 xe_next = make_float2(0.f, 0.f);
    if (grp_global < n_emb) { cd_next = cand[grp_global]; xe_next = *reinterpret_cast<const float2 *>(x_err + 2 * grp_global); }
    for (int64_t row = grp_global; row < n_emb; row += grp_total) {
        const Cand cd = cd_next;
        const float2 xe = xe_next;
        float2 xv[MAXS];
        km_load_x8<MAXS>(X + row * D, geo, xv);
        float2 xt[ROW8_TAIL];
        km_load_xtail8(X + row * D, geo, xt);
        const int64_t row_n = row + grp_total;
        if (row_n < n_emb) { cd_next = cand[row_n]; xe_next = *reinterpret_cast<const float2 *>(x_err + 2 * row_n); }
        const float tau = fp8 ? filter_tau8(xe.x, xe.y, e_mu, n_mu, w_max[2], w_max[3], D) : filter_tau(xe.x, xe.y, e_mu, n_mu, D);
        const int code = refine_decide(cd, tau, n_chunks);
        if (code == -2) {
            if (j == 0) fb_list[atomicAdd(n_fallback, 1ull)] = (int32_t)row;      // -> refine_full_kernel
            continue;
        }
        float bv;
        int bk;
        km_exact_row8<MAXS>(means, KM, D, X + row * D, xv, cd.i1, cd.i2, cd.masks, code, geo, bv, bk, xt);
        if (j == 0) { best_val[row] = bv; best_k[row] = (bk == 0x7fffffff) ? -1 : bk; }
    }
}


// ---- second-level pass over the rows the top-3 records could not decide ----------------------------------
// A model with many near-duplicate components (a diffuse k-means state: K_true << K_max) leaves more than
// three 16-component chunks inside the error bound of the best score, and an exhaustive exact scan of such a
// row costs K_max * D * 3 float operations.  Instead the undecided rows are gathered into a compact fp16 tile
// image, the SAME filter GEMM runs over them once more with a different epilogue -- a bitmap of every component
// whose filter score reaches the row's threshold (best score - tau, both known from the first pass) -- and only
// those components (tens, not thousands) are re-scored exactly.  Rigour is unchanged: the reference's winner k*
// has t^(k*) >= t(k*) - b >= t(k1) - b >= t^(k1) - 2b = m1 - tau for ANY filter scores within b of the truth.

// one warp per undecided row i: fp16 row of the tile image -> compact tile image, thr[i] = m1 - tau(row);
// the rows that pad the last 256-row work item are zeroed and get thr = +inf (empty bitmap)
__global__ void __launch_bounds__(256) gather_undecided_kernel(const uint8_t *x_tiles, const Cand *cand, const float *x_err,
                                                               const float *w_max, const int32_t *fb_list,
                                                               const unsigned long long *n_fallback, int64_t first, int64_t rows_cap,
                                                               int D, int KP, uint8_t *fb_tiles, float *thr) {
    const int lane = threadIdx.x & 31;
    const int64_t left = (int64_t)*n_fallback - first;            // this round: rows [first, first + rows_cap) of the list
    const int64_t n = left < 0 ? 0 : (left < rows_cap ? left : rows_cap);
    fb_list += first;
    const int64_t n_pad = (n + MT_ROWS - 1) / MT_ROWS * MT_ROWS;
    const int64_t tile_b = (int64_t)TILE_ROWS * KP * 2;
    const int64_t w_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t w_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = w_global; i < n_pad; i += w_total) {
        uint8_t *dst = fb_tiles + (i / TILE_ROWS) * tile_b;
        const int ri = (int)(i % TILE_ROWS);
        if (i < n) {
            const int64_t row = fb_list[i];
            const uint8_t *src = x_tiles + (row / TILE_ROWS) * tile_b;
            const int rr = (int)(row % TILE_ROWS);
            for (int ch = lane; ch < KP / 8; ch += 32)
                *reinterpret_cast<uint4 *>(dst + tile_off(ri, ch * 8)) = *reinterpret_cast<const uint4 *>(src + tile_off(rr, ch * 8));
            if (lane == 0) {
                const Cand c = cand[row];
                const float tau = filter_tau(x_err[2 * row], x_err[2 * row + 1], w_max[0], w_max[1], D);
                // unusable record or unbounded error: every component stays a candidate
                thr[i] = (c.i1 >= 0 && tau < CUDART_INF_F && c.m1 > -CUDART_INF_F) ? c.m1 - tau : -CUDART_INF_F;
            }
        } else {
            for (int ch = lane; ch < KP / 8; ch += 32)
                *reinterpret_cast<uint4 *>(dst + tile_off(ri, ch * 8)) = make_uint4(0, 0, 0, 0);
            if (lane == 0) thr[i] = CUDART_INF_F;
        }
    }
}

// eight lanes per undecided row: exact score (same routine and bits as refine_rows8_kernel) of every component
// flagged in the row's bitmap, first maximum in component order.  A row with an empty bitmap (cannot happen for
// finite data; NaN scores) goes to the exhaustive scan through unres_list.
template <int MAXS, int DC = 0>           // DC: the dimension when known at compile time (see refine_rows8_kernel)
__global__ void __launch_bounds__(REFINE_THREADS) refine_bitmap_kernel(
    segb_kmeans m, const int32_t *fb_list, const unsigned long long *n_fallback, int64_t first, int64_t rows_cap,
    const uint32_t *bitmap, int n_words, float *best_val, int32_t *best_k, unsigned long long *n_unres, int32_t *unres_list) {
    const int D = DC ? DC : m.D, KM = m.K_max;
    const int lane = threadIdx.x & 31, j = lane & 7;
    const Row8Geom geo(D, lane);
    const float *X = (const float *)m.X;
    const float *means = (const float *)m.means;
    const int64_t left = (int64_t)*n_fallback - first;
    const int64_t n = left < 0 ? 0 : (left < rows_cap ? left : rows_cap);
    fb_list += first;
    const int64_t grp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int64_t grp_total = ((int64_t)gridDim.x * blockDim.x) >> 3;
    int64_t row_next = grp_global < n ? fb_list[grp_global] : 0;
    for (int64_t i = grp_global; i < n; i += grp_total) {
        const int64_t row = row_next;
        float2 xv[MAXS];
        km_load_x8<MAXS>(X + row * D, geo, xv);
        float2 xt[ROW8_TAIL];
        km_load_xtail8(X + row * D, geo, xt);
        if (i + grp_total < n) row_next = fb_list[i + grp_total];      // consumed (and its row prefetched) at the end of this iteration
        const uint32_t *bm = bitmap + i * n_words;
        float bv = -CUDART_INF_F;
        int bk = 0x7fffffff;
        for (int w0 = 0; w0 < n_words; w0 += 8) {
            // lane j of the group fetches word w0 + j; the eight words are then walked by the whole group
            const uint32_t mine = (w0 + j < n_words) ? bm[w0 + j] : 0u;
            if (__ballot_sync(geo.gmask, mine != 0u) == 0u) continue;
            for (int t = 0; t < 8; ++t) {
                uint32_t bits = __shfl_sync(geo.gmask, mine, geo.gbase + t);
                while (bits) {
                    const int k = (w0 + t) * 32 + (__ffs(bits) - 1);
                    bits &= bits - 1;
                    if (k >= KM) continue;
                    const float v = km_exact_one8<MAXS>(means, D, X + row * D, xv, k, geo, xt);
                    if (v > bv || bk == 0x7fffffff) { bv = v; bk = k; }          // k ascending: first maximum
                }
            }
        }
        if (j == 0) {
            if (bk == 0x7fffffff) unres_list[atomicAdd(n_unres, 1ull)] = (int32_t)row;
            else { best_val[row] = bv; best_k[row] = bk; }
        }
        // the list sends the groups to scattered rows: the next one's embedding is pulled into L2 ahead of its loads
        if (i + grp_total < n) prefetch_row_l2(X + row_next * D, D, j);
    }
}


// ---- e4m3 first-level filter: operand packing ----------------------------------------------------------------
// The e4m3 pass (kind::f8f6f4: twice the MMA rate and half the operand bytes of the fp16 pass) decides the rows
// whose best component is well separated -- a trained model -- with the same rigorous-bound logic; what it cannot
// decide goes to the fp16 second-level pass.  Operands are scaled by a common power of two `scale` (exact) into
// e4m3's normal range; three constant columns (x side 256.0) carry a three-term e4m3 representation of
// -scale^2 |mu|^2 / 2 / 256.  Error norms are the ACTUAL rounding errors of the scaled operands.
__device__ __forceinline__ uint8_t to_e4m3(float v) { return (uint8_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3); }
__device__ __forceinline__ float from_e4m3(uint8_t b) { return __half2float(__half(__nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)b, __NV_E4M3))); }
constexpr float BIAS_COL8 = 256.0f;

// one warp per row; lane ch packs the 16-element chunk ch.  err[2r] = |s x - e4m3(s x)|_2, err[2r+1] = |s x|_2
__global__ void pack_x8_kernel(const float *X, int64_t n_emb, int64_t n_rows_pad, int D, int KP, float scale, uint8_t *tiles,
                               float *err, float *x_max) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n_rows_pad) return;
    uint8_t *base = tiles + (row / TILE_ROWS) * ((int64_t)TILE_ROWS * KP);
    const int r = (int)(row % TILE_ROWS);
    float e2 = 0.f, n2 = 0.f;
    for (int ch = lane; ch < KP / 16; ch += 32) {
        __align__(16) uint8_t hv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int c = ch * 16 + j;
            float x = 0.f;
            if (row < n_emb) {
                if (c < D) x = X[row * D + c] * scale;
                else if (c < D + 3) x = BIAS_COL8;
            }
            const uint8_t q = to_e4m3(x);
            hv[j] = q;
            if (c < D) { const float dl = x - from_e4m3(q); e2 += dl * dl; n2 += x * x; }
        }
        *reinterpret_cast<uint4 *>(base + tile_off8(r, ch * 16)) = *reinterpret_cast<const uint4 *>(hv);
    }
    for (int o = 16; o > 0; o >>= 1) { e2 += __shfl_xor_sync(FULL, e2, o); n2 += __shfl_xor_sync(FULL, n2, o); }
    if (lane == 0 && row < n_emb) {
        float ex = sqrtf(e2) * 1.0001f, nx = sqrtf(n2) * 1.0001f;
        if (!(ex < CUDART_INF_F) || !(nx < CUDART_INF_F)) ex = nx = CUDART_INF_F;      // NaN / inf rows: never decided here
        err[2 * row] = ex;
        err[2 * row + 1] = nx;
        atomicMax(reinterpret_cast<int *>(x_max), __float_as_int(ex));
        atomicMax(reinterpret_cast<int *>(x_max) + 1, __float_as_int(nx));
    }
}

// means image; err[4k..] = (|s mu - e4m3(s mu)|, |e4m3(s mu)|, |bias - represented bias|, |bias|); padded rows are zeros
// (the filter's epilogue masks them).
__global__ void pack_w8_kernel(const float *means, int K_max, int K_pad, int D, int KP, float scale, uint8_t *tiles, float *err) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= K_pad) return;
    uint8_t *base = tiles + (int64_t)(row / TILE_ROWS) * ((int64_t)TILE_ROWS * KP);
    const int r = row % TILE_ROWS;
    double nrm = 0.0;
    if (row < K_max)
        for (int d = lane; d < D; d += 32) { const double v = (double)means[(int64_t)row * D + d] * scale; nrm += v * v; }
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(FULL, nrm, o);
    const float bias = (row < K_max) ? (float)(-0.5 * nrm) : 0.f;                  // scaled space
    const float t0 = bias / BIAS_COL8;
    const uint8_t b0 = to_e4m3(t0);
    const float r1 = t0 - from_e4m3(b0);
    const uint8_t b1 = to_e4m3(r1);
    const float r2 = r1 - from_e4m3(b1);
    const uint8_t b2 = to_e4m3(r2);
    const float e_bias = fabsf(r2 - from_e4m3(b2)) * BIAS_COL8 * 1.0001f + fabsf(bias) * 2e-7f;   // + float32 rounding of the split
    float e2 = 0.f, n2 = 0.f;
    for (int ch = lane; ch < KP / 16; ch += 32) {
        __align__(16) uint8_t hv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int c = ch * 16 + j;
            uint8_t q = 0;
            if (c < D) {
                const float x = (row < K_max) ? means[(int64_t)row * D + c] * scale : 0.f;
                q = to_e4m3(x);
                const float dq = from_e4m3(q), dl = x - dq;
                e2 += dl * dl;
                n2 += dq * dq;
            } else if (row < K_max) {
                if (c == D) q = b0;
                else if (c == D + 1) q = b1;
                else if (c == D + 2) q = b2;
            }
            hv[j] = q;
        }
        *reinterpret_cast<uint4 *>(base + tile_off8(r, ch * 16)) = *reinterpret_cast<const uint4 *>(hv);
    }
    for (int o = 16; o > 0; o >>= 1) { e2 += __shfl_xor_sync(FULL, e2, o); n2 += __shfl_xor_sync(FULL, n2, o); }
    if (lane == 0) {
        const bool live = row < K_max;
        float em = sqrtf(e2) * 1.0001f, nm = sqrtf(n2) * 1.0001f;
        if (!(em < CUDART_INF_F) || !(nm < CUDART_INF_F) || !(e_bias < CUDART_INF_F)) em = CUDART_INF_F;
        err[4 * row] = live ? em : 0.f;
        err[4 * row + 1] = live ? nm : 0.f;
        err[4 * row + 2] = live ? e_bias : 0.f;
        err[4 * row + 3] = live ? fabsf(bias) : 0.f;
    }
}

// model-wide maxima of the four per-component quantities -> w_max[0..3]
__global__ void wmax4_kernel(const float *w_err, int K_max, float *w_max) {
    __shared__ float red[4][32];
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = threadIdx.x; k < K_max; k += blockDim.x)
        for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], w_err[4 * k + i]);
    for (int i = 0; i < 4; ++i) {
        for (int o = 16; o > 0; o >>= 1) v[i] = fmaxf(v[i], __shfl_xor_sync(FULL, v[i], o));
        if ((threadIdx.x & 31) == 0) red[i][threadIdx.x >> 5] = v[i];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float r = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, red[threadIdx.x][w]);
        w_max[threadIdx.x] = r;
    }
}

// Second level after the e4m3 pass: no resident fp16 image of X exists, so the undecided rows are converted from
// the fp32 embeddings on the fly (one warp per row, as pack_x_kernel) into the compact fp16 tile image, and the
// row's threshold for the fp16 bitmap pass follows from the e4m3 record: the reference's winner k* has
// t(k*) >= t(k1) >= m1/s^2 - b8, hence t16^(k*) >= m1/s^2 - b8 - b16 (b8, b16: the two passes' error bounds).
__global__ void __launch_bounds__(256) gather_convert_undecided_kernel(
    const float *X, const Cand *cand, const float *x_err8, const float *w_max8, float scale, const float *w_max16,
    const int32_t *fb_list, const unsigned long long *n_fallback, int64_t first, int64_t rows_cap, int D, int KP, uint8_t *fb_tiles,
    float *thr) {
    const int lane = threadIdx.x & 31;
    const int64_t left = (int64_t)*n_fallback - first;
    const int64_t n = left < 0 ? 0 : (left < rows_cap ? left : rows_cap);
    fb_list += first;
    const int64_t n_pad = (n + MT_ROWS - 1) / MT_ROWS * MT_ROWS;
    const int64_t tile_b = (int64_t)TILE_ROWS * KP * 2;
    const int64_t w_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t w_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float inv_s2 = 1.0f / (scale * scale);
    for (int64_t i = w_global; i < n_pad; i += w_total) {
        uint8_t *dst = fb_tiles + (i / TILE_ROWS) * tile_b;
        const int ri = (int)(i % TILE_ROWS);
        const int64_t row = i < n ? (int64_t)fb_list[i] : -1;
        float e2 = 0.f, n2 = 0.f;
        bool overflow = false;
        for (int ch = lane; ch < KP / 8; ch += 32) {
            __align__(16) __half hv[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int c = ch * 8 + jj;
                float x = 0.f;
                if (row >= 0) {
                    if (c < D) x = X[row * D + c];
                    else if (c < D + 3) x = 1.0f;
                }
                if (fabsf(x) > 60000.f) overflow = true;
                const __half hx = __float2half_rn(x);
                hv[jj] = hx;
                if (c < D) { const float dl = x - __half2float(hx); e2 += dl * dl; n2 += x * x; }
            }
            *reinterpret_cast<uint4 *>(dst + tile_off(ri, ch * 8)) = *reinterpret_cast<const uint4 *>(hv);
        }
        for (int o = 16; o > 0; o >>= 1) { e2 += __shfl_xor_sync(FULL, e2, o); n2 += __shfl_xor_sync(FULL, n2, o); }
        overflow = __any_sync(FULL, overflow);
        if (lane == 0) {
            float t = CUDART_INF_F;                              // padding rows: empty bitmap
            if (row >= 0) {
                const Cand c = cand[row];
                const float ex16 = overflow ? CUDART_INF_F : sqrtf(e2) * 1.0001f, nx = sqrtf(n2) * 1.0001f;
                const float tau16 = filter_tau(ex16, nx, w_max16[0], w_max16[1], D);
                const float tau8 = filter_tau8(x_err8[2 * row], x_err8[2 * row + 1], w_max8[0], w_max8[1], w_max8[2], w_max8[3], D);
                const float lo = c.m1 * inv_s2 - 0.5f * tau8 * inv_s2 - 0.5f * tau16;
                // one float32 rounding in each of the three terms: widen by 2^-21 of their magnitudes
                const float slack = ldexpf(fabsf(c.m1 * inv_s2) + tau8 * inv_s2 + tau16, -21);
                t = (c.i1 >= 0 && tau8 < CUDART_INF_F && tau16 < CUDART_INF_F && c.m1 > -CUDART_INF_F) ? lo - slack : -CUDART_INF_F;
            }
            thr[i] = t;
        }
    }
}

// Exhaustive exact scan for the rows the filter could not decide, FULLB_R rows per block pass: the rows sit
// in shared memory ([d][r], one 16-byte read serves four rows), every thread walks its components through the
// transposed means (coalesced over k) and keeps, for each of the rows, NumPy's eight running accumulators of
// the current pairwise block in registers -- the means are streamed once per FULLB_R rows instead of once per
// row (the one-row-per-block version below read 2.6 MB of L2 per undecided row: 51 ms for 230k rows of a
// diffuse model).  Same separately rounded operations and combination tree as exact_neg_dist: bit-exact.
// D <= 256 (at most one split of NumPy's pairwise sum).
constexpr int FULLB_R = 8, FULLB_THREADS = 256;
__global__ void __launch_bounds__(FULLB_THREADS) refine_full_blocked_kernel(segb_kmeans m, const int32_t *fb_list,
                                                                            const unsigned long long *n_fallback,
                                                                            long long first, float *best_val, int32_t *best_k) {
    extern __shared__ float fsm[];
    float *xs = fsm;                                   // [D][FULLB_R]
    float *rv = fsm + (size_t)m.D * FULLB_R;           // [FULLB_R][8] per-warp best value
    int *rk = (int *)(rv + FULLB_R * 8);               // [FULLB_R][8] per-warp best index
    const int D = m.D, KM = m.K_max;
    const float *X = (const float *)m.X;
    const float *meansT = (const float *)m.meansT;
    const long long n = (long long)*n_fallback;          // rows [first, n) of the list
    int n2 = 0;
    if (D > 128) { n2 = D / 2; n2 -= n2 % 8; }
    const int n_blocks = D > 128 ? 2 : 1;
    for (long long base = first + (long long)blockIdx.x * FULLB_R; base < n; base += (long long)gridDim.x * FULLB_R) {
        const int nr = (int)min((long long)FULLB_R, n - base);
        __syncthreads();
        for (int i = threadIdx.x; i < D * FULLB_R; i += blockDim.x) {
            const int r = i / D, d = i % D;
            xs[d * FULLB_R + r] = r < nr ? X[(int64_t)fb_list[base + r] * D + d] : 0.f;
        }
        __syncthreads();
        float bv[FULLB_R];
        int bk[FULLB_R];
#pragma unroll
        for (int r = 0; r < FULLB_R; ++r) { bv[r] = -CUDART_INF_F; bk[r] = 0x7fffffff; }
        for (int k = threadIdx.x; k < KM; k += blockDim.x) {
            float tot[FULLB_R];
#pragma unroll
            for (int r = 0; r < FULLB_R; ++r) tot[r] = 0.f;
            for (int b = 0; b < n_blocks; ++b) {
                const int lo = b == 0 ? 0 : n2, nb = D > 128 ? (b == 0 ? n2 : D - n2) : D;
                const int n8 = nb >= 8 ? nb - (nb % 8) : 0;
                float res[FULLB_R];
                if (n8 > 0) {
                    float acc[FULLB_R][8];
                    for (int i = 0; i < n8; i += 8) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int d = lo + i + j;
                            const float mu = meansT[(size_t)d * KM + k];
                            const float4 xa = *reinterpret_cast<const float4 *>(xs + d * FULLB_R);
                            const float4 xb = *reinterpret_cast<const float4 *>(xs + d * FULLB_R + 4);
                            const float xr[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
                            for (int r = 0; r < FULLB_R; ++r) {
                                const float dl = __fsub_rn(mu, xr[r]);
                                const float pr = __fmul_rn(dl, dl);
                                acc[r][j] = (i == 0) ? pr : __fadd_rn(acc[r][j], pr);
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < FULLB_R; ++r)
                        res[r] = __fadd_rn(__fadd_rn(__fadd_rn(acc[r][0], acc[r][1]), __fadd_rn(acc[r][2], acc[r][3])),
                                           __fadd_rn(__fadd_rn(acc[r][4], acc[r][5]), __fadd_rn(acc[r][6], acc[r][7])));
                } else {
#pragma unroll
                    for (int r = 0; r < FULLB_R; ++r) res[r] = 0.f;
                }
                for (int d = lo + n8; d < lo + nb; ++d) {                 // the block's n % 8 trailing terms
                    const float mu = meansT[(size_t)d * KM + k];
#pragma unroll
                    for (int r = 0; r < FULLB_R; ++r) {
                        const float dl = __fsub_rn(mu, xs[d * FULLB_R + r]);
                        res[r] = __fadd_rn(res[r], __fmul_rn(dl, dl));
                    }
                }
#pragma unroll
                for (int r = 0; r < FULLB_R; ++r) tot[r] = b == 0 ? res[r] : __fadd_rn(tot[r], res[r]);
            }
#pragma unroll
            for (int r = 0; r < FULLB_R; ++r) {
                const float v = -tot[r];
                if (v > bv[r] || bk[r] == 0x7fffffff) { bv[r] = v; bk[r] = k; }       // k ascending per thread: first max
            }
        }
#pragma unroll
        for (int r = 0; r < FULLB_R; ++r) {
            float v = bv[r];
            int kk = bk[r];
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(FULL, v, o);
                const int ok = __shfl_xor_sync(FULL, kk, o);
                if (ov > v || (ov == v && ok < kk)) { v = ov; kk = ok; }
            }
            if ((threadIdx.x & 31) == 0) { rv[r * 8 + (threadIdx.x >> 5)] = v; rk[r * 8 + (threadIdx.x >> 5)] = kk; }
        }
        __syncthreads();
        if (threadIdx.x < nr) {
            const int r = threadIdx.x;
            float v = rv[r * 8];
            int kk = rk[r * 8];
            for (int w = 1; w < FULLB_THREADS / 32; ++w)
                if (rv[r * 8 + w] > v || (rv[r * 8 + w] == v && rk[r * 8 + w] < kk)) { v = rv[r * 8 + w]; kk = rk[r * 8 + w]; }
            const int64_t row = fb_list[base + r];
            best_val[row] = v;
            best_k[row] = (kk == 0x7fffffff) ? -1 : kk;
        }
    }
}

// Exhaustive exact scan for the rows the filter could not decide: one block per row.
__global__ void __launch_bounds__(256) refine_full_kernel(segb_kmeans m, const int32_t *fb_list,
                                                          const unsigned long long *n_fallback, long long first,
                                                          float *best_val, int32_t *best_k) {
    extern __shared__ float fsm[];
    float *xs = fsm;                      // [D]
    float *rv = fsm + m.D;                // [8] per-warp best value
    int *rk = (int *)(rv + 8);            // [8] per-warp best index
    const int D = m.D, KM = m.K_max;
    const float *X = (const float *)m.X;
    const float *meansT = (const float *)m.meansT;
    const long long n = (long long)*n_fallback;
    for (long long i = first + blockIdx.x; i < n; i += gridDim.x) {
        const int64_t row = fb_list[i];
        __syncthreads();
        for (int d = threadIdx.x; d < D; d += blockDim.x) xs[d] = X[row * D + d];
        __syncthreads();
        float bv = -CUDART_INF_F;
        int bk = 0x7fffffff;
        for (int k = threadIdx.x; k < KM; k += blockDim.x) {
            const float v = exact_neg_dist(meansT, KM, k, xs, D);
            if (v > bv || bk == 0x7fffffff) { bv = v; bk = k; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(FULL, bv, o);
            const int ok = __shfl_xor_sync(FULL, bk, o);
            if (ov > bv || (ov == bv && ok < bk)) { bv = ov; bk = ok; }
        }
        if ((threadIdx.x & 31) == 0) { rv[threadIdx.x >> 5] = bv; rk[threadIdx.x >> 5] = bk; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
                if (rv[w] > bv || (rv[w] == bv && rk[w] < bk)) { bv = rv[w]; bk = rk[w]; }
            best_val[row] = bv;
            best_k[row] = (bk == 0x7fffffff) ? -1 : bk;
        }
    }
}

static inline int64_t rows_pad(int64_t n) { return (n + MT_ROWS - 1) / MT_ROWS * MT_ROWS; }
static inline int k_pad(int K) { return (K + NT_COLS - 1) / NT_COLS * NT_COLS; }

}  // namespace mma
}  // namespace segb

using namespace segb;
using namespace segb::mma;

extern "C" int64_t segb_mma_x_tiles_bytes(int64_t n_emb, int32_t D) { return rows_pad(n_emb) * kp_of(D) * 2; }
extern "C" int64_t segb_mma_w_tiles_bytes(int32_t K_max, int32_t D) { return (int64_t)k_pad(K_max) * kp_of(D) * 2; }
extern "C" int64_t segb_mma_cand_bytes(int64_t n_emb) { return n_emb * (int64_t)sizeof(Cand); }

extern "C" int segb_mma_pack_x(const float *X, int64_t n_emb, int32_t D, void *x_tiles, float *x_err, float *x_max,
                               void *stream) {
    SEGB_CHECK_ARG(X && x_tiles && x_err && x_max && n_emb > 0 && D > 0, "null pointer");
    const int64_t np_ = rows_pad(n_emb);
    const int wpb = 8;
    SEGB_CUDA(cudaMemsetAsync(x_max, 0, 2 * sizeof(float), (cudaStream_t)stream));
    pack_x_kernel<<<(unsigned)((np_ + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        X, n_emb, np_, D, kp_of(D), (uint8_t *)x_tiles, x_err, x_max);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_mma_pack_means(const float *means, int32_t K_max, int32_t D, void *w_tiles, float *w_err,
                                   float *w_max, void *stream) {
    SEGB_CHECK_ARG(means && w_tiles && w_err && w_max && K_max > 0 && D > 0, "null pointer");
    const int kp_rows = k_pad(K_max);
    const int wpb = 8;
    pack_w_kernel<<<(kp_rows + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(
        means, K_max, kp_rows, D, kp_of(D), (uint8_t *)w_tiles, w_err);
    SEGB_LAUNCH_CHECK();
    wmax_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(w_err, K_max, w_max);
    SEGB_LAUNCH_CHECK();
    return 0;
}

namespace segb {
namespace mma {
static int launch_filter_impl(const FilterLaunch &f, const unsigned long long *n_rows_dev, int64_t rows_first, int64_t rows_cap,
                              const float *thr, uint32_t *bitmap, cudaStream_t stream) {
    FilterParams p;
    p.n_rows_dev = n_rows_dev; p.rows_first = rows_first; p.rows_cap = rows_cap; p.thr = thr; p.bitmap = bitmap;
    p.x_tiles = (const uint8_t *)f.x_tiles; p.w_tiles = (const uint8_t *)f.w_tiles; p.cand = (Cand *)f.cand;
    p.n_emb = f.n_emb;
    p.x_max = f.x_max; p.w_max = f.w_max; p.D = f.D;
    p.tau_kind = f.tau_kind; p.tau_T = f.tau_T; p.sx = f.sx; p.alpha = f.alpha;
    p.n_mtiles = (int32_t)(rows_pad(f.n_emb) / MT_ROWS);
    p.n_ntiles = f.w_rows_pad / NT_COLS;
    {
        int last = f.w_rows > 0 ? f.w_rows - (p.n_ntiles - 1) * NT_COLS : NT_COLS;
        last = (last + 15) / 16 * 16;
        p.n_last_cols = last < 16 ? 16 : (last > NT_COLS ? NT_COLS : last);
        p.n_chunks_valid = (p.n_ntiles - 1) * (NT_COLS / CHUNK) + p.n_last_cols / CHUNK;
        const int rows = f.w_rows > 0 ? f.w_rows : f.w_rows_pad;
        p.n_valid_last = rows - (p.n_chunks_valid - 1) * CHUNK;
        if (p.n_valid_last > CHUNK) p.n_valid_last = CHUNK;
    }
    p.n_ksteps = f.fp8 ? f.KP / 32 : f.KP / 16;                          // 32 bytes of K per row and instruction either way
    p.tile_bytes = (uint32_t)((int64_t)TILE_ROWS * f.KP * (f.fp8 ? 1 : 2));
    const int nch = f.n_chunks;
    if (nch != 1 && nch != 2) { set_error("filter GEMM: 1 or 2 inner-dimension chunks"); return SEGB_E_ARG; }
    const size_t tb = p.tile_bytes;
    const size_t fixed0 = 256 + (size_t)MT_ROWS * 32 + (f.fp8 ? (size_t)4 * SNAP_STRIDE * 16 : 0);   // barriers, merge buffer, F8: parked best chunks
    size_t fixed = 2 * tb + fixed0;                                     // + B stages
    const size_t budget = 227 * 1024;
    if (2 * nch * tb + fixed > budget) {
        set_error("inner dimension %d x %d too large for the tensor-core scorer", nch, f.KP);
        return SEGB_E_UNSUPPORTED;
    }
    p.n_abuf = (4 * nch * tb + fixed <= budget) ? 2 : 1;
    p.n_bstage = 2;
    if (nch == 1)           // deepen the B ring while double-buffered A still fits (e4m3: 4 x 20 KB A + 4..8 x 20 KB B)
        while (p.n_bstage < 8 && p.n_abuf == 2 && 4 * tb + 2 * p.n_bstage * tb + fixed0 <= budget) p.n_bstage *= 2;
    if (const char *env = getenv("SEGB_FILTER_BSTAGES")) { const int v = atoi(env); if (nch == 1 && (v == 2 || v == 4 || v == 8) && (size_t)p.n_abuf * 2 * tb + v * tb + fixed0 <= budget) p.n_bstage = v; }
    fixed = (size_t)p.n_bstage * tb + fixed0;
    const size_t smem = (size_t)p.n_abuf * 2 * nch * tb + fixed;
    int n_sm = 0;
    { const int rc = device_info(nullptr, &n_sm, nullptr); if (rc) return rc; }
    const int grid = p.n_mtiles < n_sm ? p.n_mtiles : n_sm;
    void (*kern)(FilterParams);
    if (f.fp8) {
        if (nch != 1 || n_rows_dev || f.w_rows <= 0) { set_error("fp8 filter pass: one chunk, top-3 epilogue, known component count"); return SEGB_E_ARG; }
        if (p.n_ntiles * (NT_COLS / CHUNK) > (1 << KEY_ID_BITS)) { set_error("fp8 filter pass: K_max <= 65536 (chunk ids ride in 12 mantissa bits of the packed top-3 keys)"); return SEGB_E_UNSUPPORTED; }
        kern = p.n_ksteps == 5 ? kmeans_filter_kernel<5, 1, 0, true> : kmeans_filter_kernel<0, 1, 0, true>;   // 5: D = 130 (KP = 160)
    } else if (n_rows_dev) {
        if (nch != 1) { set_error("second-level filter pass: one inner-dimension chunk"); return SEGB_E_ARG; }
        kern = p.n_ksteps == 9 ? kmeans_filter_kernel<9, 1, 1> : kmeans_filter_kernel<0, 1, 1>;
    } else if (nch == 1) kern = p.n_ksteps == 9 ? kmeans_filter_kernel<9, 1> : kmeans_filter_kernel<0, 1>;     // 9: D = 130 (KP = 144)
    else kern = p.n_ksteps == 9 ? kmeans_filter_kernel<9, 2> : kmeans_filter_kernel<0, 2>;
    SEGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, N_THREADS, smem, stream>>>(p);
    SEGB_LAUNCH_CHECK();
    return 0;
}
int launch_filter(const FilterLaunch &f, cudaStream_t stream) { return launch_filter_impl(f, nullptr, 0, 0, nullptr, nullptr, stream); }
}  // namespace mma
}  // namespace segb

extern "C" int segb_mma_filter(const void *x_tiles, const void *w_tiles, int64_t n_emb, int32_t K_max, int32_t D,
                               const float *x_max, const float *w_max, void *cand, void *stream) {
    SEGB_CHECK_ARG(x_tiles && w_tiles && cand && x_max && w_max && n_emb > 0 && K_max > 0, "null pointer");
    FilterLaunch f;
    f.x_tiles = x_tiles; f.w_tiles = w_tiles; f.cand = cand; f.n_emb = n_emb;
    f.w_rows_pad = k_pad(K_max); f.w_rows = K_max; f.KP = kp_of(D); f.D = D; f.x_max = x_max; f.w_max = w_max;
    f.n_chunks = 1; f.tau_kind = TAU_KMEANS; f.tau_T = 0.f;
    return launch_filter(f, (cudaStream_t)stream);
}

namespace segb {
namespace mma {
// exhaustive exact scan of the rows listed in fb_list[0 .. *n_fallback)
static int launch_refine_full_from(const segb_kmeans *m, const int32_t *fb_list, const int64_t *n_fallback, int64_t first,
                                   float *best_val, int32_t *best_k, cudaStream_t st) {
    const size_t smem_b = sizeof(float) * ((size_t)m->D * FULLB_R + FULLB_R * 8) + sizeof(int) * FULLB_R * 8;
    if (m->D <= 256 && smem_b <= 48 * 1024)
        refine_full_blocked_kernel<<<148 * 4, FULLB_THREADS, smem_b, st>>>(
            *m, fb_list, (const unsigned long long *)n_fallback, (long long)first, best_val, best_k);
    else
        refine_full_kernel<<<148 * 8, 256, sizeof(float) * (m->D + 16), st>>>(
            *m, fb_list, (const unsigned long long *)n_fallback, (long long)first, best_val, best_k);
    SEGB_LAUNCH_CHECK();
    return 0;
}
int launch_refine_full(const segb_kmeans *m, const int32_t *fb_list, const int64_t *n_fallback, float *best_val,
                       int32_t *best_k, cudaStream_t st) {
    return launch_refine_full_from(m, fb_list, n_fallback, 0, best_val, best_k, st);
}
}  // namespace mma
}  // namespace segb

extern "C" int64_t segb_mma_refine_work_bytes(int64_t n_emb, int32_t K_max) {
    (void)K_max;
    return (n_emb + 64) * (int64_t)sizeof(int32_t);
}

// first stage of both refine entry points: every row's surviving candidates re-scored exactly, the rows the
// record cannot decide appended to fb_list (count in n_fallback)
static int refine_rows_stage(const segb_kmeans *m, const void *cand, const float *x_err, const float *w_max,
                             int64_t n_emb, void *work, float *best_val, int32_t *best_k, int64_t *n_fallback,
                             void *stream, int fp8 = 0) {
    SEGB_CHECK_ARG(m && cand && x_err && w_max && work && best_val && best_k && n_fallback, "null pointer");
    SEGB_CHECK_ARG(!m->x_is_f64, "tensor-core scorer needs float32 embeddings");
    SEGB_CHECK_ARG(n_emb < (1ll << 31), "too many embeddings for one refine call");
    {   // the half-warp scheme covers NumPy's pairwise structure up to one split with blocks <= 128
        int n2 = m->D / 2; n2 -= n2 % 8;
        SEGB_CHECK_ARG(m->D <= 128 || (m->D <= 256 && m->D - n2 <= 128), "refine: unsupported D");
    }
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *fb_list = (int32_t *)work;
    SEGB_CUDA(cudaMemsetAsync(n_fallback, 0, sizeof(int64_t), st));
    const bool lanes8 = (m->D % 2 == 0);          // 8-byte loads need even D (rows 8-byte aligned)
    int64_t blocks = (n_emb * (lanes8 ? 8 : 16) + REFINE_THREADS - 1) / REFINE_THREADS;
    if (blocks > 148 * 16) blocks = 148 * 16;
    // accumulator steps of the longer NumPy block: <= 8 for D <= 143 when split (D = 130: 64 + 66 terms)
    int steps_max;
    {
        int n2 = 0;
        if (m->D > 128) { n2 = m->D / 2; n2 -= n2 % 8; }
        const int longest = (m->D > 128) ? (m->D - n2 > n2 ? m->D - n2 : n2) : m->D;
        steps_max = longest / 8;
    }
    if (fp8 && !lanes8) { set_error("e4m3 filter records need an even D"); return SEGB_E_UNSUPPORTED; }
    if (lanes8 && m->D == 130)               // the dimension of the BASELINE configurations, specialised
        refine_rows8_kernel<8, 130><<<(unsigned)blocks, REFINE_THREADS, 0, st>>>(
            *m, (const Cand *)cand, x_err, w_max, n_emb, k_pad(m->K_max) / CHUNK, best_val, best_k,
            (unsigned long long *)n_fallback, fb_list, fp8);
    else if (lanes8 && steps_max <= 8)
        refine_rows8_kernel<8><<<(unsigned)blocks, REFINE_THREADS, 0, st>>>(
            *m, (const Cand *)cand, x_err, w_max, n_emb, k_pad(m->K_max) / CHUNK, best_val, best_k,
            (unsigned long long *)n_fallback, fb_list, fp8);
    else if (lanes8)
        refine_rows8_kernel<REFINE_MAX_STEPS><<<(unsigned)blocks, REFINE_THREADS, 0, st>>>(
            *m, (const Cand *)cand, x_err, w_max, n_emb, k_pad(m->K_max) / CHUNK, best_val, best_k,
            (unsigned long long *)n_fallback, fb_list, fp8);
    else
        refine_rows_kernel<<<(unsigned)blocks, REFINE_THREADS, 0, st>>>(
            *m, (const Cand *)cand, x_err, w_max, n_emb, k_pad(m->K_max) / CHUNK, best_val, best_k,
            (unsigned long long *)n_fallback, fb_list);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_mma_refine(const segb_kmeans *m, const void *cand, const float *x_err, const float *w_max,
                               int64_t n_emb, void *work, float *best_val, int32_t *best_k, int64_t *n_fallback,
                               void *stream) {
    const int rc = refine_rows_stage(m, cand, x_err, w_max, n_emb, work, best_val, best_k, n_fallback, stream);
    if (rc) return rc;
    return launch_refine_full(m, (const int32_t *)work, n_fallback, best_val, best_k, (cudaStream_t)stream);
}

// ---- refine with the second-level tensor pass (segb_mma_refine2) --------------------------------------------
// work layout: [fb_list (n_emb + 64) int32 | n_unres (256 B) | unres_list (n_emb + 64) int32 | thr cap float |
//               fb_tiles cap * KP * 2 | bitmap cap * (K_pad / 32) * 4], every part 256-byte aligned;
// cap = undecided rows one ROUND of the second-level pass takes (the list is worked off in ceil(n_emb / cap) rounds).
namespace {
struct Work2 { int64_t off_unres_n, off_unres, off_thr, off_tiles, off_bitmap, total; };
inline int64_t al256(int64_t v) { return (v + 255) / 256 * 256; }
inline Work2 work2_layout(int64_t n_emb, int32_t K_max, int32_t D, int64_t cap) {
    Work2 w;
    w.off_unres_n = al256((n_emb + 64) * (int64_t)sizeof(int32_t));
    w.off_unres = w.off_unres_n + 256;
    w.off_thr = w.off_unres + al256((n_emb + 64) * (int64_t)sizeof(int32_t));       // any row may end up unresolved (NaN data)
    w.off_tiles = w.off_thr + al256(cap * 4);
    w.off_bitmap = w.off_tiles + al256(cap * (int64_t)kp_of(D) * 2);
    w.total = w.off_bitmap + al256(cap * (int64_t)(k_pad(K_max) / 32) * 4);
    return w;
}
inline int64_t default_cap(int64_t n_emb) {
    int64_t cap = rows_pad(n_emb / 8);
    if (cap < 16 * MT_ROWS) cap = 16 * MT_ROWS;
    if (cap > rows_pad(n_emb)) cap = rows_pad(n_emb);
    return cap;
}
}  // namespace

extern "C" int64_t segb_mma_refine2_work_bytes(int64_t n_emb, int32_t K_max, int32_t D) {
    return work2_layout(n_emb, K_max, D, default_cap(n_emb)).total;
}

extern "C" int segb_mma_refine2(const segb_kmeans *m, const void *x_tiles, const void *w_tiles, const void *cand,
                                const float *x_err, const float *w_max, int64_t n_emb, void *work, int64_t work_bytes,
                                int32_t max_rounds, float *best_val, int32_t *best_k, int64_t *n_fallback, void *stream) {
    SEGB_CHECK_ARG(m && x_tiles && w_tiles, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    // the largest cap (a multiple of 256 rows) whose layout fits the caller's buffer
    int64_t cap = default_cap(n_emb);
    while (cap > 0 && work2_layout(n_emb, m->K_max, m->D, cap).total > work_bytes) cap -= MT_ROWS;
    const bool second = cap > 0 && row8_supported(m->D) && row8_steps_max(m->D) <= REFINE_MAX_STEPS;
    if (!second) {
        SEGB_CHECK_ARG(work_bytes >= segb_mma_refine_work_bytes(n_emb, m->K_max), "refine2: work buffer too small");
        return segb_mma_refine(m, cand, x_err, w_max, n_emb, work, best_val, best_k, n_fallback, stream);
    }
    int rc = refine_rows_stage(m, cand, x_err, w_max, n_emb, work, best_val, best_k, n_fallback, stream);
    if (rc) return rc;
    const Work2 w = work2_layout(n_emb, m->K_max, m->D, cap);
    uint8_t *wb = (uint8_t *)work;
    const int32_t *fb_list = (const int32_t *)work;
    unsigned long long *n_unres = (unsigned long long *)(wb + w.off_unres_n);
    int32_t *unres_list = (int32_t *)(wb + w.off_unres);
    float *thr = (float *)(wb + w.off_thr);
    uint8_t *fb_tiles = wb + w.off_tiles;
    uint32_t *bitmap = (uint32_t *)(wb + w.off_bitmap);
    const unsigned long long *n_fb = (const unsigned long long *)n_fallback;
    SEGB_CUDA(cudaMemsetAsync(n_unres, 0, sizeof(unsigned long long), st));
    FilterLaunch f;
    f.x_tiles = fb_tiles; f.w_tiles = w_tiles; f.cand = nullptr; f.n_emb = cap;       // grid sized for cap; rows from the device
    f.w_rows_pad = k_pad(m->K_max); f.w_rows = m->K_max; f.KP = kp_of(m->D); f.D = m->D; f.x_max = w_max; f.w_max = w_max;
    f.n_chunks = 1; f.tau_kind = TAU_KMEANS; f.tau_T = 0.f;
    const int n_words = k_pad(m->K_max) / 32;
    // the undecided list is worked off in rounds of `cap` rows: the launch count is fixed (the list length lives on
    // the device), rounds past its end find nothing to do
    int64_t first = 0;
    for (int round = 0; first < n_emb && (max_rounds <= 0 || round < max_rounds); first += cap, ++round) {
        gather_undecided_kernel<<<148 * 8, 256, 0, st>>>((const uint8_t *)x_tiles, (const Cand *)cand, x_err, w_max, fb_list, n_fb,
                                                         first, cap, m->D, kp_of(m->D), fb_tiles, thr);
        SEGB_LAUNCH_CHECK();
        rc = launch_filter_impl(f, n_fb, first, cap, thr, bitmap, st);
        if (rc) return rc;
        if (m->D == 130)
            refine_bitmap_kernel<8, 130><<<148 * 8, REFINE_THREADS, 0, st>>>(*m, fb_list, n_fb, first, cap, bitmap, n_words, best_val,
                                                                             best_k, n_unres, unres_list);
        else if (row8_steps_max(m->D) <= 8)
            refine_bitmap_kernel<8><<<148 * 8, REFINE_THREADS, 0, st>>>(*m, fb_list, n_fb, first, cap, bitmap, n_words, best_val,
                                                                        best_k, n_unres, unres_list);
        else
            refine_bitmap_kernel<REFINE_MAX_STEPS><<<148 * 8, REFINE_THREADS, 0, st>>>(*m, fb_list, n_fb, first, cap, bitmap, n_words,
                                                                                       best_val, best_k, n_unres, unres_list);
        SEGB_LAUNCH_CHECK();
    }
    // undecided rows beyond the rounds the caller asked for, and rows the bitmap pass could not resolve (empty bitmap:
    // NaN scores): exhaustive exact scan
    if (first < n_emb) { rc = launch_refine_full_from(m, fb_list, n_fallback, first, best_val, best_k, st); if (rc) return rc; }
    return launch_refine_full_from(m, unres_list, (const int64_t *)n_unres, 0, best_val, best_k, st);
}


// ---- e4m3 first-level filter (segb_mma8_*) -------------------------------------------------------------------
static inline int64_t rows_pad8(int64_t n) { return (n + MT_ROWS - 1) / MT_ROWS * MT_ROWS; }

extern "C" int64_t segb_mma8_x_tiles_bytes(int64_t n_emb, int32_t D) { return rows_pad8(n_emb) * kp8_of(D); }
extern "C" int64_t segb_mma8_w_tiles_bytes(int32_t K_max, int32_t D) { return (int64_t)k_pad(K_max) * kp8_of(D); }

extern "C" int segb_mma8_pack_x(const float *X, int64_t n_emb, int32_t D, float scale, void *x_tiles8, float *x_err8,
                                float *x_max8, void *stream) {
    SEGB_CHECK_ARG(X && x_tiles8 && x_err8 && x_max8 && n_emb > 0 && D > 0 && scale > 0.f, "null pointer");
    const int64_t np_ = rows_pad8(n_emb);
    const int wpb = 8;
    SEGB_CUDA(cudaMemsetAsync(x_max8, 0, 2 * sizeof(float), (cudaStream_t)stream));
    pack_x8_kernel<<<(unsigned)((np_ + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        X, n_emb, np_, D, kp8_of(D), scale, (uint8_t *)x_tiles8, x_err8, x_max8);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_mma8_pack_means(const float *means, int32_t K_max, int32_t D, float scale, void *w_tiles8, float *w_err8,
                                    float *w_max8, void *stream) {
    SEGB_CHECK_ARG(means && w_tiles8 && w_err8 && w_max8 && K_max > 0 && D > 0 && scale > 0.f, "null pointer");
    const int kp_rows = k_pad(K_max);
    const int wpb = 8;
    pack_w8_kernel<<<(kp_rows + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(
        means, K_max, kp_rows, D, kp8_of(D), scale, (uint8_t *)w_tiles8, w_err8);
    SEGB_LAUNCH_CHECK();
    wmax4_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(w_err8, K_max, w_max8);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_mma8_filter(const void *x_tiles8, const void *w_tiles8, int64_t n_emb, int32_t K_max, int32_t D,
                                const float *x_max8, const float *w_max8, void *cand, void *stream) {
    SEGB_CHECK_ARG(x_tiles8 && w_tiles8 && cand && x_max8 && w_max8 && n_emb > 0 && K_max > 0, "null pointer");
    FilterLaunch f;
    f.x_tiles = x_tiles8; f.w_tiles = w_tiles8; f.cand = cand; f.n_emb = n_emb;
    f.w_rows_pad = k_pad(K_max); f.w_rows = K_max; f.KP = kp8_of(D); f.D = D; f.x_max = x_max8; f.w_max = w_max8;
    f.n_chunks = 1; f.tau_kind = TAU_KMEANS_FP8; f.tau_T = 0.f; f.fp8 = 1;
    return launch_filter(f, (cudaStream_t)stream);
}

extern "C" int segb_mma8_refine(const segb_kmeans *m, const void *cand, const float *x_err8, const float *w_max8, float scale,
                                const void *w_tiles16, const float *w_max16, int64_t n_emb, void *work, int64_t work_bytes,
                                int32_t max_rounds, float *best_val, int32_t *best_k, int64_t *n_fallback, void *stream) {
    SEGB_CHECK_ARG(m && cand && x_err8 && w_max8 && w_tiles16 && w_max16 && scale > 0.f, "null pointer");
    SEGB_CHECK_ARG(row8_supported(m->D) && row8_steps_max(m->D) <= REFINE_MAX_STEPS, "e4m3 scorer: unsupported D");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t cap = default_cap(n_emb);
    while (cap > 0 && work2_layout(n_emb, m->K_max, m->D, cap).total > work_bytes) cap -= MT_ROWS;
    SEGB_CHECK_ARG(cap > 0, "segb_mma8_refine: work buffer too small (segb_mma_refine2_work_bytes)");
    int rc = refine_rows_stage(m, cand, x_err8, w_max8, n_emb, work, best_val, best_k, n_fallback, stream, 1);
    if (rc) return rc;
    const Work2 w = work2_layout(n_emb, m->K_max, m->D, cap);
    uint8_t *wb = (uint8_t *)work;
    const int32_t *fb_list = (const int32_t *)work;
    unsigned long long *n_unres = (unsigned long long *)(wb + w.off_unres_n);
    int32_t *unres_list = (int32_t *)(wb + w.off_unres);
    float *thr = (float *)(wb + w.off_thr);
    uint8_t *fb_tiles = wb + w.off_tiles;
    uint32_t *bitmap = (uint32_t *)(wb + w.off_bitmap);
    const unsigned long long *n_fb = (const unsigned long long *)n_fallback;
    SEGB_CUDA(cudaMemsetAsync(n_unres, 0, sizeof(unsigned long long), st));
    FilterLaunch f;
    f.x_tiles = fb_tiles; f.w_tiles = w_tiles16; f.cand = nullptr; f.n_emb = cap;
    f.w_rows_pad = k_pad(m->K_max); f.w_rows = m->K_max; f.KP = kp_of(m->D); f.D = m->D; f.x_max = w_max16; f.w_max = w_max16;
    f.n_chunks = 1; f.tau_kind = TAU_KMEANS; f.tau_T = 0.f;
    const int n_words = k_pad(m->K_max) / 32;
    int64_t first = 0;
    for (int round = 0; first < n_emb && (max_rounds <= 0 || round < max_rounds); first += cap, ++round) {   // rounds (see segb_mma_refine2)
        gather_convert_undecided_kernel<<<148 * 8, 256, 0, st>>>((const float *)m->X, (const Cand *)cand, x_err8, w_max8, scale,
                                                                 w_max16, fb_list, n_fb, first, cap, m->D, kp_of(m->D), fb_tiles, thr);
        SEGB_LAUNCH_CHECK();
        rc = launch_filter_impl(f, n_fb, first, cap, thr, bitmap, st);
        if (rc) return rc;
        if (m->D == 130)
            refine_bitmap_kernel<8, 130><<<148 * 8, REFINE_THREADS, 0, st>>>(*m, fb_list, n_fb, first, cap, bitmap, n_words, best_val,
                                                                             best_k, n_unres, unres_list);
        else if (row8_steps_max(m->D) <= 8)
            refine_bitmap_kernel<8><<<148 * 8, REFINE_THREADS, 0, st>>>(*m, fb_list, n_fb, first, cap, bitmap, n_words, best_val,
                                                                        best_k, n_unres, unres_list);
        else
            refine_bitmap_kernel<REFINE_MAX_STEPS><<<148 * 8, REFINE_THREADS, 0, st>>>(*m, fb_list, n_fb, first, cap, bitmap, n_words,
                                                                                       best_val, best_k, n_unres, unres_list);
        SEGB_LAUNCH_CHECK();
    }
    if (first < n_emb) { rc = launch_refine_full_from(m, fb_list, n_fallback, first, best_val, best_k, st); if (rc) return rc; }
    return launch_refine_full_from(m, unres_list, (const int64_t *)n_unres, 0, best_val, best_k, st);
}
