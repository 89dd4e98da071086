// Fused scoring: fp32 embeddings in, exact per-embedding results out, in ONE kernel.
//
// The two-kernel scorers (kmeans_mma.cu / fixedvar_filter.cu: filter GEMM over a pre-packed fp16 tile
// image of X, then a refine kernel) read the embeddings twice -- 6.05 GB of fp16 image in the filter plus
// 10.9 GB of fp32 rows in the refine for 10.9 GB of data -- keep both copies resident, write 0.67 GB of
// filter records and spend 17 % of the sweep in a latency-bound refine.  This kernel reads the fp32
// embeddings ONCE from HBM:
//
//   convert warps (4)  turn the NEXT work item's 256 fp32 rows into the fp16 UMMA tile layout directly in
//                      shared memory (8 rows in flight per warp, coalesced 8-byte loads; the rounding-error
//                      norms behind the rigorous candidate threshold fall out of the same pass, per row)
//   refine warps (8)   re-score the PREVIOUS work item's surviving candidates exactly (eight lanes per
//                      embedding; the rows were just read, so they come from L2) while the tensor pipe works
//                      on the current one
//   warp 0             bulk-copy producer of the model tiles (B operand, L2-resident)
//   warp 1             tcgen05.mma issuer (one elected thread; M=128 N=128 K=16, fp32 TMEM accumulators)
//   warp 2             TMEM allocator
//   epilogue warps (16) running top-3 chunk maxima per embedding, as in kmeans_filter_kernel, but with the
//                      embedding's OWN threshold; the per-row record goes to shared memory, not HBM
//
// Per embedding 8 bytes come back (k-means: best value + component; FBGMM: float64 log_marg_i + MAP slot).
// Policies: POL_KMEANS (max / first argmax of -|mu_k - x|^2, float32 NumPy order, bit-exact;
// kmeans_components.py:225-232) and POL_FV (FBGMM.log_marg_i, isotropic variances; fbgmm.py:256-285).
#include "fv_refine.cuh"
#include "refine_rows.cuh"

namespace segb {
namespace fused {

using namespace segb::mma;

constexpr uint32_t KSTEP_BYTES = 2 * (TILE_ROWS / 8) * 128;        // one K=16 step of a 128-row tile image
constexpr int EPI_PARTS = 2;
constexpr int N_EPI_WARPS = 8 * EPI_PARTS, N_CVT_WARPS = 4, N_REF_WARPS = 8;
constexpr int CVT_WARP0 = 4 + N_EPI_WARPS, REF_WARP0 = CVT_WARP0 + N_CVT_WARPS;
constexpr int N_THREADS = 32 * (4 + N_EPI_WARPS + N_CVT_WARPS + N_REF_WARPS);    // 1024: 64 registers per thread
constexpr int N_CVT = N_CVT_WARPS + 2;                             // warps 2 and 3 (idle after the TMEM allocation) convert too
constexpr uint32_t TMEM_COLS = 512;
constexpr int POL_KMEANS = 0, POL_FV = 1;
constexpr int XBUF_BYTES = MT_ROWS * 32, TAU_BYTES = 2 * MT_ROWS * 4, BAR_BYTES = 256;

struct Params {
    const float *X;
    int64_t n_emb;
    int32_t D, KP;
    const uint8_t *w_tiles;
    int32_t n_mtiles, n_ntiles, n_ksteps, n_chunks;
    uint32_t tile_bytes;
    const float *w_max;            // k-means: (e_mu, n_mu); FBGMM: (eW, nW, |A|, p/2)
    float tau_T;
    // k-means refine
    const float *means;
    int32_t K_max;
    float *best_val;
    int32_t *best_k;
    // FBGMM refine
    const double *model_rows;
    double *log_marg;
    int32_t *map_k;
    unsigned long long *n_fallback;
    int32_t *fb_list;
    RowRec *rec_out;               // optional [n_emb]: the row records, for a later component draw (segb_fvf_choose_tokens)
    int32_t dbg;                   // development: bit 0 skip the conversion work, bit 1 skip the refine work (timing only)
};

// One convert warp turns CVT_ROWS consecutive fp32 rows of a work item into fp16 operand rows of the tile
// image in shared memory, eight rows at a time.  Lane (rr, pp) = (lane / 4, lane % 4) owns elements
// (8 ch + 2 pp, + 1) of row rr of the group for every 8-column chunk ch: one store instruction of the warp
// writes ONE 8x8 core matrix of the UMMA layout -- 128 contiguous bytes, no bank conflicts (the tensor pipe
// reads its operands from the same shared memory at close to its full bandwidth, so conflicted stores cost
// MMA time) -- and one load instruction fetches eight full 32-byte sectors.  The row norm (input of the row's
// candidate threshold) is reduced over the four lanes of a row; the rounding-error norm is bounded by
// 2^-11 |x| + sqrt(D) 2^-25 (fp16 has 11 significant bits; the second term covers fp16 subnormals) instead
// of being measured -- a looser threshold costs a few more exact re-scores, never exactness.  Even D.
// k-means columns [x^ (D), 1, 1, 1, 0..]; FBGMM [x^ (D), 1, 1, 1, n2h, n2h, n2l, 0..] (fixedvar_filter.cu).
template <int POL>
__device__ __forceinline__ void convert_rows_warp(const Params &p, int mt, uint8_t *sA_ab, uint32_t tb, float *tau_ab,
                                                  int wc, int lane) {
    const int D = p.D, KP = p.KP;
    const int rr = lane >> 2, pp = lane & 3;
    const int64_t row0 = (int64_t)mt * MT_ROWS;
    const int n_data = (D + 7) / 8, n_ch = KP / 8;               // chunks holding data columns / all chunks
    const float w0 = p.w_max[0], w1 = p.w_max[1];
    const W4 w4{w0, w1, POL == POL_FV ? p.w_max[2] : 0.f, POL == POL_FV ? p.w_max[3] : 0.f};
    const float sub_err = sqrtf((float)D) * ldexpf(1.f, -25);
    constexpr int UNR = 9;                                        // loads in flight per lane
#pragma unroll 1
    for (int b = 8 * wc; b < MT_ROWS; b += 8 * N_CVT) {           // 8-row groups of the work item, round-robin over the convert warps
        const int r = (b & (TILE_ROWS - 1)) + rr;
        const int64_t row = row0 + b + rr;
        const bool live = row < p.n_emb;
        const float *xr = p.X + row * D + 2 * pp;
        uint8_t *dst = sA_ab + (size_t)(b / TILE_ROWS) * tb + tile_off(r, 2 * pp);    // + ch * (TILE_ROWS / 8) * 128 per chunk
        float f2 = 0.f;
        float2 last = make_float2(0.f, 0.f);                      // the chunk that also holds constant columns
#pragma unroll 1
        for (int c0 = 0; c0 < n_data; c0 += UNR) {
            float2 v[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int ch = c0 + u;
                v[u] = (live && ch < n_data && 8 * ch + 2 * pp < D) ? *reinterpret_cast<const float2 *>(xr + 8 * ch)
                                                                   : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int ch = c0 + u;
                if (ch < n_data) {
                    f2 = fmaf(v[u].x, v[u].x, fmaf(v[u].y, v[u].y, f2));
                    if (8 * ch + 8 <= D)
                        *reinterpret_cast<__half2 *>(dst + (size_t)ch * (TILE_ROWS / 8) * 128) = __floats2half2_rn(v[u].x, v[u].y);
                    else last = v[u];
                }
            }
        }
        f2 += __shfl_xor_sync(FULL, f2, 1);
        f2 += __shfl_xor_sync(FULL, f2, 2);                       // |x|^2 of the row on its four lanes
        __half n2h = __float2half_rn(0.f), n2l = n2h;
        if (POL == POL_FV) { n2h = __float2half_rn(f2); n2l = __float2half_rn(f2 - __half2float(n2h)); }
        // chunks from the first one that holds a column >= D: data tail, constants, zeros
        for (int ch = D / 8; ch < n_ch; ++ch) {
            float o[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = 8 * ch + 2 * pp + e;
                float val = 0.f;
                if (live) {
                    if (c < D) val = e == 0 ? last.x : last.y;
                    else {
                        const int ecol = c - D;
                        if (ecol < 3) val = 1.f;
                        else if (POL == POL_FV && (ecol == 3 || ecol == 4)) val = __half2float(n2h);
                        else if (POL == POL_FV && ecol == 5) val = __half2float(n2l);
                    }
                }
                o[e] = val;
            }
            *reinterpret_cast<__half2 *>(dst + (size_t)ch * (TILE_ROWS / 8) * 128) = __floats2half2_rn(o[0], o[1]);
        }
        if (pp == 0) {
            float tau = CUDART_INF_F;
            if (live) {
                const float nx = sqrtf(f2) * 1.0001f;
                const float ex = (nx < 60000.f) ? (ldexpf(nx, -11) + sub_err) : CUDART_INF_F;   // |x_d| > 60000 => |x| > 60000
                tau = POL == POL_KMEANS ? filter_tau(ex, nx, w0, w1, D) : lse_tau(ex, nx, w4, KP, p.tau_T);
            }
            tau_ab[b + rr] = tau;
        }
    }
}

// KS = K=16 steps of the inner dimension (compile-time unrolled issue loop), 0 = runtime.
// MAXS = accumulator steps of the k-means 8-lane refine (refine_rows.cuh); unused for POL_FV.
template <int POL, int KS, int MAXS>
__global__ void __launch_bounds__(N_THREADS, 1) score_fused_kernel(Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tb = p.tile_bytes;
    const int n_ks = KS > 0 ? KS : p.n_ksteps;
    uint8_t *sA = smem;                                        // 2 buffers x 2 tiles
    uint8_t *sB = smem + 4 * (size_t)tb;                       // 2 stages = accumulator buffer = tile parity
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + 2 * (size_t)tb);
    uint8_t *xbuf = reinterpret_cast<uint8_t *>(bars) + BAR_BYTES;   // [MT_ROWS] x 32 B: merge slots, then the row records
    float *tau_s = reinterpret_cast<float *>(xbuf + XBUF_BYTES);     // [2][MT_ROWS]
    constexpr int A_FULL = 0, A_EMPTY = 2, B_FULL = 4, MMA_DONE = 6, ACC_EMPTY = 8, CAND_FULL = 10, CAND_EMPTY = 11,
                  N_BARS = 12;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + N_BARS);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };

    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(BAR(A_FULL + b), N_CVT);
            mbar_init(BAR(A_EMPTY + b), 1);
            mbar_init(BAR(B_FULL + b), 1);
            mbar_init(BAR(MMA_DONE + b), 1);
            mbar_init(BAR(ACC_EMPTY + b), N_EPI_WARPS);
        }
        mbar_init(BAR(CAND_FULL), N_EPI_WARPS / EPI_PARTS);
        mbar_init(BAR(CAND_EMPTY), N_REF_WARPS);
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
        // ===================== model-tile producer =====================
        if (elect_one()) {
            uint32_t n_use = 0;
            for (int mt = blockIdx.x; mt < p.n_mtiles; mt += gridDim.x) {
                for (int nt = 0; nt < p.n_ntiles; ++nt, ++n_use) {
                    const uint32_t s = n_use & 1, use = n_use >> 1;
                    mbar_wait(BAR(MMA_DONE + s), (use & 1) ^ 1);       // the MMAs of the stage's previous tile are done
                    mbar_expect_tx(BAR(B_FULL + s), tb);
                    bulk_g2s(smem_u32(sB + (size_t)s * tb), p.w_tiles + (size_t)nt * tb, tb, BAR(B_FULL + s));
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (as kmeans_filter_kernel) =====================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_mn(TILE_ROWS, NT_COLS);
            constexpr uint32_t KSTEP = KSTEP_BYTES >> 4;
            const uint32_t a_lo_base = make_desc_lo(smem_u32(sA), TILE_ROWS);
            const uint32_t b_lo_base = make_desc_lo(smem_u32(sB), TILE_ROWS);
            uint32_t n_use = 0;
            int it = 0;
            for (int mt = blockIdx.x; mt < p.n_mtiles; mt += gridDim.x, ++it) {
                const int ab = it & 1;
                mbar_wait(BAR(A_FULL + ab), (uint32_t)(it >> 1) & 1);
                tc_fence_after();
                const uint32_t a_lo0 = a_lo_base + (uint32_t)(2 * ab) * (tb >> 4), a_lo1 = a_lo0 + (tb >> 4);
                for (int nt = 0; nt < p.n_ntiles; ++nt, ++n_use) {
                    const uint32_t buf = n_use & 1, use = n_use >> 1;
                    mbar_wait2(BAR(B_FULL + buf), use & 1, BAR(ACC_EMPTY + buf), (use & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t b_lo = b_lo_base + buf * (tb >> 4);
                    const uint32_t d0 = tmem_base + (buf * 2) * NT_COLS, d1 = d0 + NT_COLS;
                    if (KS > 0) {
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
                    tc_commit(BAR(MMA_DONE + buf));       // B stage free + accumulators ready
                }
                tc_commit(BAR(A_EMPTY + ab));             // A tiles free
            }
        }
    } else if (warp >= 4 && warp < CVT_WARP0) {
        // ===================== epilogue: running top-3 chunk maxima per embedding =====================
        const int e = warp - 4, q = warp & 3, h = (e >> 2) & 1, part = e >> 3;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        constexpr int COLS = NT_COLS / EPI_PARTS;
        static_assert(COLS == 64, "two x32 loads per warp and tile");
        const int r_local = h * TILE_ROWS + q * 32 + lane;
        float4 *merge = reinterpret_cast<float4 *>(xbuf);
        uint32_t n_use = 0;
        int it = 0;
        for (int mt = blockIdx.x; mt < p.n_mtiles; mt += gridDim.x, ++it) {
            const int ab = it & 1;
            mbar_wait(BAR(A_FULL + ab), (uint32_t)(it >> 1) & 1);      // the row thresholds of this work item are in place
            const float tau_row = tau_s[ab * MT_ROWS + r_local];
            float m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F;
            int i1 = -1, i2 = -1;
            uint32_t k1 = 0, k2 = 0;
            for (int nt = 0; nt < p.n_ntiles; ++nt, ++n_use) {
                const uint32_t buf = n_use & 1, acc_phase = (n_use >> 1) & 1;
                mbar_wait(BAR(MMA_DONE + buf), acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + lane_base + (buf * 2 + h) * NT_COLS + part * COLS;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float v[32];
                    tc_ld32_wait(taddr + half * 32, v);
                    if (half == 1) {                                   // accumulator drained: hand it back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(BAR(ACC_EMPTY + buf));
                    }
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        float cm = v[c * 16];
#pragma unroll
                        for (int j = 1; j < 16; ++j) cm = fmaxf(cm, v[c * 16 + j]);
                        top3_insert(&v[c * 16], cm, nt * (NT_COLS / CHUNK) + (part * COLS) / CHUNK + half * 2 + c, tau_row,
                                    m1, m2, m3, i1, i2, k1, k2);
                    }
                }
            }
            // the record slots are free once the aux warps have re-scored the previous work item
            mbar_wait(BAR(CAND_EMPTY), ((uint32_t)it & 1) ^ 1);
            if (part == 1) {
                merge[2 * r_local] = make_float4(m1, m2, m3, __int_as_float(i1));
                merge[2 * r_local + 1] = make_float4(__int_as_float(i2), __uint_as_float(k1), __uint_as_float(k2), 0.f);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * N_EPI_WARPS) : "memory");
            if (part == 0) {
                const float4 a = merge[2 * r_local], b = merge[2 * r_local + 1];
                top3_merge(a.x, __float_as_int(a.w), __float_as_uint(b.y), m1, m2, m3, i1, i2, k1, k2);
                top3_merge(a.y, __float_as_int(b.x), __float_as_uint(b.z), m1, m2, m3, i1, i2, k1, k2);
                if (a.z > m3) m3 = a.z;
                Cand c;
                c.m1 = m1; c.m2 = m2; c.m3 = m3; c.i1 = i1; c.i2 = i2;
                RowRec rec;
                rec.i1 = i1; rec.i2 = i2; rec.masks = (k1 & 0xffffu) | (k2 << 16);
                rec.code = refine_decide(c, tau_row, p.n_chunks);
                *reinterpret_cast<RowRec *>(xbuf + 32 * r_local) = rec;      // over this row's own merge slot
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(CAND_FULL));
            }
        }
    } else if (warp == 2 || warp == 3 || (warp >= CVT_WARP0 && warp < REF_WARP0)) {
        // ===================== convert: the NEXT work item's fp32 rows -> fp16 operand tiles =====================
        const int wc = warp < 4 ? warp - 2 : warp - CVT_WARP0 + 2;
        int it = 0;
        long long t_wait = 0, t_cvt = 0;
        for (int mt = blockIdx.x; mt < p.n_mtiles; mt += gridDim.x, ++it) {
            const int ab = it & 1;
            const long long t0 = clock64();
            // Ask the copy engine to bring the work item AFTER this one into L2 now: by the time this warp
            // converts it (one work-item period later) its loads hit L2 instead of paying HBM latency
            // (the conversion is a chain of dependent batches: 16 x HBM latency per item was ~55 us
            // against a 35 us MMA period).
            if (wc == 0 && lane == 0) {
                const int64_t r_pf = ((int64_t)mt + gridDim.x) * MT_ROWS;
                if (r_pf < p.n_emb) {
                    int64_t n_r = p.n_emb - r_pf;
                    if (n_r > MT_ROWS) n_r = MT_ROWS;
                    const uint32_t bytes = (uint32_t)((n_r * p.D * 4) & ~15ll);
                    const float *src = p.X + r_pf * p.D;
                    if (bytes && (reinterpret_cast<uintptr_t>(src) & 15) == 0)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
                }
            }
            mbar_wait(BAR(A_EMPTY + ab), ((uint32_t)(it >> 1) & 1) ^ 1);       // the MMAs that read this buffer are done
            const long long t1 = clock64();
            if (!(p.dbg & 1)) convert_rows_warp<POL>(p, mt, sA + (size_t)(2 * ab) * tb, tb, tau_s + ab * MT_ROWS, wc, lane);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic writes -> tensor-core reads
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(A_FULL + ab));
            t_wait += t1 - t0; t_cvt += clock64() - t1;
        }
        if ((p.dbg & 4) && blockIdx.x == 0 && lane == 0)
            printf("cvt warp %d: items %d, avg wait %lld clk, avg convert %lld clk\n", wc, it, t_wait / max(it, 1), t_cvt / max(it, 1));
    } else if (warp >= REF_WARP0) {
        // ===================== refine: the PREVIOUS work item's surviving candidates, exactly =====================
        const int j = lane & 7, grp = (warp - REF_WARP0) * 4 + (lane >> 3);
        const unsigned gmask = 0xffu << (lane & 24);
        const Row8Geom geo(p.D, lane);
        const fvf::ModelRows tabs = fvf::model_view(p.model_rows, p.K_max, p.D, 0);
        int it = 0;
        for (int mt = blockIdx.x; mt < p.n_mtiles; mt += gridDim.x, ++it) {
            mbar_wait(BAR(CAND_FULL), (uint32_t)it & 1);
#pragma unroll 1
            for (int r = grp; r < MT_ROWS; r += 4 * N_REF_WARPS) {
                const int64_t row = (int64_t)mt * MT_ROWS + r;
                if (row >= p.n_emb || (p.dbg & 2)) break;              // uniform within the 8-lane group
                const RowRec rec = *reinterpret_cast<const RowRec *>(xbuf + 32 * r);
                if (p.rec_out && j == 0) p.rec_out[row] = rec;
                if (rec.code == -2) {
                    if (j == 0) p.fb_list[atomicAdd(p.n_fallback, 1ull)] = (int32_t)row;
                    continue;
                }
                const float *xr = p.X + row * p.D;
                if (POL == POL_KMEANS) {
                    float bv;
                    int bk;
                    km_exact_row8<MAXS>(p.means, p.K_max, p.D, xr, rec.i1, rec.i2, rec.masks, rec.code, geo, bv, bk);
                    if (j == 0) { p.best_val[row] = bv; p.best_k[row] = (bk == 0x7fffffff) ? -1 : bk; }
                } else {
                    const fvf::LseAcc acc = fvf::fv_exact_row8<false>(tabs, p.K_max + 1, p.D, xr, rec.i1, rec.i2,
                                                                      rec.masks, rec.code, j, gmask);
                    if (j == 0) {
                        p.log_marg[row] = acc.lse();
                        if (p.map_k) p.map_k[row] = (acc.bk == 0x7fffffff) ? -1 : acc.bk;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(CAND_EMPTY));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

static inline int64_t rows_pad(int64_t n) { return (n + MT_ROWS - 1) / MT_ROWS * MT_ROWS; }

template <int POL>
static int launch(Params &p, cudaStream_t st) {
    p.n_mtiles = (int32_t)(rows_pad(p.n_emb) / MT_ROWS);
    p.n_ksteps = p.KP / 16;
    p.tile_bytes = (uint32_t)((int64_t)TILE_ROWS * p.KP * 2);
    const size_t smem = 6 * (size_t)p.tile_bytes + BAR_BYTES + XBUF_BYTES + TAU_BYTES;
    if (smem > 227 * 1024) { set_error("inner dimension %d too large for the fused scorer", p.KP); return SEGB_E_UNSUPPORTED; }
    int n_sm = 0;
    { const int rc = device_info(nullptr, &n_sm, nullptr); if (rc) return rc; }
    const int grid = p.n_mtiles < n_sm ? p.n_mtiles : n_sm;
    void (*kern)(Params);
    if (POL == POL_KMEANS) {
        const int sm = row8_steps_max(p.D);
        if (!row8_supported(p.D) || sm > 16) { set_error("fused k-means scorer: unsupported D=%d", p.D); return SEGB_E_UNSUPPORTED; }
        if (p.n_ksteps == 9 && sm <= 8) kern = score_fused_kernel<POL_KMEANS, 9, 8>;
        else if (sm <= 8) kern = score_fused_kernel<POL_KMEANS, 0, 8>;
        else kern = score_fused_kernel<POL_KMEANS, 0, 16>;
    } else {
        kern = p.n_ksteps == 9 ? score_fused_kernel<POL_FV, 9, 8> : score_fused_kernel<POL_FV, 0, 8>;
    }
    SEGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, N_THREADS, smem, st>>>(p);
    SEGB_LAUNCH_CHECK();
    return 0;
}

}  // namespace fused

// exhaustive scans of the rows the filter could not decide (kmeans_mma.cu / fixedvar_filter.cu)
namespace mma { int launch_refine_full(const segb_kmeans *m, const int32_t *fb_list, const int64_t *n_fallback,
                                       float *best_val, int32_t *best_k, cudaStream_t st); }
namespace fvf { int launch_full(const float *X, int D, int K_max, int aniso, const void *model, const int32_t *fb_list,
                                const int64_t *n_fallback, double *log_marg, int32_t *map_k, cudaStream_t st); }
}  // namespace segb

using namespace segb;

extern "C" int segb_fused_kmeans_best(const segb_kmeans *m, const void *w_tiles, const float *w_max, int64_t n_emb,
                                      void *work, float *best_val, int32_t *best_k, int64_t *n_fallback, void *stream) {
    SEGB_CHECK_ARG(m && w_tiles && w_max && work && best_val && best_k && n_fallback, "null pointer");
    SEGB_CHECK_ARG(!m->x_is_f64, "tensor-core scorer needs float32 embeddings");
    SEGB_CHECK_ARG(n_emb > 0 && n_emb <= m->n_emb && n_emb < (1ll << 31), "row count");
    cudaStream_t st = (cudaStream_t)stream;
    fused::Params p;
    memset(&p, 0, sizeof(p));
    p.X = (const float *)m->X; p.n_emb = n_emb; p.D = m->D; p.KP = (m->D + 3 + 15) / 16 * 16;
    p.w_tiles = (const uint8_t *)w_tiles; p.w_max = w_max;
    const int k_pad = (m->K_max + mma::NT_COLS - 1) / mma::NT_COLS * mma::NT_COLS;
    p.n_ntiles = k_pad / mma::NT_COLS; p.n_chunks = k_pad / mma::CHUNK;
    p.means = (const float *)m->means; p.K_max = m->K_max; p.best_val = best_val; p.best_k = best_k;
    p.n_fallback = (unsigned long long *)n_fallback; p.fb_list = (int32_t *)work;
    SEGB_CUDA(cudaMemsetAsync(n_fallback, 0, sizeof(int64_t), st));
    if (const char *e = getenv("SEGB_FUSED_DBG")) p.dbg = atoi(e);
    const int rc = fused::launch<fused::POL_KMEANS>(p, st);
    if (rc) return rc;
    return mma::launch_refine_full(m, p.fb_list, n_fallback, best_val, best_k, st);
}

extern "C" int segb_fused_fv_log_marg(const float *X, int64_t n_emb, int32_t D, int32_t K_max, const void *w_tiles,
                                      const void *model, const float *w_max, float T, void *work, double *log_marg,
                                      int32_t *map_k, void *rec_out, int64_t *n_fallback, void *stream) {
    SEGB_CHECK_ARG(X && w_tiles && model && w_max && work && log_marg && n_fallback, "null pointer");
    SEGB_CHECK_ARG(n_emb > 0 && n_emb < (1ll << 31) && T > 0.f, "row count / threshold");
    cudaStream_t st = (cudaStream_t)stream;
    fused::Params p;
    memset(&p, 0, sizeof(p));
    p.X = X; p.n_emb = n_emb; p.D = D; p.KP = fvf::kp_of(D, 0);
    p.w_tiles = (const uint8_t *)w_tiles; p.w_max = w_max; p.tau_T = T;
    const int w_pad = (K_max + 1 + mma::NT_COLS - 1) / mma::NT_COLS * mma::NT_COLS;
    p.n_ntiles = w_pad / mma::NT_COLS; p.n_chunks = w_pad / mma::CHUNK;
    p.model_rows = (const double *)model; p.K_max = K_max; p.log_marg = log_marg; p.map_k = map_k;
    p.n_fallback = (unsigned long long *)n_fallback; p.fb_list = (int32_t *)work; p.rec_out = (mma::RowRec *)rec_out;
    SEGB_CUDA(cudaMemsetAsync(n_fallback, 0, sizeof(int64_t), st));
    const int rc = fused::launch<fused::POL_FV>(p, st);
    if (rc) return rc;
    return fvf::launch_full(X, D, K_max, 0, model, p.fb_list, n_fallback, log_marg, map_k, st);
}
