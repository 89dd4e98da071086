// Segmentation DP over landmarks: forward filter + backward sample / Viterbi.
//
// Replaces (reference paths relative to segmentalist/):
//   forward_backward                unigram_acoustic_wordseg.py:653-756   (mode 0)
//   forward_backward_viterbi        unigram_acoustic_wordseg.py:759-864   (mode 1)
//   forward_backward_kmeans_viterbi kmeans_acoustic_wordseg.py:449-555    (mode 2)
//
// One warp per utterance.  The span-limited recurrence keeps log_alphas in
// shared memory; lane i of the warp owns the candidate with span i+1, the
// per-position max / logsumexp / argmax / inverse-CDF draw run on warp
// shuffles.  Sums are formed in the reference's element order (ascending start
// landmark for logsumexp, ascending span for the draw) so that results differ
// from the CPU only through exp()/log() rounding.  All arithmetic is float64.
#include <stdlib.h>
#include <type_traits>
#include "mma_common.cuh"
#include "dp_warp.cuh"

namespace segb {

__global__ void __launch_bounds__(128) dp_banded_kernel(DpParams p) {
    extern __shared__ double smem_alpha[];
    const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u_local = blockIdx.x * (blockDim.x >> 5) + warp_in_block;
    if (u_local >= p.n_utt) return;
    const int u = p.utt_first + u_local;
    double *al = smem_alpha + (size_t)warp_in_block * p.N_cap;
    const int64_t off = p.pos_off[u];
    const int N = (int)(p.pos_off[u + 1] - off);
    const double *sc = p.scores + (p.scores_local ? 0 : off * p.S);
    if (N <= 0) { if (lane == 0) { p.status[u_local] = SEGB_DP_OK; p.log_prob[u_local] = 0.0; } return; }
    const int64_t ubase = p.u_counter ? *p.u_counter : off;
    double total;
    int status, used;
    dp_warp_body(p, sc, p.bounds + off, N, al, p.alphas ? p.alphas + off : nullptr, ubase, total, status, used);
    if (lane == 0) {
        p.log_prob[u_local] = (status == SEGB_DP_OK) ? total : CUDART_NAN;
        p.status[u_local] = status;
        if (p.n_draws) p.n_draws[u_local] = used;
        if (p.u_counter) *p.u_counter = ubase + used;
    }
}

// ---------------------------------------------------------------------------------------
// Short-span fast path: one THREAD per utterance (band width <= 8, N <= DP_SMALL_N,
// n_slices_min <= 1).  With S = 6 a warp-per-utterance layout leaves 26 of 32 lanes idle and
// spends its time in shuffles; here the S candidates of a position live in registers.  The
// body is shared by two kernels that differ only in where rows and alphas live:
//   dp_small_kernel  -- rows straight from global memory, alphas in a per-thread local array
//                       (single utterances of the sequential sweeps, odd band widths);
//   dp_staged_kernel -- every utterance's score rows brought into shared memory by ONE
//                       cp.async.bulk (TMA) per utterance, alphas in shared memory: HBM is
//                       streamed by the copy engine in 1 KB bursts instead of being walked
//                       48 bytes at a time by each thread.
// Same operation order as the warp kernel / the reference, so results are identical.
constexpr int DP_SMALL_N = 64;
constexpr int DP_SMALL_S = 8;

struct GlobalRows {            // banded rows in global memory
    const double *sc; int S;
    __device__ __forceinline__ void load(int r, double *row) const {
#pragma unroll
        for (int l = 0; l < DP_SMALL_S; ++l) row[l] = (l < S) ? sc[(int64_t)r * S + l] : neg_inf();
    }
};
struct SharedRows {            // rows staged in shared memory, 16-byte aligned, S even
    const double *sc; int S;
    __device__ __forceinline__ void load(int r, double *row) const {
        const double2 *q = reinterpret_cast<const double2 *>(sc + r * S);
#pragma unroll
        for (int g = 0; g < DP_SMALL_S / 2; ++g) {
            if (2 * g < S) { const double2 v = q[g]; row[2 * g] = v.x; row[2 * g + 1] = v.y; }
            else { row[2 * g] = neg_inf(); row[2 * g + 1] = neg_inf(); }
        }
    }
};
struct LocalAlphas {           // per-thread array (local memory, L1-resident)
    double al[DP_SMALL_N];
    __device__ __forceinline__ double get(int j) const { return al[j]; }
    __device__ __forceinline__ void set(int j, double v) { al[j] = v; }
};
struct SharedAlphas {          // [N_cap][32] in shared memory: lane-interleaved, conflict-free
    double *al;
    __device__ __forceinline__ double get(int j) const { return al[j * 32]; }
    __device__ __forceinline__ void set(int j, double v) { al[j * 32] = v; }
};

struct DpThreadResult { unsigned long long bmask; double total; int status, used; };

template <int MODE, typename Rows, typename Alphas>
__device__ __forceinline__ DpThreadResult dp_backward(const DpParams &p, const Rows &rows, Alphas &A, int N,
                                                      int64_t ubase, int status);

template <int MODE, typename Rows, typename Alphas>
__device__ __forceinline__ DpThreadResult dp_thread_body(const DpParams &p, const Rows &rows, Alphas &A,
                                                         int N, int64_t ubase) {
    const int S = p.S;
    const int Wlim = (p.n_max == 0 || p.n_max > S) ? S : p.n_max;
    A.set(0, 0.0);
    int status = SEGB_DP_OK;
    double cur[DP_SMALL_S], nxt[DP_SMALL_S];
    rows.load(0, cur);                                                               // row of t = 1

    // c[l-1] = score(t, l) + alpha[t-l] for l = 1..W; returns max, flags NaN
    auto cands = [&](const double *row, int t, int W, double *c, bool &has_nan) {
        double m = neg_inf();
        has_nan = false;
#pragma unroll
        for (int l = 1; l <= DP_SMALL_S; ++l) {
            if (l <= W) {
                const double v = row[l - 1] + A.get(t - l);
                c[l - 1] = v;
                has_nan |= (v != v);
                m = fmax(m, v);
            } else c[l - 1] = neg_inf();
        }
        return m;
    };
    // logsumexp in the reference's element order: descending span
    auto lse_desc = [&](const double *c, int W, double m) {
        double s = 0.0;
#pragma unroll
        for (int l = DP_SMALL_S; l >= 1; --l) if (l <= W) s += exp(c[l - 1] - m);
        return log(s) + m;
    };

    for (int t = 1; t < N; ++t) {
        rows.load(t, nxt);                       // request the next row before reducing this one
        const int W = min(t, Wlim);
        double c[DP_SMALL_S];
        bool has_nan;
        const double m = cands(cur, t, W, c, has_nan);
        double a_t;
        if (has_nan) { status = SEGB_DP_NAN; a_t = neg_inf(); }
        else if (m == neg_inf()) a_t = neg_inf();
        else if (MODE == SEGB_DP_FFBS) a_t = lse_desc(c, W, m) + p.log_p_continue;
        else a_t = m;
        A.set(t, a_t);
#pragma unroll
        for (int l = 0; l < DP_SMALL_S; ++l) cur[l] = nxt[l];
    }

    return dp_backward<MODE>(p, rows, A, N, ubase, status);
}

// General backward pass (sampling / exp-normalised argmax / raw argmax) given all alphas.
template <int MODE, typename Rows, typename Alphas>
__device__ __forceinline__ DpThreadResult dp_backward(const DpParams &p, const Rows &rows, Alphas &A, int N,
                                                      int64_t ubase, int status) {
    const int S = p.S;
    const int Wlim = (p.n_max == 0 || p.n_max > S) ? S : p.n_max;
    auto cands = [&](const double *row, int t, int W, double *c, bool &has_nan) {
        double m = neg_inf();
        has_nan = false;
#pragma unroll
        for (int l = 1; l <= DP_SMALL_S; ++l) {
            if (l <= W) {
                const double v = row[l - 1] + A.get(t - l);
                c[l - 1] = v;
                has_nan |= (v != v);
                m = fmax(m, v);
            } else c[l - 1] = neg_inf();
        }
        return m;
    };
    auto lse_desc = [&](const double *c, int W, double m) {
        double s = 0.0;
#pragma unroll
        for (int l = DP_SMALL_S; l >= 1; --l) if (l <= W) s += exp(c[l - 1] - m);
        return log(s) + m;
    };
    unsigned long long bmask = 1ull << (N - 1);
    double total = 0.0;
    int used = 0;
    int t = N;
    for (int guard = 0; guard <= N && status == SEGB_DP_OK; ++guard) {
        int W = min(t, Wlim);
        double c[DP_SMALL_S], row[DP_SMALL_S];
        bool has_nan;
        rows.load(t - 1, row);
        double m = cands(row, t, W, c, has_nan);
        if (has_nan && MODE != SEGB_DP_VITERBI_GMM) { status = SEGB_DP_NAN; break; }
        if (m == neg_inf()) {
            while (m == neg_inf()) {
                t = t - 1;
                if (t == 0) break;
                W = min(t, Wlim);
                rows.load(t - 1, row);
                m = cands(row, t, W, c, has_nan);
            }
            if (t == 0) { status = SEGB_DP_INFEASIBLE; break; }
            bmask |= 1ull << (t - 1);
        }
        int idx = 0;
        if (MODE == SEGB_DP_VITERBI_KMEANS) {
            idx = W - 1;
#pragma unroll
            for (int l = DP_SMALL_S; l >= 1; --l) if (l <= W && c[l - 1] == m) idx = l - 1;   // first maximum
        } else {
            const double lse = lse_desc(c, W, m);
            double pr[DP_SMALL_S];
            if (MODE == SEGB_DP_FFBS && p.anneal_temp != 1.0) {
                const double inv_t = 1. / p.anneal_temp;
                double q[DP_SMALL_S], mq = neg_inf();
#pragma unroll
                for (int l = 1; l <= DP_SMALL_S; ++l) if (l <= W) { q[l - 1] = inv_t * (c[l - 1] - lse); mq = fmax(mq, q[l - 1]); }
                double s2 = 0.0;
#pragma unroll
                for (int l = 1; l <= DP_SMALL_S; ++l) if (l <= W) s2 += exp(q[l - 1] - mq);      // ascending span
                const double lse2 = log(s2) + mq;
#pragma unroll
                for (int l = 1; l <= DP_SMALL_S; ++l) pr[l - 1] = (l <= W) ? exp(q[l - 1] - lse2) : 0.0;
            } else {
#pragma unroll
                for (int l = 1; l <= DP_SMALL_S; ++l) pr[l - 1] = (l <= W) ? exp(c[l - 1] - lse) : 0.0;
            }
            if (MODE == SEGB_DP_FFBS) {
                double uu = p.uniforms[ubase + used];
                used++;
                idx = W - 1;
                bool done = false;
#pragma unroll
                for (int l = 1; l <= DP_SMALL_S; ++l)
                    if (l <= W && !done) { uu = uu - pr[l - 1]; if (uu < 0) { idx = l - 1; done = true; } }
            } else {
                double pm = -1.0;
#pragma unroll
                for (int l = 1; l <= DP_SMALL_S; ++l) if (l <= W) pm = fmax(pm, pr[l - 1]);
                idx = 0;
                bool found = false;
#pragma unroll
                for (int l = 1; l <= DP_SMALL_S; ++l) if (l <= W && !found && pr[l - 1] == pm) { idx = l - 1; found = true; }
            }
        }
        const int k = idx + 1;
        total += row[k - 1];
        if (t - k - 1 < 0) break;
        bmask |= 1ull << (t - k - 1);
        t = t - k;
    }
    DpThreadResult r;
    r.bmask = bmask; r.total = total; r.status = status; r.used = used;
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(128) dp_small_kernel(DpParams p) {
    const int u_local = blockIdx.x * blockDim.x + threadIdx.x;
    if (u_local >= p.n_utt) return;
    const int u = p.utt_first + u_local;
    const int64_t off = p.pos_off[u];
    const int N = (int)(p.pos_off[u + 1] - off);
    uint8_t *bo = p.bounds + off;
    if (N <= 0) { p.status[u_local] = SEGB_DP_OK; p.log_prob[u_local] = 0.0; return; }
    GlobalRows rows;
    rows.sc = p.scores + (p.scores_local ? 0 : off * p.S); rows.S = p.S;
    LocalAlphas A;
    const int64_t ubase = p.u_counter ? *p.u_counter : off;
    const DpThreadResult r = dp_thread_body<MODE>(p, rows, A, N, ubase);
    if (p.alphas) for (int j = 0; j < N; ++j) p.alphas[off + j] = A.get(j);
    for (int j = 0; j < N; ++j) bo[j] = (uint8_t)((r.bmask >> j) & 1ull);
    p.log_prob[u_local] = (r.status == SEGB_DP_OK) ? r.total : CUDART_NAN;
    p.status[u_local] = r.status;
    if (p.n_draws) p.n_draws[u_local] = r.used;
    if (p.u_counter) *p.u_counter = ubase + r.used;
}

// Staged variant: one warp (= one block) per 32 consecutive utterances, band width SB known at
// compile time, window limit == band (the common case: n_slices_max == S).  Lane i issues ONE
// cp.async.bulk for the whole banded score block of utterance i (N_i * SB * 8 bytes, contiguous
// in HBM) into its shared-memory slot; slots are G granules (16 B) apart with G == 1 (mod 8), so
// the 128-bit row loads of a quarter-warp, which run in lock-step over t, hit eight different
// bank groups.  The forward pass keeps the last SB alphas in a register ring (the t loop is
// unrolled SB times so ring indices are compile-time); k-means Viterbi records the winning span
// of every position as a back-pointer byte and the backward pass only chases pointers, the
// other two modes keep their alphas in shared memory and run the general backward pass.
// Boundaries are assembled in shared memory and written back as one contiguous run of bytes.
// Shared memory of one warp: the group's score rows (tightly packed, one bulk copy), SB rows of
// slack for the row ring of the skewed forward pass, back-pointer bytes, boundary bytes.
__host__ __device__ inline size_t dp_staged_smem_bytes(int N_cap, int S, int G, bool alphas_in_smem) {
    return 128 + (alphas_in_smem ? (size_t)32 * N_cap * 8 : 0) + (size_t)32 * (N_cap + 1) /* back-pointers */
           + (size_t)((G * N_cap + 4 + 15) / 16) * 16 /* boundary bytes (+ alignment phase) */
           + (size_t)G * N_cap * S * 8 + (size_t)(S + 1) * S * 8;
}
// Utterances per warp: the choice that keeps most utterances resident per SM (228 KB of shared
// memory, 1 KB reserved per block), ties to the wider group.
inline int dp_staged_pick_group(int N_cap, int S, bool alphas_in_smem) {
    int best = 0, best_res = 0;
    for (int G = 32; G >= 16; --G) {
        const size_t b = dp_staged_smem_bytes(N_cap, S, G, alphas_in_smem) + 1024;
        const int per_sm = (int)min((size_t)32, (size_t)(228 * 1024) / b);
        if (G * per_sm > best_res) { best_res = G * per_sm; best = G; }
    }
    return best;
}

template <int J, int E, typename F>
__device__ __forceinline__ bool static_for_while(F &&f) {
    if constexpr (J < E) {
        if (!f(std::integral_constant<int, J>{})) return false;
        return static_for_while<J + 1, E>(f);
    }
    return true;
}

// First maximum of c[LO..HI) as a balanced tree: the right half wins only when strictly
// greater, so ties resolve to the lowest index exactly like a left-to-right scan.
template <int LO, int HI, int W>
__device__ __forceinline__ void tree_first_max(const double (&c)[W], double &m, int &idx) {
    if constexpr (HI - LO == 1) { m = c[LO]; idx = LO; }
    else {
        constexpr int MID = LO + (HI - LO + 1) / 2;
        double ml, mr;
        int il, ir;
        tree_first_max<LO, MID, W>(c, ml, il);
        tree_first_max<MID, HI, W>(c, mr, ir);
        const bool right = mr > ml;
        m = right ? mr : ml;
        idx = right ? ir : il;
    }
}

// One forward step at position t with t % SB == J (FIRST: t == J < SB, spans limited to l <= t).
// w[] is the alpha ring: alpha[t'] lives in w[t' % SB].  `live` is false for lanes whose
// utterance ended before t: they keep executing (the warp's trip count is uniform, no
// divergence) on whatever their slot holds, and only their NaN flag is masked.
template <int MODE, int SB, int J, bool FIRST>
__device__ __forceinline__ void dp_fwd_step(const double *row_ptr, double (&w)[SB], bool live, bool &bad, uint8_t *bp,
                                            double *al, double log_p_continue) {
    double row[SB];
    const double2 *q = reinterpret_cast<const double2 *>(row_ptr);
#pragma unroll
    for (int g = 0; g < SB / 2; ++g) { const double2 v = q[g]; row[2 * g] = v.x; row[2 * g + 1] = v.y; }
    constexpr int W = FIRST ? J : SB;
    double c[W];
    bool nan = false;
#pragma unroll
    for (int l = 1; l <= W; ++l) {
        c[l - 1] = row[l - 1] + w[(J - l + 2 * SB) % SB];
        nan |= (c[l - 1] != c[l - 1]);
    }
    bad |= (nan && live);
    // first maximum in ascending span order.  c[0] depends on the alpha formed one step ago and
    // is compared last, so the recurrence's critical path is one add + one compare-select.
    double m = c[0];
    int idx = 0;
    if constexpr (W >= 2) {
        double mr;
        int ir;
        tree_first_max<1, W, W>(c, mr, ir);
        if (mr > m) { m = mr; idx = ir; }
    }
    double a_t = m;
    if (MODE == SEGB_DP_FFBS && m != neg_inf()) {
        double s = 0.0;
#pragma unroll
        for (int l = W; l >= 1; --l) s += exp(c[l - 1] - m);    // descending span = the reference's order
        a_t = log(s) + m + log_p_continue;
    }
    if (MODE == SEGB_DP_VITERBI_KMEANS) *bp = (m == neg_inf()) ? (uint8_t)0xff : (uint8_t)idx;
    else *al = a_t;
    w[J % SB] = a_t;
}

// Skewed ("systolic") forward pass for the two Viterbi modes.  At time tau the alpha of
// position tau has just been formed; it is a candidate of the SB positions t = tau + l
// (l = 1..SB), so every step issues SB independent add + compare-select pairs, one for each
// of the next SB positions, and only the l = 1 pair is on the recurrence's critical path.
// Per position the candidates therefore arrive in DESCENDING span order; `keep the running
// maximum only if strictly greater` resolves ties to the smallest span, exactly like the
// reference's first-maximum over the reversed window.
//   R[t % SB]  score row of position t (ring of SB rows in registers)
//   m/ix/nf[t % SB]  running maximum, its span index, NaN flag of position t
template <int SB>
struct DpSkewState {
    double R[SB][SB];
    double m[SB];
    int ix[SB];
    bool nf[SB];
    double a;
};

template <int SB>
__device__ __forceinline__ void dp_load_row(const double *row_ptr, double (&row)[SB]) {
    const double2 *q = reinterpret_cast<const double2 *>(row_ptr);
#pragma unroll
    for (int g = 0; g < SB / 2; ++g) { const double2 v = q[g]; row[2 * g] = v.x; row[2 * g + 1] = v.y; }
}

// tau % SB == J; FIRST: tau == 0 (every candidate opens its position).
// NANCHK = false: the caller vouches for scores without NaN / +inf (SEGB_DP_SCORES_FINITE), so no candidate
// can be NaN: the per-candidate NaN compares go (3 of the step's 10 DSETPs, the FP64 pipe's slowest instruction),
// and the -inf test of the new alpha becomes an integer compare of its high word.
template <int MODE, int SB, int J, bool FIRST, bool NANCHK>
__device__ __forceinline__ void dp_skew_step(DpSkewState<SB> &st, const double *sc, int tau, int Nf, bool &bad,
                                             bool &any_dead, uint8_t *bp, double *al) {
#pragma unroll
    for (int l = 1; l <= SB; ++l) {
        constexpr int dummy = 0; (void)dummy;
        const int r = (J + l) % SB;
        const double c = st.R[r][l - 1] + st.a;
        const bool nanc = NANCHK ? (c != c) : false;
        if (FIRST || l == SB) { st.m[r] = c; st.ix[r] = l - 1; if (NANCHK) st.nf[r] = nanc; }
        else {
            const bool keep = st.m[r] > c;
            st.m[r] = keep ? st.m[r] : c;
            st.ix[r] = keep ? st.ix[r] : l - 1;
            if (NANCHK) st.nf[r] |= nanc;
        }
    }
    constexpr int r1 = (J + 1) % SB;
    const double a_new = st.m[r1];
    if (NANCHK) bad |= (st.nf[r1] && (tau + 1 <= Nf));
    const bool dead = NANCHK ? (a_new == neg_inf()) : (__double2hiint(a_new) == (int)0xFFF00000);
    any_dead |= dead;                    // some window was all -inf: the backward pass needs its walk-left branch
    if (MODE == SEGB_DP_VITERBI_KMEANS) bp[(tau + 1) * 32] = dead ? (uint8_t)0xff : (uint8_t)st.ix[r1];
    else al[(tau + 1) * 32] = a_new;
    st.a = a_new;
    dp_load_row<SB>(sc + (tau + SB) * SB, st.R[r1]);          // row of position tau + 1 + SB
}

template <int MODE, int SB, bool NANCHK>
__device__ __forceinline__ void dp_forward_skewed(const double *sc, int Nf, int Nf_w, bool &bad, bool &any_dead, uint8_t *bp,
                                                  double *al) {
    DpSkewState<SB> st;
    static_for_while<1, SB + 1>([&](auto tc) {
        constexpr int T = decltype(tc)::value;
        dp_load_row<SB>(sc + (T - 1) * SB, st.R[T % SB]);
        return true;
    });
    st.a = 0.0;
    if (Nf_w < 1) return;
    dp_skew_step<MODE, SB, 0, true, NANCHK>(st, sc, 0, Nf, bad, any_dead, bp, al);
    static_for_while<1, SB>([&](auto jc) {
        constexpr int J = decltype(jc)::value;
        if (J >= Nf_w) return false;
        dp_skew_step<MODE, SB, J, false, NANCHK>(st, sc, J, Nf, bad, any_dead, bp, al);
        return true;
    });
    int tau0 = SB;
    for (; tau0 + SB <= Nf_w; tau0 += SB) {                   // SB steps, straight-line
        static_for_while<0, SB>([&](auto jc) {
            constexpr int J = decltype(jc)::value;
            dp_skew_step<MODE, SB, J, false, NANCHK>(st, sc, tau0 + J, Nf, bad, any_dead, bp, al);
            return true;
        });
    }
    static_for_while<0, SB>([&](auto jc) {
        constexpr int J = decltype(jc)::value;
        if (tau0 + J >= Nf_w) return false;
        dp_skew_step<MODE, SB, J, false, NANCHK>(st, sc, tau0 + J, Nf, bad, any_dead, bp, al);
        return true;
    });
}

template <int MODE, int SB, bool NANCHK = true>
__global__ void __launch_bounds__(32) dp_staged_kernel(DpParams p) {
    extern __shared__ __align__(128) unsigned char dp_sm[];
    constexpr bool AL_SMEM = (MODE != SEGB_DP_VITERBI_KMEANS);
    const int lane = threadIdx.x;
    const int N_cap = p.N_cap, G = p.group;
    double *al_s = reinterpret_cast<double *>(dp_sm + 128);
    uint8_t *bp_s = dp_sm + 128 + (AL_SMEM ? (size_t)32 * N_cap * 8 : 0);
    uint8_t *bo_s = bp_s + (size_t)32 * (N_cap + 1);
    const double *sc_s = reinterpret_cast<const double *>(bo_s + (size_t)((G * N_cap + 4 + 15) / 16) * 16);
    const uint32_t bar = mma::smem_u32(dp_sm);
    uint8_t *bp = bp_s + lane;
    double *al = al_s + lane;

    if (lane == 0) {
        mma::mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    // persistent warp: a contiguous range of utterances, worked off in groups of consecutive utterances
    // (lane = utterance); the next group's offsets are requested before this group's copy is awaited and
    // consumed after its compute.  The group's score rows are contiguous in HBM: ONE bulk copy.
    // A group takes as many utterances (<= 32) as fit the staging buffers (G * N_cap landmarks): the buffers
    // are sized for G utterances of the LONGEST length, so with average-length utterances all 32 lanes work
    // (round 1 always took G = 26 of them: 8 rounds per warp at 200k utterances instead of 7).
    const int u_lo = (int)((long long)blockIdx.x * p.n_utt / gridDim.x);
    const int u_hi = (int)((long long)(blockIdx.x + 1) * p.n_utt / gridDim.x);
    // every lane runs to the group's longest utterance and prefetches SB rows ahead, i.e. it reads up to N_cap + SB rows
    // past ITS OWN start: the group's rows must end N_cap rows before the end of the staging buffer (G * N_cap + S + 1)
    const int budget = (G - 1) * N_cap;
    auto group_offsets = [&](int u_start, int64_t &o0, int64_t &o1, bool &valid) {
        const int idx = u_start + lane;
        const bool in_range = idx < u_hi;
        const int u = p.utt_first + (in_range ? idx : u_hi - 1);      // out-of-range lanes re-read the last utterance
        o0 = p.pos_off[u];
        o1 = p.pos_off[u + 1];
        const int64_t base = __shfl_sync(FULL, o0, 0);
        valid = in_range && (o1 - base) <= budget;                     // a prefix of the lanes (offsets ascend)
    };
    int64_t off = 0, o1 = 0, off_next = 0, o1_next = 0;
    bool valid = false, valid_next = false;
    int u_start = u_lo;
    if (u_start < u_hi) group_offsets(u_start, off, o1, valid);
    int N = valid ? (int)(o1 - off) : 0;
    uint32_t parity = 0;
    while (u_start < u_hi) {
        const int n_here = __popc(__ballot_sync(FULL, valid));        // >= 1: one utterance always fits
        const int u_next = u_start + n_here;
        const int64_t off0 = __shfl_sync(FULL, off, 0);
        int64_t end = valid ? off + N : 0;
        // uniform trip count: every lane runs to the longest utterance of the group
        const int Nf = (MODE == SEGB_DP_VITERBI_KMEANS) ? N : N - 1;
        int Nf_w = Nf;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Nf_w = max(Nf_w, __shfl_xor_sync(FULL, Nf_w, o));
            end = max(end, __shfl_xor_sync(FULL, end, o));
        }
        const int n_bytes = (int)(end - off0);                  // landmarks of the group
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)n_bytes * SB * 8;
            if (bytes) {
                mma::mbar_expect_tx(bar, bytes);
                mma::bulk_g2s(mma::smem_u32(sc_s), p.scores + off0 * SB, bytes, bar);
            } else {
                mma::mbar_arrive(bar);
            }
        }
        valid_next = false;
        if (u_next < u_hi) group_offsets(u_next, off_next, o1_next, valid_next);
        const double *sc = sc_s + (valid ? (off - off0) * SB : 0);
        const int u_local = u_start + lane;
        // boundary bytes are staged with the same 4-byte phase as their destination in HBM
        uint8_t *bo = p.bounds + off0;
        const int phase = (int)((uintptr_t)bo & 3);
        uint8_t *bs0 = bo_s + phase;
        for (int wd = lane; wd * 4 < phase + n_bytes; wd += 32) reinterpret_cast<uint32_t *>(bo_s)[wd] = 0u;
        bool bad = false, any_dead = false;
        if (AL_SMEM) al[0] = 0.0;
        mma::mbar_wait(bar, parity);
        parity ^= 1;

        if (MODE == SEGB_DP_FFBS) {
            double w[SB];
            w[0] = 0.0;
#pragma unroll
            for (int i = 1; i < SB; ++i) w[i] = neg_inf();
            static_for_while<1, SB>([&](auto jc) {
                constexpr int J = decltype(jc)::value;
                if (J > Nf_w) return false;
                dp_fwd_step<MODE, SB, J, true>(sc + (J - 1) * SB, w, J <= Nf, bad, bp + J * 32, al + J * 32, p.log_p_continue);
                return true;
            });
            for (int t0 = SB; t0 <= Nf_w; t0 += SB) {
                static_for_while<0, SB>([&](auto jc) {
                    constexpr int J = decltype(jc)::value;
                    const int t = t0 + J;
                    if (t > Nf_w) return false;
                    dp_fwd_step<MODE, SB, J, false>(sc + (t - 1) * SB, w, t <= Nf, bad, bp + t * 32, al + t * 32, p.log_p_continue);
                    return true;
                });
            }
        } else {
            dp_forward_skewed<MODE, SB, NANCHK>(sc, Nf, Nf_w, bad, any_dead, bp, al);
        }

        int status = bad ? SEGB_DP_NAN : SEGB_DP_OK, used = 0;
        double total = 0.0;
        __syncwarp();                                       // zero-fill of the boundary bytes is complete
        if (N > 0) {
            if (AL_SMEM && p.alphas) for (int j = 0; j < N; ++j) p.alphas[off + j] = al[j * 32];
            // ---------------- backward
            uint8_t *bs = bs0 + (off - off0);
            if (MODE == SEGB_DP_VITERBI_KMEANS) {
                // pointer chase (t strictly decreases, so no trip guard); the chosen score of
                // one hop is added while the next hop's back-pointer is in flight -- same
                // summation order as the reference: right to left
                bs[N - 1] = 1;
                int t = N;
                double pend = 0.0;
                if (status == SEGB_DP_OK && !any_dead) {
                    // every window had a finite candidate: no walk-left branch in the chase
                    while (true) {
                        const int b = bp[t * 32];
                        total += pend;
                        pend = sc[(t - 1) * SB + b];
                        t = t - b - 1;
                        if (t <= 0) break;
                        bs[t - 1] = 1;
                    }
                    total += pend;
                } else if (status == SEGB_DP_OK) {
                    while (true) {
                        int b = bp[t * 32];
                        total += pend;
                        if (b == 0xff) {                   // window all -inf: walk left until feasible
                            while (b == 0xff) {
                                t = t - 1;
                                if (t == 0) break;
                                b = bp[t * 32];
                            }
                            if (t == 0) { status = SEGB_DP_INFEASIBLE; pend = 0.0; break; }
                            bs[t - 1] = 1;
                        }
                        pend = sc[(t - 1) * SB + b];
                        t = t - b - 1;
                        if (t <= 0) break;
                        bs[t - 1] = 1;
                    }
                    total += pend;
                }
            } else {
                SharedRows rows;
                rows.sc = sc; rows.S = SB;
                SharedAlphas A;
                A.al = al;
                const DpThreadResult r = dp_backward<MODE>(p, rows, A, N, off, status);
                total = r.total; status = r.status; used = r.used;
                for (int j = 0; j < N; ++j) bs[j] = (uint8_t)((r.bmask >> j) & 1ull);
            }
        }
        if (valid) {
            p.log_prob[u_local] = (N <= 0) ? 0.0 : ((status == SEGB_DP_OK) ? total : CUDART_NAN);
            p.status[u_local] = (N <= 0) ? SEGB_DP_OK : status;
            if (p.n_draws) p.n_draws[u_local] = used;
        }
        __syncwarp();
        // the group's boundaries are one contiguous run of bytes in HBM: whole aligned words,
        // plus the ragged head / tail bytes
        {
            const int lo = phase, hi = phase + n_bytes;           // byte range inside bo_s
            const int w_lo = (lo + 3) >> 2, w_hi = hi >> 2;       // full words [w_lo, w_hi)
            uint8_t *g0 = bo - phase;                             // 4-byte aligned
            if (w_hi > w_lo) {
                for (int wd = w_lo + lane; wd < w_hi; wd += 32)
                    reinterpret_cast<uint32_t *>(g0)[wd] = reinterpret_cast<const uint32_t *>(bo_s)[wd];
                if (lo + lane < 4 * w_lo) g0[lo + lane] = bo_s[lo + lane];
                if (4 * w_hi + lane < hi) g0[4 * w_hi + lane] = bo_s[4 * w_hi + lane];
            } else {
                for (int b = lo + lane; b < hi; b += 32) g0[b] = bo_s[b];
            }
        }
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // slots are rewritten by the copy engine
        u_start = u_next;
        off = off_next;
        valid = valid_next;
        N = valid ? (int)(o1_next - off_next) : 0;
    }
}

int launch_dp(const segb_corpus *c, int32_t utt_first, int32_t n_utt, const double *scores, int32_t mode,
              double log_p_continue, double anneal_temp, const double *uniforms, int64_t *u_counter,
              uint8_t *bounds_out, double *log_prob, double *alphas, int32_t *n_draws, int32_t *status,
              cudaStream_t stream, int scores_local = 0) {
    const bool scores_finite = (mode & SEGB_DP_SCORES_FINITE) != 0;
    mode &= ~SEGB_DP_SCORES_FINITE;
    DpParams p;
    p.scores_local = scores_local;
    p.group = 32;
    p.pos_off = c->pos_off; p.scores = scores; p.uniforms = uniforms; p.u_counter = u_counter;
    p.bounds = bounds_out; p.log_prob = log_prob; p.alphas = alphas; p.n_draws = n_draws; p.status = status;
    p.utt_first = utt_first; p.n_utt = n_utt; p.S = c->S; p.n_min = c->n_slices_min; p.n_max = c->n_slices_max;
    p.mode = mode; p.N_cap = c->N_max; p.log_p_continue = log_p_continue; p.anneal_temp = anneal_temp;
    if (c->S <= DP_SMALL_S && c->N_max <= DP_SMALL_N && c->n_slices_min <= 1) {
        // batched calls: stage the score blocks through shared memory with the copy engine
        const int Wlim = (c->n_slices_max == 0 || c->n_slices_max > c->S) ? c->S : c->n_slices_max;
        const bool al_smem = (mode != SEGB_DP_VITERBI_KMEANS);
        int G = dp_staged_pick_group(c->N_max, c->S, al_smem);
        if (const char *env = getenv("SEGB_DP_GROUP")) {            // tuning aid
            const int g_env = atoi(env);
            if (g_env >= 16 && g_env <= 32) G = g_env;
        }
        const size_t st_smem = G ? dp_staged_smem_bytes(c->N_max, c->S, G, al_smem) : 0;
        if (!scores_local && !u_counter && n_utt >= 32 && (c->S % 2) == 0 && Wlim == c->S && G >= 16 &&
            ((uintptr_t)scores & 15) == 0 && st_smem <= 72 * 1024 && !(alphas && !al_smem)) {
            p.group = G;
            const int groups = (n_utt + 31) / 32;               // warps worth launching: each takes a contiguous range
            auto launch = [&](auto kern) -> int {
                // resident warps per device for this (kernel, shared-memory size, device): queried when
                // any of the three changes (the shared-memory opt-in is a per-device function attribute);
                // thread_local: concurrent host threads each keep their own small cache
                static thread_local void *cached_kern = nullptr;
                static thread_local size_t cached_smem = 0;
                static thread_local int cached_resident = 0, cached_dev = -1;
                int dev = 0, n_sm = 0;
                { const int rc = device_info(&dev, &n_sm, nullptr); if (rc) return rc; }
                if (cached_kern != (void *)kern || cached_smem != st_smem || cached_dev != dev) {
                    if (st_smem > 48 * 1024)
                        SEGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st_smem));
                    int per_sm = 0;
                    SEGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, st_smem));
                    cached_kern = (void *)kern; cached_smem = st_smem; cached_dev = dev;
                    cached_resident = max(1, per_sm) * n_sm;
                }
                const int blocks = min(groups, cached_resident);            // persistent warps
                kern<<<blocks, 32, st_smem, stream>>>(p);
                SEGB_LAUNCH_CHECK();
                return 0;
            };
            auto by_band = [&](auto mc) -> int {
                constexpr int M = decltype(mc)::value;
                if (M == SEGB_DP_VITERBI_KMEANS && scores_finite && c->S == 6)      // the frozen k-means sweep's shape
                    return launch(dp_staged_kernel<SEGB_DP_VITERBI_KMEANS, 6, false>);
                switch (c->S) {
                    case 2: return launch(dp_staged_kernel<M, 2>);
                    case 4: return launch(dp_staged_kernel<M, 4>);
                    case 6: return launch(dp_staged_kernel<M, 6>);
                    default: return launch(dp_staged_kernel<M, 8>);
                }
            };
            if (mode == SEGB_DP_FFBS) return by_band(std::integral_constant<int, SEGB_DP_FFBS>{});
            if (mode == SEGB_DP_VITERBI_GMM) return by_band(std::integral_constant<int, SEGB_DP_VITERBI_GMM>{});
            return by_band(std::integral_constant<int, SEGB_DP_VITERBI_KMEANS>{});
        }
        const int threads = 128, blocks = (n_utt + threads - 1) / threads;
        if (mode == SEGB_DP_FFBS) dp_small_kernel<SEGB_DP_FFBS><<<blocks, threads, 0, stream>>>(p);
        else if (mode == SEGB_DP_VITERBI_GMM) dp_small_kernel<SEGB_DP_VITERBI_GMM><<<blocks, threads, 0, stream>>>(p);
        else dp_small_kernel<SEGB_DP_VITERBI_KMEANS><<<blocks, threads, 0, stream>>>(p);
        SEGB_LAUNCH_CHECK();
        return 0;
    }
    const int warps = 4;
    const size_t smem = (size_t)warps * c->N_max * sizeof(double);
    if (smem > 200 * 1024) { set_error("utterance too long for the DP kernel (N_max=%d)", c->N_max); return SEGB_E_UNSUPPORTED; }
    if (smem > 48 * 1024)
        SEGB_CUDA(cudaFuncSetAttribute(dp_banded_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (n_utt + warps - 1) / warps;
    dp_banded_kernel<<<blocks, warps * 32, smem, stream>>>(p);
    SEGB_LAUNCH_CHECK();
    return 0;
}

// Single utterance, scores indexed from the utterance's first slot (sequential sweeps).
int launch_dp_local(const segb_corpus *c, int32_t utt, const double *local_scores, int32_t mode,
                    double log_p_continue, double anneal_temp, const double *uniforms, int64_t *u_counter,
                    double *log_prob, int32_t *status, cudaStream_t stream) {
    return launch_dp(c, utt, 1, local_scores, mode, log_p_continue, anneal_temp, uniforms, u_counter,
                     c->bounds, log_prob, nullptr, nullptr, status, stream, 1);
}

}  // namespace segb

extern "C" int segb_dp_banded(const segb_corpus *c, int32_t utt_first, int32_t n_utt, const double *scores,
                              int32_t mode, double log_p_continue, double anneal_temp,
                              const double *uniforms, int64_t *u_counter, uint8_t *bounds_out,
                              double *log_prob, double *alphas, int32_t *n_draws, int32_t *status,
                              void *stream) {
    SEGB_CHECK_ARG(c && scores && bounds_out && log_prob && status, "null pointer");
    SEGB_CHECK_ARG((mode & ~SEGB_DP_SCORES_FINITE) >= 0 && (mode & ~SEGB_DP_SCORES_FINITE) <= 2, "mode");
    SEGB_CHECK_ARG(n_utt >= 0 && utt_first >= 0 && utt_first + n_utt <= c->n_utt, "utterance range");
    SEGB_CHECK_ARG(c->S >= 1, "band width");
    SEGB_CHECK_ARG((mode & ~SEGB_DP_SCORES_FINITE) != SEGB_DP_FFBS || uniforms, "FFBS needs uniforms");
    SEGB_CHECK_ARG(!u_counter || n_utt <= 1, "u_counter implies a single utterance");
    if (n_utt == 0) return 0;
    return segb::launch_dp(c, utt_first, n_utt, scores, mode, log_p_continue, anneal_temp, uniforms,
                           u_counter, bounds_out, log_prob, alphas, n_draws, status, (cudaStream_t)stream);
}
