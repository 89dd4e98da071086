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
#include "common.cuh"

namespace segb {

struct DpParams {
    const int64_t *pos_off;
    const double *scores;
    const double *uniforms;
    int64_t *u_counter;
    uint8_t *bounds;
    double *log_prob;
    double *alphas;
    int32_t *n_draws;
    int32_t *status;
    int32_t utt_first, n_utt, S, n_min, n_max, mode, N_cap, scores_local;
    double log_p_continue, anneal_temp;
};

// Candidates of position t live in lanes/iterations idx = 0..W-1 with span l = idx+1.
// The reference orders a window by ascending start j = t - l, i.e. DESCENDING span.

// logsumexp over spans l in [l_lo, W] of c(l), summed in descending-span order
// (= the reference's array order, _cython_utils.pyx:13-25).  `cval(l)` returns the
// candidate held by the calling lane for span l = chunk*32 + lane + 1.
template <typename F>
__device__ double warp_lse_desc(F cval, int l_lo, int W, double m) {
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    const int n_chunks = (W + 31) >> 5;
    for (int ch = n_chunks - 1; ch >= 0; --ch) {
        const int l = ch * 32 + lane + 1;
        const double e = (l >= l_lo && l <= W) ? exp(cval(l) - m) : 0.0;
        const int hi = min(W - ch * 32, 32);         // lanes [0, hi) hold spans of this chunk
        for (int i = hi - 1; i >= 0; --i) {
            const double ei = __shfl_sync(FULL, e, i);
            if (ch * 32 + i + 1 >= l_lo) s += ei;
        }
    }
    return log(s) + m;
}

__global__ void __launch_bounds__(128) dp_banded_kernel(DpParams p) {
    extern __shared__ double smem_alpha[];
    const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u_local = blockIdx.x * (blockDim.x >> 5) + warp_in_block;
    if (u_local >= p.n_utt) return;
    const int u = p.utt_first + u_local;
    double *al = smem_alpha + (size_t)warp_in_block * p.N_cap;

    const int64_t off = p.pos_off[u];
    const int N = (int)(p.pos_off[u + 1] - off);
    const int S = p.S;
    const int Wlim = (p.n_max == 0 || p.n_max > S) ? S : p.n_max;   // window limit in spans
    const int n_min = p.n_min;
    const int l_cut = n_min > 1 ? n_min : 1;                        // [-S : -(n_min-1)] keeps spans >= n_min
    const double *sc = p.scores + (p.scores_local ? 0 : off * S);
    uint8_t *bo = p.bounds + off;
    int status = SEGB_DP_OK;

    if (N <= 0) { if (lane == 0) { p.status[u_local] = SEGB_DP_OK; p.log_prob[u_local] = 0.0; } return; }
    for (int j = lane; j < N; j += 32) bo[j] = (j == N - 1);
    if (lane == 0) al[0] = 0.0;
    __syncwarp();

    // score of the candidate with span l ending at t, plus alpha[t - l]
    auto cand = [&](int t, int l) -> double { return sc[(int64_t)(t - 1) * S + (l - 1)] + al[t - l]; };

    // ---------------- forward pass (t = 1 .. N-1)
    for (int t = 1; t < N && status == SEGB_DP_OK; ++t) {
        const int W = min(t, Wlim);
        double m_all = neg_inf(), m_cut = neg_inf();
        bool has_nan = false;
        for (int l = lane + 1; l <= W; l += 32) {
            const double c = cand(t, l);
            has_nan |= (c != c);
            m_all = fmax(m_all, c);
            if (l >= l_cut) m_cut = fmax(m_cut, c);
        }
        m_all = warp_max(m_all);
        m_cut = warp_max(m_cut);
        has_nan = __any_sync(FULL, has_nan);
        double a_t;
        if (has_nan) { status = SEGB_DP_NAN; a_t = neg_inf(); }
        else if (m_all == neg_inf()) a_t = neg_inf();
        else if (W < l_cut) { status = SEGB_DP_EMPTY_SLICE; a_t = neg_inf(); }
        else if (p.mode == SEGB_DP_FFBS) {
            if (m_cut == neg_inf()) a_t = CUDART_NAN;   // reference: exp(-inf - -inf) -> nan
            else a_t = warp_lse_desc([&](int l) { return cand(t, l); }, l_cut, W, m_cut) + p.log_p_continue;
        } else a_t = m_cut;
        __syncwarp();
        if (lane == 0) al[t] = a_t;
        __syncwarp();
    }
    if (p.alphas) for (int j = lane; j < N; j += 32) p.alphas[off + j] = al[j];

    // ---------------- backward pass
    double total = 0.0;
    int used = 0;
    int64_t ubase = p.u_counter ? *p.u_counter : off;
    int t = N;
    for (int guard = 0; guard <= N && status == SEGB_DP_OK; ++guard) {
        int W = min(t, Wlim);
        int l_lo = l_cut;                      // current window keeps spans >= l_lo
        if (W < l_lo) { status = SEGB_DP_EMPTY_SLICE; break; }
        // all -inf over the (cut) window?
        auto window_max = [&](int tt, int ww, int ll, bool &nan_seen) {
            double m = neg_inf();
            bool hn = false;
            for (int l = lane + 1; l <= ww; l += 32)
                if (l >= ll) { const double c = cand(tt, l); hn |= (c != c); m = fmax(m, c); }
            nan_seen = __any_sync(FULL, hn);
            return warp_max(m);
        };
        bool nan_seen;
        double m = window_max(t, W, l_lo, nan_seen);
        if (nan_seen && p.mode != SEGB_DP_VITERBI_GMM) { status = SEGB_DP_NAN; break; }
        if (m == neg_inf()) {
            // walk left until some candidate is feasible; the recomputed window is not
            // trimmed by n_slices_min (unigram_acoustic_wordseg.py:723-728)
            while (m == neg_inf()) {
                t = t - 1;
                if (t == 0) break;
                W = min(t, Wlim);
                l_lo = 1;
                m = window_max(t, W, l_lo, nan_seen);
            }
            if (t == 0) { status = SEGB_DP_INFEASIBLE; break; }
            if (lane == 0) bo[t - 1] = 1;
        }
        // choose the span index (0-based position in the reversed window: idx 0 <-> span l_lo)
        int idx;
        const int n_w = W - l_lo + 1;
        if (p.mode == SEGB_DP_VITERBI_KMEANS) {
            // first maximum of the reversed raw scores (kmeans_acoustic_wordseg.py:535-536)
            int best = 0x7fffffff;
            for (int l = l_lo + lane; l <= W; l += 32)
                if (cand(t, l) == m) { best = l - l_lo; break; }
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(FULL, best, o));
            idx = best;
        } else {
            const double lse = warp_lse_desc([&](int l) { return cand(t, l); }, l_lo, W, m);
            double lse2 = 0.0, inv_t = 1.0;
            const bool anneal = (p.mode == SEGB_DP_FFBS && p.anneal_temp != 1.0);
            if (anneal) {
                // log_p_k_anneal = 1/T * lp - logsumexp(1/T * lp), lp already reversed
                // (unigram_acoustic_wordseg.py:731-736): sum runs in ASCENDING span order
                inv_t = 1. / p.anneal_temp;
                double mq = neg_inf();
                for (int l = l_lo + lane; l <= W; l += 32) mq = fmax(mq, inv_t * (cand(t, l) - lse));
                mq = warp_max(mq);
                double s = 0.0;
                for (int base = l_lo; base <= W; base += 32) {
                    const int l = base + lane;
                    const double e = (l <= W) ? exp(inv_t * (cand(t, l) - lse) - mq) : 0.0;
                    const int hi = min(W - base + 1, 32);
                    for (int i = 0; i < hi; ++i) s += __shfl_sync(FULL, e, i);
                }
                lse2 = log(s) + mq;
            }
            auto prob = [&](int l) -> double {
                const double lp = cand(t, l) - lse;
                return anneal ? exp(inv_t * lp - lse2) : exp(lp);
            };
            if (p.mode == SEGB_DP_FFBS) {
                // inverse-CDF draw, sequential subtraction in ascending-span order
                // (_cython_utils.pyx:75-89); falls through to the last index
                double uu = p.uniforms[ubase + used];
                used++;
                idx = n_w - 1;
                bool done = false;
                for (int base = l_lo; base <= W && !done; base += 32) {
                    const int l = base + lane;
                    const double pl = (l <= W) ? prob(l) : 0.0;
                    const int hi = min(W - base + 1, 32);
                    for (int i = 0; i < hi; ++i) {
                        uu = uu - __shfl_sync(FULL, pl, i);
                        if (uu < 0) { idx = base + i - l_lo; done = true; break; }
                    }
                }
            } else {
                // argmax of exp-normalised values, first maximum (unigram_acoustic_wordseg.py:843-844)
                double pm = -1.0;
                for (int l = l_lo + lane; l <= W; l += 32) pm = fmax(pm, prob(l));
                pm = warp_max(pm);
                int best = 0x7fffffff;
                for (int l = l_lo + lane; l <= W; l += 32)
                    if (prob(l) == pm) { best = l - l_lo; break; }
                for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(FULL, best, o));
                idx = (best == 0x7fffffff) ? 0 : best;   // all-NaN: np.argmax returns 0
            }
        }
        int k = idx + 1;
        if (n_min > 1) k += n_min - 1;             // reference adds this even after back-tracking
        if (k > t) { status = SEGB_DP_EMPTY_SLICE; break; }
        // spans beyond the band carry no embedding: the packed vector holds -inf there
        total += (k <= S) ? sc[(int64_t)(t - 1) * S + (k - 1)] : neg_inf();
        if (t - k - 1 < 0) break;
        if (lane == 0) bo[t - k - 1] = 1;
        t = t - k;
    }
    __syncwarp();
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
// spends its time in shuffles; here the S candidates of a position live in registers, the
// next position's scores are requested before the current one is reduced, and the float64
// alphas sit in a per-thread local array (L1-resident).  Same operation order as the
// warp kernel / the reference, so results are identical.
constexpr int DP_SMALL_N = 64;
constexpr int DP_SMALL_S = 8;

template <int MODE>
__global__ void __launch_bounds__(128) dp_small_kernel(DpParams p) {
    const int u_local = blockIdx.x * blockDim.x + threadIdx.x;
    if (u_local >= p.n_utt) return;
    const int u = p.utt_first + u_local;
    const int64_t off = p.pos_off[u];
    const int N = (int)(p.pos_off[u + 1] - off);
    const int S = p.S;
    const int Wlim = (p.n_max == 0 || p.n_max > S) ? S : p.n_max;
    const double *sc = p.scores + (p.scores_local ? 0 : off * S);
    uint8_t *bo = p.bounds + off;
    if (N <= 0) { p.status[u_local] = SEGB_DP_OK; p.log_prob[u_local] = 0.0; return; }

    double al[DP_SMALL_N];
    al[0] = 0.0;
    int status = SEGB_DP_OK;
    double cur[DP_SMALL_S], nxt[DP_SMALL_S];
#pragma unroll
    for (int l = 0; l < DP_SMALL_S; ++l) cur[l] = (l < S) ? sc[l] : neg_inf();      // row of t = 1

    // c[l-1] = score(t, l) + alpha[t-l] for l = 1..W; returns max, flags NaN
    auto cands = [&](const double *row, int t, int W, double *c, bool &has_nan) {
        double m = neg_inf();
        has_nan = false;
#pragma unroll
        for (int l = 1; l <= DP_SMALL_S; ++l) {
            if (l <= W) {
                const double v = row[l - 1] + al[t - l];
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
        // request the next row before reducing this one
#pragma unroll
        for (int l = 0; l < DP_SMALL_S; ++l) nxt[l] = (l < S) ? sc[(int64_t)t * S + l] : neg_inf();
        const int W = min(t, Wlim);
        double c[DP_SMALL_S];
        bool has_nan;
        const double m = cands(cur, t, W, c, has_nan);
        double a_t;
        if (has_nan) { status = SEGB_DP_NAN; a_t = neg_inf(); }
        else if (m == neg_inf()) a_t = neg_inf();
        else if (MODE == SEGB_DP_FFBS) a_t = lse_desc(c, W, m) + p.log_p_continue;
        else a_t = m;
        al[t] = a_t;
#pragma unroll
        for (int l = 0; l < DP_SMALL_S; ++l) cur[l] = nxt[l];
    }
    if (p.alphas) for (int j = 0; j < N; ++j) p.alphas[off + j] = al[j];

    unsigned long long bmask = 1ull << (N - 1);
    double total = 0.0;
    int used = 0;
    const int64_t ubase = p.u_counter ? *p.u_counter : off;
    int t = N;
    for (int guard = 0; guard <= N && status == SEGB_DP_OK; ++guard) {
        int W = min(t, Wlim);
        double c[DP_SMALL_S], row[DP_SMALL_S];
        bool has_nan;
#pragma unroll
        for (int l = 0; l < DP_SMALL_S; ++l) row[l] = (l < S) ? sc[(int64_t)(t - 1) * S + l] : neg_inf();
        double m = cands(row, t, W, c, has_nan);
        if (has_nan && MODE != SEGB_DP_VITERBI_GMM) { status = SEGB_DP_NAN; break; }
        if (m == neg_inf()) {
            while (m == neg_inf()) {
                t = t - 1;
                if (t == 0) break;
                W = min(t, Wlim);
#pragma unroll
                for (int l = 0; l < DP_SMALL_S; ++l) row[l] = (l < S) ? sc[(int64_t)(t - 1) * S + l] : neg_inf();
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
    for (int j = 0; j < N; ++j) bo[j] = (uint8_t)((bmask >> j) & 1ull);
    p.log_prob[u_local] = (status == SEGB_DP_OK) ? total : CUDART_NAN;
    p.status[u_local] = status;
    if (p.n_draws) p.n_draws[u_local] = used;
    if (p.u_counter) *p.u_counter = ubase + used;
}

int launch_dp(const segb_corpus *c, int32_t utt_first, int32_t n_utt, const double *scores, int32_t mode,
              double log_p_continue, double anneal_temp, const double *uniforms, int64_t *u_counter,
              uint8_t *bounds_out, double *log_prob, double *alphas, int32_t *n_draws, int32_t *status,
              cudaStream_t stream, int scores_local = 0) {
    DpParams p;
    p.scores_local = scores_local;
    p.pos_off = c->pos_off; p.scores = scores; p.uniforms = uniforms; p.u_counter = u_counter;
    p.bounds = bounds_out; p.log_prob = log_prob; p.alphas = alphas; p.n_draws = n_draws; p.status = status;
    p.utt_first = utt_first; p.n_utt = n_utt; p.S = c->S; p.n_min = c->n_slices_min; p.n_max = c->n_slices_max;
    p.mode = mode; p.N_cap = c->N_max; p.log_p_continue = log_p_continue; p.anneal_temp = anneal_temp;
    if (c->S <= DP_SMALL_S && c->N_max <= DP_SMALL_N && c->n_slices_min <= 1) {
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
    SEGB_CHECK_ARG(mode >= 0 && mode <= 2, "mode");
    SEGB_CHECK_ARG(n_utt >= 0 && utt_first >= 0 && utt_first + n_utt <= c->n_utt, "utterance range");
    SEGB_CHECK_ARG(c->S >= 1, "band width");
    SEGB_CHECK_ARG(mode != SEGB_DP_FFBS || uniforms, "FFBS needs uniforms");
    SEGB_CHECK_ARG(!u_counter || n_utt <= 1, "u_counter implies a single utterance");
    if (n_utt == 0) return 0;
    return segb::launch_dp(c, utt_first, n_utt, scores, mode, log_p_continue, anneal_temp, uniforms,
                           u_counter, bounds_out, log_prob, alphas, n_draws, status, (cudaStream_t)stream);
}
