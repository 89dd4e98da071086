// Warp-per-utterance segmentation DP body shared by dp_banded_kernel (dp.cu) and the persistent
// Gibbs sweep (fixedvar_gibbs.cu).  See dp.cu for the reference functions it replaces.
#pragma once
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
    int32_t group;             // utterances per warp of the staged kernel (<= 32)
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

// Forward-filter / backward-sample (and the GMM Viterbi variant) for the common case of the sequential Gibbs
// sweep: n_slices_min <= 1 and a window of at most 8 spans, with or without annealing.  Lane l-1 holds span l, so the window
// maximum is three xor steps inside lanes 0..7, every exponential is evaluated once, and the sums run in
// the reference's order (descending span for the logsumexp, _cython_utils.pyx:13-25; ascending span for
// the draw, :75-89) -- the same values, bit for bit, as the general routine below, at ~1/3 of its
// latency.  Returns false without side effects that matter (bo / al are re-initialised by the caller's
// general path) on anything unusual: NaN scores, an infeasible window in the backward pass.
__device__ __forceinline__ bool dp_warp_ffbs_small(const DpParams &p, const double *sc, uint8_t *bo, int N, double *al,
                                                   int64_t ubase, int Wlim, double &total_out, int &used_out) {
    const int lane = threadIdx.x & 31;
    const int S = p.S;
    const bool viterbi = (p.mode == SEGB_DP_VITERBI_GMM);        // max instead of logsumexp, argmax instead of a draw
    const bool anneal = !viterbi && p.anneal_temp != 1.0;
    const double inv_t = 1. / p.anneal_temp;
    for (int j = lane; j < N; j += 32) bo[j] = (j == N - 1);
    if (lane == 0) al[0] = 0.0;
    __syncwarp();
    auto win = [&](int t, int W, double &m) -> double {          // candidate of this lane's span and the window maximum
        const double c = (lane < W) ? sc[(int64_t)(t - 1) * S + lane] + al[t - 1 - lane] : neg_inf();
        m = c;
        m = fmax(m, __shfl_xor_sync(FULL, m, 1));
        m = fmax(m, __shfl_xor_sync(FULL, m, 2));
        m = fmax(m, __shfl_xor_sync(FULL, m, 4));
        m = __shfl_sync(FULL, m, 0);
        return c;
    };
    auto lse_desc = [&](double c, int W, double m) -> double {   // log(sum_{l = W..1} exp(c_l - m)) + m
        const double e = (lane < W) ? exp(c - m) : 0.0;
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __shfl_sync(FULL, e, i);
        double ssum = 0.0;
#pragma unroll
        for (int i = 7; i >= 0; --i) if (i < W) ssum += x[i];
        return log(ssum) + m;
    };
    bool bad = false;
    for (int t = 1; t < N; ++t) {
        const int W = min(t, Wlim);
        double m;
        const double c = win(t, W, m);
        bad |= (c != c);
        const double a_t = (m == neg_inf()) ? neg_inf() : (viterbi ? m : lse_desc(c, W, m) + p.log_p_continue);
        __syncwarp();
        if (lane == 0) al[t] = a_t;
        __syncwarp();
    }
    if (__any_sync(FULL, bad)) return false;
    double total = 0.0;
    int used = 0, t = N;
    for (int guard = 0; guard <= N; ++guard) {
        const int W = min(t, Wlim);
        double m;
        const double c = win(t, W, m);
        if (m == neg_inf() || m != m) return false;               // back-tracking: general routine
        const double lse = lse_desc(c, W, m);
        double pl;
        if (anneal) {
            // p = softmax((1/T) * log-normalised p) (unigram_acoustic_wordseg.py:731-736); the second
            // logsumexp runs over the reversed window, i.e. in ASCENDING span order
            const double q = (lane < W) ? inv_t * (c - lse) : neg_inf();
            double mq = q;
            mq = fmax(mq, __shfl_xor_sync(FULL, mq, 1));
            mq = fmax(mq, __shfl_xor_sync(FULL, mq, 2));
            mq = fmax(mq, __shfl_xor_sync(FULL, mq, 4));
            mq = __shfl_sync(FULL, mq, 0);
            const double e2 = (lane < W) ? exp(q - mq) : 0.0;
            double y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = __shfl_sync(FULL, e2, i);
            double s2 = 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i) if (i < W) s2 += y[i];
            const double lse2 = log(s2) + mq;
            pl = (lane < W) ? exp(q - lse2) : 0.0;
        } else {
            pl = (lane < W) ? exp(c - lse) : 0.0;
        }
        int idx = W - 1;
        if (viterbi) {
            // argmax of the exp-normalised window, first maximum = shortest span (:843-844)
            double pm = pl;
            pm = fmax(pm, __shfl_xor_sync(FULL, pm, 1));
            pm = fmax(pm, __shfl_xor_sync(FULL, pm, 2));
            pm = fmax(pm, __shfl_xor_sync(FULL, pm, 4));
            pm = __shfl_sync(FULL, pm, 0);
            const unsigned hit = __ballot_sync(FULL, lane < W && pl == pm);
            if (hit == 0) return false;
            idx = __ffs(hit) - 1;
        } else {
            double x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __shfl_sync(FULL, pl, i);
            double uu = p.uniforms[ubase + used];
            used++;
            bool done = false;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < W && !done) {
                    uu = uu - x[i];
                    if (uu < 0) { idx = i; done = true; }
                }
            }
        }
        const int k = idx + 1;
        total += sc[(int64_t)(t - 1) * S + (k - 1)];
        if (t - k - 1 < 0) break;
        if (lane == 0) bo[t - k - 1] = 1;
        t = t - k;
    }
    __syncwarp();
    total_out = total;
    used_out = used;
    return true;
}

// One warp runs the DP of one utterance: `sc` = its banded scores (row t-1, column l-1), `bo` =
// its N boundary bytes, `al` = N doubles of scratch (shared memory), draws from
// p.uniforms[ubase ...].  Outputs are warp-uniform.
__device__ __forceinline__ void dp_warp_body(const DpParams &p, const double *sc, uint8_t *bo, int N, double *al,
                                             double *alphas_out, int64_t ubase, double &total_out, int &status_out,
                                             int &used_out) {
    const int lane = threadIdx.x & 31;
    const int S = p.S;
    const int Wlim = (p.n_max == 0 || p.n_max > S) ? S : p.n_max;   // window limit in spans
    const int n_min = p.n_min;
    const int l_cut = n_min > 1 ? n_min : 1;                        // [-S : -(n_min-1)] keeps spans >= n_min
    int status = SEGB_DP_OK;
    if ((p.mode == SEGB_DP_FFBS || p.mode == SEGB_DP_VITERBI_GMM) && n_min <= 1 && Wlim <= 8 && alphas_out == nullptr) {
        double tot;
        int usd;
        if (dp_warp_ffbs_small(p, sc, bo, N, al, ubase, Wlim, tot, usd)) {
            total_out = tot; status_out = SEGB_DP_OK; used_out = usd;
            return;
        }
        __syncwarp();
    }
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
    if (alphas_out) for (int j = lane; j < N; j += 32) alphas_out[j] = al[j];

    // ---------------- backward pass
    double total = 0.0;
    int used = 0;
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
    total_out = total; status_out = status; used_out = used;
}

}  // namespace segb
