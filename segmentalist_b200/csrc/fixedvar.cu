// Fixed-variance Bayesian GMM: component statistics, predictive scoring,
// log_marg_i, assignment sampling and the sequential Gibbs sweep.
//
// Replaces (reference paths relative to segmentalist/):
//   GaussianComponentsFixedVar.add_item/del_item/del_component  gaussian_components_fixedvar.py:153-221
//   GaussianComponentsFixedVar.log_post_pred / log_prior        gaussian_components_fixedvar.py:224-253
//   FBGMM.log_marg_i                                            fbgmm.py:256-285
//   FBGMM.gibbs_sample_inside_loop_i / map_assign_i             fbgmm.py:422-494
//   UnigramAcousticWordseg.gibbs_sample_i / get_vec_embed_log_probs
//                                                               unigram_acoustic_wordseg.py:252-360,474-511
//
// All statistics and scores are float64, as in the reference (float32 embeddings
// are promoted at the first subtraction, gaussian_components_fixedvar.py:247-248).
// State updates use separately rounded multiply/add (no FMA contraction) so the
// sufficient statistics carry the same bits as the NumPy code.
#include "fixedvar_common.cuh"

namespace segb {

int launch_dp_local(const segb_corpus *c, int32_t utt, const double *local_scores, int32_t mode,
                    double log_p_continue, double anneal_temp, const double *uniforms, int64_t *u_counter,
                    double *log_prob, int32_t *status, cudaStream_t stream);

#define NEG_HALF_LOG_2PI (-0.91893853320467274178)

// Recompute precision_pred, mu_N and log_prod_precision_pred of component k (:317-325).
// Block-cooperative; `tmp` is shared scratch of >= D doubles.
__device__ void fv_refresh(const segb_fixedvar &m, int k, double *tmp) {
    const int D = m.D, KM = m.K_max;
    if (m.model == SEGB_MODEL_DIAG) {
        // _update_log_prod_vars_and_inv_vars (gaussian_components_diag.py:332-345)
        const double k_N = m.k_0 + (double)m.counts[k], v_N = (double)(m.v_0 + m.counts[k]);
        const double f = __ddiv_rn(k_N + 1., __dmul_rn(k_N, v_N));
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const size_t o = (size_t)d * KM + k;
            const double mN = __ddiv_rn(m.mu_N_numT[o], k_N);
            const double var = __dmul_rn(f, __dsub_rn(m.prec_NT[o], __dmul_rn(k_N, __dmul_rn(mN, mN))));
            m.mu_NT[o] = mN;
            m.prec_predT[o] = __ddiv_rn(1., var);
            tmp[d] = log(var);
        }
    } else {
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const double pN = m.prec_NT[(size_t)d * KM + k];
            const double pr = m.precision[d];
            const double pp = __ddiv_rn(__dmul_rn(pN, pr), __dadd_rn(pN, pr));
            m.prec_predT[(size_t)d * KM + k] = pp;
            m.mu_NT[(size_t)d * KM + k] = __ddiv_rn(m.mu_N_numT[(size_t)d * KM + k], pN);
            tmp[d] = log(pp);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)   // np.log(.).sum(): NumPy pairwise order
        m.log_prod_prec_pred[k] = pairwise_sum<double>([&](int i) { return tmp[i]; }, D);
    __syncthreads();
}

__device__ void fv_add_item(const segb_fixedvar &m, int id, int k, double *tmp) {
    const int D = m.D, KM = m.K_max;
    const int K = *m.K;
    __syncthreads();
    const bool fresh = (k == K);
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const size_t o = (size_t)d * KM + k;
        double num = m.mu_N_numT[o], pN = m.prec_NT[o];
        if (m.model == SEGB_MODEL_DIAG) {               // gaussian_components_diag.py:162-177
            if (fresh) {
                num = __dmul_rn(m.k_0, m.mu_0[d]);
                pN = __dadd_rn(m.precision_0[d], __dmul_rn(m.k_0, __dmul_rn(m.mu_0[d], m.mu_0[d])));
            }
            m.mu_N_numT[o] = __dadd_rn(num, fv_x(m, id, d));
            m.prec_NT[o] = __dadd_rn(pN, fv_xsq(m, id, d));
        } else {
            if (fresh) { num = __dmul_rn(m.precision_0[d], m.mu_0[d]); pN = m.precision_0[d]; }
            m.mu_N_numT[o] = __dadd_rn(num, __dmul_rn(m.precision[d], fv_x(m, id, d)));
            m.prec_NT[o] = __dadd_rn(pN, m.precision[d]);
        }
    }
    if (threadIdx.x == 0) {
        if (fresh) *m.K = K + 1;
        m.counts[k] += 1;
        *m.n_total += 1;
        m.assignments[id] = k;
    }
    __syncthreads();
    fv_refresh(m, k, tmp);
}

// del_component (:190-221): move the last component into slot k, relabel its members.
__device__ void fv_del_component(const segb_fixedvar &m, int k, const int32_t *relabel_ids, int64_t relabel_n,
                                 const segb_bigram_lm *lm = nullptr) {
    const int D = m.D, KM = m.K_max;
    const int last = *m.K - 1;
    __syncthreads();
    if (lm) {
        // tied bigram LM counts move with the component (:205-208), in the reference's order: the
        // unigram count, then row `last` -> row k, then column `last` -> column k; then the old
        // slot is cleared (:218-221)
        const int KL = lm->K;
        int32_t *bi = lm->bigram_counts;
        if (k != last) {
            if (threadIdx.x == 0) lm->unigram_counts[k] = lm->unigram_counts[last];
            for (int i = threadIdx.x; i < KL; i += blockDim.x) bi[(size_t)k * KL + i] = bi[(size_t)last * KL + i];
            __syncthreads();
            for (int j = threadIdx.x; j < KL; j += blockDim.x) bi[(size_t)j * KL + k] = bi[(size_t)j * KL + last];
            __syncthreads();
        }
        if (threadIdx.x == 0) lm->unigram_counts[last] = 0;
        for (int i = threadIdx.x; i < KL; i += blockDim.x) { bi[(size_t)last * KL + i] = 0; bi[(size_t)i * KL + last] = 0; }
        __syncthreads();
    }
    if (k != last) {
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const size_t a = (size_t)d * KM + k, b = (size_t)d * KM + last;
            m.mu_N_numT[a] = m.mu_N_numT[b];
            m.prec_NT[a] = m.prec_NT[b];
            m.prec_predT[a] = m.prec_predT[b];
            m.mu_NT[a] = m.mu_NT[b];
        }
        if (relabel_ids) {
            for (int64_t i = threadIdx.x; i < relabel_n; i += blockDim.x) {
                const int id = relabel_ids[i];
                if (id >= 0 && m.assignments[id] == last) m.assignments[id] = k;
            }
        } else {
            for (int64_t i = threadIdx.x; i < m.n_emb; i += blockDim.x)
                if (m.assignments[i] == last) m.assignments[i] = k;
        }
    }
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const size_t b = (size_t)d * KM + last;
        m.mu_N_numT[b] = 0.; m.prec_NT[b] = 0.; m.prec_predT[b] = 0.; m.mu_NT[b] = 0.;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (k != last) { m.log_prod_prec_pred[k] = m.log_prod_prec_pred[last]; m.counts[k] = m.counts[last]; }
        m.log_prod_prec_pred[last] = 0.;
        m.counts[last] = 0;
        *m.K = last;
    }
    __syncthreads();
}

__device__ void fv_del_item(const segb_fixedvar &m, int id, double *tmp, const int32_t *relabel_ids,
                            int64_t relabel_n, const segb_bigram_lm *lm = nullptr) {
    __syncthreads();
    const int k = m.assignments[id];
    if (k == -1) return;            // uniform across the block
    __syncthreads();
    const int cnt = m.counts[k] - 1;
    __syncthreads();
    if (threadIdx.x == 0) { m.counts[k] = cnt; m.assignments[id] = -1; *m.n_total -= 1; }
    __syncthreads();
    if (cnt == 0) {
        fv_del_component(m, k, relabel_ids, relabel_n, lm);
    } else {
        const int D = m.D, KM = m.K_max;
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const size_t o = (size_t)d * KM + k;
            if (m.model == SEGB_MODEL_DIAG) {           // gaussian_components_diag.py:190-193
                m.mu_N_numT[o] = __dsub_rn(m.mu_N_numT[o], fv_x(m, id, d));
                m.prec_NT[o] = __dsub_rn(m.prec_NT[o], fv_xsq(m, id, d));
            } else {
                m.mu_N_numT[o] = __dsub_rn(m.mu_N_numT[o], __dmul_rn(m.precision[d], fv_x(m, id, d)));
                m.prec_NT[o] = __dsub_rn(m.prec_NT[o], m.precision[d]);
            }
        }
        __syncthreads();
        fv_refresh(m, k, tmp);
    }
}

// ---------------------------------------------------------------- scoring

// log N(x; mu_0, 1/precision_0) summed over dimensions (:224-231).  sum(log precision_0)
// is a model constant supplied by the host (formed in the reference's sequential order);
// the quadratic form is reduced across the block.
__device__ double fv_log_prior_x(const segb_fixedvar &m, const double *xs, double *red) {
    if (m.model == SEGB_MODEL_DIAG) {
        // gaussian_components_diag.py:216-223 via _log_prod_students_t (:347-360)
        const double f = (m.k_0 + 1.) / (m.k_0 * m.v_0), iv = 1. / m.v_0;
        double acc = 0.0, lpv = 0.0;
        for (int d = threadIdx.x; d < m.D; d += blockDim.x) {
            const double var = f * m.precision_0[d];
            const double dl = xs[d] - m.mu_0[d];
            lpv += log(var);
            acc += log(1. + iv * (dl * dl) * (1. / var));
        }
        acc = block_sum(acc, red);
        lpv = block_sum(lpv, red);
        double cst, hv, iv2;
        diag_consts(m, 0, cst, hv, iv2);
        return cst - 0.5 * lpv - hv * acc;
    }
    double sq = 0.0;
    for (int d = threadIdx.x; d < m.D; d += blockDim.x) {
        const double dl = xs[d] - m.mu_0[d];
        sq += dl * dl * m.precision_0[d];
    }
    sq = block_sum(sq, red);
    return fv_norm_const(m.D) + 0.5 * m.sum_log_precision_0 - 0.5 * sq;
}

// log_post_pred of the item in xs for every active component -> sk[k] (k < K)  (:242-253)
__device__ void fv_post_pred_all(const segb_fixedvar &m, const double *xs, double *sk, int K) {
    const int D = m.D, KM = m.K_max;
    const double c0 = fv_norm_const(D);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double acc = 0.0;
        const double *mu = m.mu_NT + k, *pp = m.prec_predT + k;
        if (m.model == SEGB_MODEL_DIAG) {               // gaussian_components_diag.py:237-259
            double cst, hv, iv;
            diag_consts(m, m.counts[k], cst, hv, iv);
#pragma unroll 2
            for (int d = 0; d < D; ++d) {
                const double dl = mu[(size_t)d * KM] - xs[d];
                acc += log(1. + (dl * dl) * pp[(size_t)d * KM] * iv);
            }
            sk[k] = cst - 0.5 * m.log_prod_prec_pred[k] - hv * acc;
            continue;
        }
#pragma unroll 4
        for (int d = 0; d < D; ++d) {
            const double dl = mu[(size_t)d * KM] - xs[d];
            acc += dl * dl * pp[(size_t)d * KM];
        }
        sk[k] = c0 + 0.5 * m.log_prod_prec_pred[k] - 0.5 * acc;
    }
}

// FBGMM.log_marg_i for one item per block.
__global__ void __launch_bounds__(256) fv_log_marg_kernel(segb_fixedvar m, const int32_t *ids, const double *durs,
                                                          int64_t n, double tpt, double wip, double *out) {
    extern __shared__ double smem[];
    ScoreSmem s(smem, m.D);
    for (int64_t it = blockIdx.x; it < n; it += gridDim.x) {
        const int id = ids[it];
        const bool dead = (id < 0) || (durs && durs[it] != durs[it]);
        if (dead) { if (threadIdx.x == 0) out[it] = neg_inf(); continue; }   // -inf + wip == -inf
        __syncthreads();
        for (int d = threadIdx.x; d < m.D; d += blockDim.x) s.xs[d] = fv_x(m, id, d);
        __syncthreads();
        const int K = *m.K, KM = m.K_max;
        const double log_norm = log((double)(*m.n_total) + m.alpha);
        fv_post_pred_all(m, s.xs, s.sk, K);
        const double lprior = fv_log_prior_x(m, s.xs, s.red);
        double mx = neg_inf();
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const double v = m.lms * (log(m.alpha / KM + m.counts[k]) - log_norm) + s.sk[k];
            s.sk[k] = v;
            mx = fmax(mx, v);
        }
        const int n_empty = KM - K;
        const double e = m.lms * (log(m.alpha / KM + 0.) - log_norm) + lprior;
        if (n_empty > 0) mx = fmax(mx, e);
        mx = block_max(mx, s.red);
        double sum = 0.0;
        for (int k = threadIdx.x; k < K; k += blockDim.x) sum += exp(s.sk[k] - mx);
        sum = block_sum(sum, s.red);
        if (threadIdx.x == 0) {
            if (n_empty > 0) sum += n_empty * exp(e - mx);
            double v = log(sum) + mx;
            if (durs) {
                const double du = durs[it];
                v *= (tpt == 1.0) ? du : pow(du, tpt);
            }
            out[it] = v + wip;
        }
    }
}

// One row: log_post_pred for active slots, log_prior for the rest.
__global__ void __launch_bounds__(256) fv_pred_row_kernel(segb_fixedvar m, int id, double *out) {
    extern __shared__ double smem[];
    ScoreSmem s(smem, m.D);
    for (int d = threadIdx.x; d < m.D; d += blockDim.x) s.xs[d] = fv_x(m, id, d);
    __syncthreads();
    const int K = *m.K;
    fv_post_pred_all(m, s.xs, s.sk, K);
    const double lprior = fv_log_prior_x(m, s.xs, s.red);
    for (int k = threadIdx.x; k < m.K_max; k += blockDim.x) out[k] = k < K ? s.sk[k] : lprior;
}

// ---------------------------------------------------------------- assignment (sequential)

// Sample / MAP-assign one item whose embedding is already in s.xs; returns k (block-uniform).
// mode 0: gibbs_sample_inside_loop_i (fbgmm.py:422-463); mode 1: map_assign_i (:465-494).
__device__ int fv_choose(const segb_fixedvar &m, ScoreSmem &s, int mode, double anneal_temp, double u) {
    const int K = *m.K, KM = m.K_max;
    fv_post_pred_all(m, s.xs, s.sk, K);
    const double lprior = fv_log_prior_x(m, s.xs, s.red);
    const double scale = (mode == 0) ? m.lms : 1.0;
    // physical slots: K actives, then KM-K empties with identical value
    double mx = neg_inf();
    for (int k = threadIdx.x; k < KM; k += blockDim.x) {
        const double v = scale * log(m.alpha / KM + (k < K ? m.counts[k] : 0)) + (k < K ? s.sk[k] : lprior);
        s.sk[k] = v;
        mx = fmax(mx, v);
    }
    return fv_decide(s, K, KM, mode, anneal_temp, u, mx);
}

__global__ void __launch_bounds__(1024) fv_assign_list_kernel(segb_fixedvar m, const int32_t *ids, int n, int mode,
                                                              double anneal_temp, const double *uniforms,
                                                              int64_t *u_counter, int32_t *ks_out) {
    extern __shared__ double smem[];
    ScoreSmem s(smem, m.D);
    int64_t upos = (mode == 0) ? *u_counter : 0;
    for (int i = 0; i < n; ++i) {
        const int id = ids[i];
        if (id < 0) { if (ks_out && threadIdx.x == 0) ks_out[i] = -1; continue; }
        __syncthreads();
        for (int d = threadIdx.x; d < m.D; d += blockDim.x) s.xs[d] = fv_x(m, id, d);
        __syncthreads();
        const double u = (mode == 0) ? uniforms[upos++] : 0.0;
        const int k = fv_choose(m, s, mode, anneal_temp, u);
        fv_add_item(m, id, k, s.sk);
        if (ks_out && threadIdx.x == 0) ks_out[i] = k;
    }
    __syncthreads();
    if (mode == 0 && threadIdx.x == 0) *u_counter = upos;
}

__global__ void __launch_bounds__(256) fv_add_list_kernel(segb_fixedvar m, const int32_t *ids, const int32_t *ks, int n) {
    extern __shared__ double smem[];
    for (int i = 0; i < n; ++i) {
        int k = ks[i];
        __syncthreads();
        const int K = *m.K;
        if (k > K) k = K;
        fv_add_item(m, ids[i], k, smem);
    }
}

__global__ void __launch_bounds__(256) fv_del_list_kernel(segb_fixedvar m, const int32_t *ids, int n,
                                                          const int32_t *relabel_ids, int64_t relabel_n) {
    extern __shared__ double smem[];
    for (int i = 0; i < n; ++i)
        if (ids[i] >= 0) fv_del_item(m, ids[i], smem, relabel_ids, relabel_n);
}

__global__ void __launch_bounds__(256) fv_del_list_lm_kernel(segb_fixedvar m, segb_bigram_lm lm, const int32_t *ids, int n,
                                                             const int32_t *relabel_ids, int64_t relabel_n) {
    extern __shared__ double smem[];
    for (int i = 0; i < n; ++i)
        if (ids[i] >= 0) fv_del_item(m, ids[i], smem, relabel_ids, relabel_n, &lm);
}

// ---------------------------------------------------------------- per-utterance Gibbs steps

// Remove the current tokens of utterance u from the model (unigram_acoustic_wordseg.py:270-273).
__global__ void __launch_bounds__(256) fv_remove_utt_kernel(segb_fixedvar m, segb_corpus c, int u) {
    extern __shared__ double smem[];
    const int64_t off = c.pos_off[u];
    const int N = (int)(c.pos_off[u + 1] - off);
    for (int j = 0; j < N; ++j) {
        const int id = c.tok_id[off + j];
        if (id >= 0) fv_del_item(m, id, smem, c.tok_id, c.n_pos);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += blockDim.x) c.tok_id[off + j] = -1;
}

// get_vec_embed_log_probs for the banded slots of utterance u (:474-511): one block per slot.
__global__ void __launch_bounds__(256) fv_score_utt_kernel(segb_fixedvar m, segb_corpus c, int u, double tpt,
                                                           double wip, double *local_scores) {
    extern __shared__ double smem[];
    ScoreSmem s(smem, m.D);
    const int64_t off = c.pos_off[u];
    const int N = (int)(c.pos_off[u + 1] - off);
    const int n_slots = N * c.S;
    for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
        const int id = c.seg_id[off * c.S + slot];
        const double du = c.seg_dur[off * c.S + slot];
        if (id < 0 || du != du) { if (threadIdx.x == 0) local_scores[slot] = neg_inf(); continue; }
        __syncthreads();
        for (int d = threadIdx.x; d < m.D; d += blockDim.x) s.xs[d] = fv_x(m, id, d);
        __syncthreads();
        const int K = *m.K, KM = m.K_max;
        const double log_norm = log((double)(*m.n_total) + m.alpha);
        fv_post_pred_all(m, s.xs, s.sk, K);
        const double lprior = fv_log_prior_x(m, s.xs, s.red);
        double mx = neg_inf();
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const double v = m.lms * (log(m.alpha / KM + m.counts[k]) - log_norm) + s.sk[k];
            s.sk[k] = v;
            mx = fmax(mx, v);
        }
        const int n_empty = KM - K;
        const double e = m.lms * (log(m.alpha / KM + 0.) - log_norm) + lprior;
        if (n_empty > 0) mx = fmax(mx, e);
        mx = block_max(mx, s.red);
        double sum = 0.0;
        for (int k = threadIdx.x; k < K; k += blockDim.x) sum += exp(s.sk[k] - mx);
        sum = block_sum(sum, s.red);
        if (threadIdx.x == 0) {
            if (n_empty > 0) sum += n_empty * exp(e - mx);
            double v = log(sum) + mx;
            v *= (tpt == 1.0) ? du : pow(du, tpt);
            local_scores[slot] = v + wip;
        }
    }
}

// Assign the tokens of the new segmentation left to right (:339-349).
__global__ void __launch_bounds__(1024) fv_assign_utt_kernel(segb_fixedvar m, segb_corpus c, int u, int mode,
                                                             double anneal_temp, const double *uniforms,
                                                             int64_t *u_counter, const int32_t *dp_status) {
    extern __shared__ double smem[];
    ScoreSmem s(smem, m.D);
    if (*dp_status != SEGB_DP_OK) return;
    const int64_t off = c.pos_off[u];
    const int N = (int)(c.pos_off[u + 1] - off);
    int64_t upos = (mode == 0) ? *u_counter : 0;
    int j_prev = 0;
    for (int j = 0; j < N; ++j) {
        if (!c.bounds[off + j]) continue;
        const int t = j + 1, l = t - j_prev;
        j_prev = j + 1;
        const int id = (l <= c.S) ? c.seg_id[(off + t - 1) * c.S + (l - 1)] : -1;
        __syncthreads();
        if (threadIdx.x == 0) c.tok_id[off + j] = id;
        if (id < 0) continue;       // back-tracking leftovers are skipped (:340-342)
        for (int d = threadIdx.x; d < m.D; d += blockDim.x) s.xs[d] = fv_x(m, id, d);
        __syncthreads();
        const double uu = (mode == 0) ? uniforms[upos++] : 0.0;
        const int k = fv_choose(m, s, mode, anneal_temp, uu);
        fv_add_item(m, id, k, s.sk);
    }
    __syncthreads();
    if (mode == 0 && threadIdx.x == 0) *u_counter = upos;
}

// ---------------------------------------------------------------- bulk construction

// The constructor's loop `for k: for i in where(assignments == k): add_item(i, k)` (:111-120) for a model
// that is still EMPTY, without its one-item-at-a-time dependency: a component's statistics only depend on
// its own members in index order, so thread (k, d) replays that component's additions for one dimension --
// the same sequence of separately rounded operations as fv_add_item, bit for bit.  `order` = item ids
// stably sorted by assignment, seg_off[k] = start of component k's members (as for the diagnostics).
__global__ void fv_build_stats_kernel(segb_fixedvar m, const int64_t *order, const int64_t *seg_off, int K_new) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)K_new * m.D) return;
    const int k = (int)(idx / m.D), d = (int)(idx % m.D);
    const int KM = m.K_max;
    double num, pN;
    if (m.model == SEGB_MODEL_DIAG) {                   // gaussian_components_diag.py:162-177
        num = __dmul_rn(m.k_0, m.mu_0[d]);
        pN = __dadd_rn(m.precision_0[d], __dmul_rn(m.k_0, __dmul_rn(m.mu_0[d], m.mu_0[d])));
        for (int64_t j = seg_off[k]; j < seg_off[k + 1]; ++j) {
            num = __dadd_rn(num, fv_x(m, order[j], d));
            pN = __dadd_rn(pN, fv_xsq(m, order[j], d));
        }
    } else {
        const double pr = m.precision[d];
        num = __dmul_rn(m.precision_0[d], m.mu_0[d]);
        pN = m.precision_0[d];
        for (int64_t j = seg_off[k]; j < seg_off[k + 1]; ++j) {
            num = __dadd_rn(num, __dmul_rn(pr, fv_x(m, order[j], d)));
            pN = __dadd_rn(pN, pr);
        }
    }
    const size_t o = (size_t)d * KM + k;
    m.mu_N_numT[o] = num;
    m.prec_NT[o] = pN;
}

// counts, assignments and the derived tables (fv_refresh) of the freshly built components: block per component.
__global__ void __launch_bounds__(256) fv_build_finish_kernel(segb_fixedvar m, const int64_t *order, const int64_t *seg_off,
                                                              int K_new) {
    extern __shared__ double smem[];
    const int k = blockIdx.x;
    const int64_t lo = seg_off[k], hi = seg_off[k + 1];
    for (int64_t j = lo + threadIdx.x; j < hi; j += blockDim.x) m.assignments[order[j]] = k;
    if (threadIdx.x == 0) {
        m.counts[k] = (int32_t)(hi - lo);
        if (k == 0) { *m.K = K_new; *m.n_total = seg_off[K_new] - seg_off[0]; }
    }
    __syncthreads();
    fv_refresh(m, k, smem);
}

// ---------------------------------------------------------------- bigram LM + bigram cluster sampling

// counts_from_utterance (+1, bigram_lms.py:98-105) / remove_counts_from_utterance (-1, :107-113):
// a handful of dependent integer updates -- one thread.
__device__ void lm_update(const segb_bigram_lm &lm, const int32_t *tr, int n, int sign) {
    if (threadIdx.x == 0) {
        int j_prev = -1;
        for (int t = 0; t < n; ++t) {
            const int i = tr[t];
            if (i < 0) continue;
            lm.unigram_counts[i] += sign;
            if (j_prev >= 0) lm.bigram_counts[(size_t)j_prev * lm.K + i] += sign;
            j_prev = i;
        }
    }
    __syncthreads();
}

// sum_ints(unigram_counts) (_cython_utils.pyx), exact: block-wide integer sum.  red: >= 33 doubles.
__device__ long long lm_total(const segb_bigram_lm &lm, double *red) {
    long long t = 0;
    for (int k = threadIdx.x; k < lm.K; k += blockDim.x) t += lm.unigram_counts[k];
    return (long long)block_sum((double)t, red);      // partial sums are far below 2^53: exact
}

__global__ void __launch_bounds__(256) bg_lm_update_kernel(segb_bigram_lm lm, const int32_t *tr, int n, int sign) {
    lm_update(lm, tr, n, sign);
}

__global__ void __launch_bounds__(256) bg_lm_row_kernel(segb_bigram_lm lm, int j_prev, double *out) {
    __shared__ double red[40];
    const double sum_a = __dadd_rn((double)lm_total(lm, red), lm.a);
    for (int k = threadIdx.x; k < lm.K; k += blockDim.x) out[k] = lm_log_prob(lm, j_prev, k, sum_a);
}

// gibbs_sample_i, first part (bigram_acoustic_wordseg.py:410-417): the utterance's transcript leaves the
// LM, then its tokens leave the components (the LM tie follows every component move).
__global__ void __launch_bounds__(256) bg_remove_utt_kernel(segb_fixedvar m, segb_bigram_lm lm, segb_corpus c, int u) {
    extern __shared__ double smem[];
    const int64_t off = c.pos_off[u];
    const int N = (int)(c.pos_off[u + 1] - off);
    if (threadIdx.x == 0) {
        int j_prev = -1;
        for (int j = 0; j < N; ++j) {
            const int id = c.tok_id[off + j];
            if (id < 0) continue;
            const int i = m.assignments[id];
            lm.unigram_counts[i] -= 1;
            if (j_prev >= 0) lm.bigram_counts[(size_t)j_prev * lm.K + i] -= 1;
            j_prev = i;
        }
    }
    __syncthreads();
    for (int j = 0; j < N; ++j) {
        const int id = c.tok_id[off + j];
        if (id >= 0) fv_del_item(m, id, smem, c.tok_id, c.n_pos, &lm);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += blockDim.x) c.tok_id[off + j] = -1;
}

// gibbs_sample_inside_loop_i_embed (:333-384) for the item staged in s.xs.
__device__ int bg_choose(const segb_fixedvar &m, const segb_bigram_lm &lm, ScoreSmem &s, int j_prev,
                         double anneal_temp, double u) {
    const int K = *m.K, KM = m.K_max;
    fv_post_pred_all(m, s.xs, s.sk, K);
    const double lprior = fv_log_prior_x(m, s.xs, s.red);
    const double sum_a = __dadd_rn((double)lm_total(lm, s.red), lm.a);
    double mx = neg_inf();
    for (int k = threadIdx.x; k < KM; k += blockDim.x) {
        const double v = __dadd_rn(__dmul_rn(lm_log_prob(lm, j_prev, k, sum_a), m.lms), (k < K ? s.sk[k] : lprior));
        s.sk[k] = v;
        mx = fmax(mx, v);
    }
    return fv_decide(s, K, KM, 0, anneal_temp, u, mx);
}

// gibbs_sample_i, last part (:483-499): sample the new tokens' components left to right under the
// bigram prior, then add the new transcript to the LM.  When the boundaries are kept
// (assignments_only) the tokens are those of the current segmentation.
__global__ void __launch_bounds__(1024) bg_assign_utt_kernel(segb_fixedvar m, segb_bigram_lm lm, segb_corpus c, int u,
                                                             double anneal_temp, const double *uniforms,
                                                             int64_t *u_counter, const int32_t *dp_status) {
    extern __shared__ double smem[];
    ScoreSmem s(smem, m.D);
    if (dp_status && *dp_status != SEGB_DP_OK) return;
    const int64_t off = c.pos_off[u];
    const int N = (int)(c.pos_off[u + 1] - off);
    int64_t upos = *u_counter;
    int j_prev = 0, k_prev = -1;
    for (int j = 0; j < N; ++j) {
        if (!c.bounds[off + j]) continue;
        const int t = j + 1, l = t - j_prev;
        j_prev = j + 1;
        const int id = (l <= c.S) ? c.seg_id[(off + t - 1) * c.S + (l - 1)] : -1;
        __syncthreads();
        if (threadIdx.x == 0) c.tok_id[off + j] = id;
        if (id < 0) continue;       // back-tracking leftovers are skipped (:485-487)
        for (int d = threadIdx.x; d < m.D; d += blockDim.x) s.xs[d] = fv_x(m, id, d);
        __syncthreads();
        int k = bg_choose(m, lm, s, k_prev, anneal_temp, uniforms[upos++]);
        __syncthreads();
        const int Kact = *m.K;
        if (k > Kact) k = Kact;     // several empty slots may follow the active ones (:371-372)
        fv_add_item(m, id, k, s.sk);
        k_prev = k;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *u_counter = upos;
        int jp = -1;                // counts_from_utterance over the new transcript (:499)
        for (int j = 0; j < N; ++j) {
            const int id = c.tok_id[off + j];
            if (id < 0) continue;
            const int i = m.assignments[id];
            lm.unigram_counts[i] += 1;
            if (jp >= 0) lm.bigram_counts[(size_t)jp * lm.K + i] += 1;
            jp = i;
        }
    }
}

static int score_smem_attr(const void *fn, size_t bytes) {
    if (bytes > 220 * 1024) { set_error("K_max too large for shared-memory scoring (%zu bytes)", bytes); return SEGB_E_UNSUPPORTED; }
    if (bytes > 48 * 1024) SEGB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

}  // namespace segb

using namespace segb;

extern "C" int segb_fixedvar_add_items(const segb_fixedvar *m, const int32_t *ids, const int32_t *ks, int32_t n,
                                       void *stream) {
    SEGB_CHECK_ARG(m && ids && ks && n >= 0, "null pointer");
    if (n == 0) return 0;
    fv_add_list_kernel<<<1, 256, sizeof(double) * m->D, (cudaStream_t)stream>>>(*m, ids, ks, n);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_fixedvar_del_items(const segb_fixedvar *m, const int32_t *ids, int32_t n,
                                       const int32_t *relabel_ids, int64_t relabel_n, void *stream) {
    SEGB_CHECK_ARG(m && ids && n >= 0, "null pointer");
    if (n == 0) return 0;
    fv_del_list_kernel<<<1, 256, sizeof(double) * m->D, (cudaStream_t)stream>>>(*m, ids, n, relabel_ids, relabel_n);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_fixedvar_log_pred_row(const segb_fixedvar *m, int32_t id, double *out, void *stream) {
    SEGB_CHECK_ARG(m && out && id >= 0 && id < m->n_emb, "item id");
    const size_t bytes = ScoreSmem::bytes(m->D, m->K_max);
    int r = score_smem_attr((const void *)fv_pred_row_kernel, bytes);
    if (r) return r;
    fv_pred_row_kernel<<<1, 256, bytes, (cudaStream_t)stream>>>(*m, id, out);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_fixedvar_log_marg(const segb_fixedvar *m, const int32_t *ids, const double *durs, int64_t n,
                                      double time_power_term, double wip, double *out, void *stream) {
    SEGB_CHECK_ARG(m && ids && out && n >= 0, "null pointer");
    if (n == 0) return 0;
    const size_t bytes = ScoreSmem::bytes(m->D, m->K_max);
    int r = score_smem_attr((const void *)fv_log_marg_kernel, bytes);
    if (r) return r;
    const int blocks = (int)(n < 148 * 16 ? n : 148 * 16);
    fv_log_marg_kernel<<<blocks, 256, bytes, (cudaStream_t)stream>>>(*m, ids, durs, n, time_power_term, wip, out);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_fixedvar_assign_items(const segb_fixedvar *m, const int32_t *ids, int32_t n, int32_t mode,
                                          double anneal_temp, const double *uniforms, int64_t *u_counter,
                                          int32_t *ks_out, void *stream) {
    SEGB_CHECK_ARG(m && ids && n >= 0, "null pointer");
    SEGB_CHECK_ARG(mode == 1 || (uniforms && u_counter), "sampling needs uniforms and a counter");
    if (n == 0) return 0;
    const size_t bytes = ScoreSmem::bytes(m->D, m->K_max);
    int r = score_smem_attr((const void *)fv_assign_list_kernel, bytes);
    if (r) return r;
    fv_assign_list_kernel<<<1, 1024, bytes, (cudaStream_t)stream>>>(*m, ids, n, mode, anneal_temp, uniforms, u_counter, ks_out);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_gibbs_sweep_fixedvar(const segb_fixedvar *m, const segb_corpus *c, const int32_t *h_order,
                                         int32_t n_order, int32_t fb_mode, double time_power_term, double wip,
                                         double anneal_temp, int32_t anneal_gibbs_am, const double *uniforms,
                                         int64_t *u_counter, double *scratch_scores, double *log_probs,
                                         int32_t *status, void *stream) {
    SEGB_CHECK_ARG(m && c && h_order && scratch_scores && log_probs && status, "null pointer");
    SEGB_CHECK_ARG(fb_mode == SEGB_DP_FFBS || fb_mode == SEGB_DP_VITERBI_GMM, "fb_mode");
    SEGB_CHECK_ARG(fb_mode == SEGB_DP_VITERBI_GMM || (uniforms && u_counter), "FFBS needs uniforms");
    SEGB_CHECK_ARG(c->tok_id && c->bounds, "corpus needs bounds and tok_id");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = ScoreSmem::bytes(m->D, m->K_max);
    int r;
    if ((r = score_smem_attr((const void *)fv_score_utt_kernel, bytes))) return r;
    if ((r = score_smem_attr((const void *)fv_assign_utt_kernel, bytes))) return r;
    const int score_blocks = c->N_max * c->S;
    const int assign_mode = (fb_mode == SEGB_DP_FFBS) ? 0 : 1;
    const double assign_temp = anneal_gibbs_am ? anneal_temp : 1.0;
    for (int i = 0; i < n_order; ++i) {
        const int u = h_order[i];
        SEGB_CHECK_ARG(u >= 0 && u < c->n_utt, "utterance index");
        fv_remove_utt_kernel<<<1, 256, sizeof(double) * m->D, st>>>(*m, *c, u);
        SEGB_LAUNCH_CHECK();
        fv_score_utt_kernel<<<score_blocks, 256, bytes, st>>>(*m, *c, u, time_power_term, wip, scratch_scores);
        SEGB_LAUNCH_CHECK();
        r = launch_dp_local(c, u, scratch_scores, fb_mode, 0.0, anneal_temp, uniforms, u_counter,
                            log_probs + i, status + i, st);
        if (r) return r;
        fv_assign_utt_kernel<<<1, 1024, bytes, st>>>(*m, *c, u, assign_mode, assign_temp, uniforms, u_counter, status + i);
        SEGB_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int segb_fixedvar_del_items_lm(const segb_fixedvar *m, const segb_bigram_lm *lm, const int32_t *ids, int32_t n,
                                          const int32_t *relabel_ids, int64_t relabel_n, void *stream) {
    SEGB_CHECK_ARG(m && lm && ids && n >= 0 && lm->K == m->K_max, "null pointer");
    if (n == 0) return 0;
    fv_del_list_lm_kernel<<<1, 256, sizeof(double) * m->D, (cudaStream_t)stream>>>(*m, *lm, ids, n, relabel_ids, relabel_n);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_bigram_lm_update(const segb_bigram_lm *lm, const int32_t *transcript, int32_t n, int32_t sign,
                                     void *stream) {
    SEGB_CHECK_ARG(lm && lm->unigram_counts && lm->bigram_counts && (sign == 1 || sign == -1) && n >= 0, "bigram lm");
    if (n == 0) return 0;
    SEGB_CHECK_ARG(transcript, "null pointer");
    bg_lm_update_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(*lm, transcript, n, sign);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_bigram_lm_log_prob_row(const segb_bigram_lm *lm, int32_t j_prev, double *out, void *stream) {
    SEGB_CHECK_ARG(lm && lm->unigram_counts && lm->bigram_counts && out && j_prev < lm->K, "bigram lm");
    bg_lm_row_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(*lm, j_prev, out);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_gibbs_sweep_bigram(const segb_fixedvar *m, const segb_bigram_lm *lm, const segb_corpus *c,
                                       const int32_t *h_order, int32_t n_order, int32_t assignments_only,
                                       double time_power_term, double wip, double anneal_temp,
                                       int32_t anneal_gibbs_am, const double *uniforms, int64_t *u_counter,
                                       double *scratch_scores, double *log_probs, int32_t *status, void *stream) {
    SEGB_CHECK_ARG(m && lm && c && h_order && scratch_scores && log_probs && status && uniforms && u_counter, "null pointer");
    SEGB_CHECK_ARG(lm->K == m->K_max && lm->unigram_counts && lm->bigram_counts, "the LM covers the K_max component labels");
    SEGB_CHECK_ARG(m->model == SEGB_MODEL_FIXEDVAR, "bigram sampling: fixed-variance components");
    SEGB_CHECK_ARG(c->tok_id && c->bounds, "corpus needs bounds and tok_id");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = ScoreSmem::bytes(m->D, m->K_max);
    int r;
    if ((r = score_smem_attr((const void *)fv_score_utt_kernel, bytes))) return r;
    if ((r = score_smem_attr((const void *)bg_assign_utt_kernel, bytes))) return r;
    const int score_blocks = c->N_max * c->S;
    const double assign_temp = anneal_gibbs_am ? anneal_temp : 1.0;
    for (int i = 0; i < n_order; ++i) {
        const int u = h_order[i];
        SEGB_CHECK_ARG(u >= 0 && u < c->n_utt, "utterance index");
        // remove_utt clears tok_id; when the boundaries are kept they are re-derived from `bounds` by the assign kernel
        bg_remove_utt_kernel<<<1, 256, sizeof(double) * m->D, st>>>(*m, *lm, *c, u);
        SEGB_LAUNCH_CHECK();
        if (!assignments_only) {
            fv_score_utt_kernel<<<score_blocks, 256, bytes, st>>>(*m, *c, u, time_power_term, wip, scratch_scores);
            SEGB_LAUNCH_CHECK();
            r = launch_dp_local(c, u, scratch_scores, SEGB_DP_FFBS, 0.0, anneal_temp, uniforms, u_counter,
                                log_probs + i, status + i, st);
            if (r) return r;
        } else {
            SEGB_CUDA(cudaMemsetAsync(log_probs + i, 0, sizeof(double), st));
            SEGB_CUDA(cudaMemsetAsync(status + i, 0, sizeof(int32_t), st));
        }
        bg_assign_utt_kernel<<<1, 1024, bytes, st>>>(*m, *lm, *c, u, assign_temp, uniforms, u_counter,
                                                     assignments_only ? nullptr : status + i);
        SEGB_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int segb_fixedvar_build(const segb_fixedvar *m, const int64_t *order, const int64_t *seg_off, int32_t K_new,
                                   void *stream) {
    SEGB_CHECK_ARG(m && order && seg_off && K_new >= 0 && K_new <= m->K_max, "null pointer / K");
    if (K_new == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = (int64_t)K_new * m->D;
    fv_build_stats_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(*m, order, seg_off, K_new);
    SEGB_LAUNCH_CHECK();
    fv_build_finish_kernel<<<K_new, 256, sizeof(double) * m->D, st>>>(*m, order, seg_off, K_new);
    SEGB_LAUNCH_CHECK();
    return 0;
}
