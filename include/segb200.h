/*
 * segb200.h -- C ABI of libsegb200.so: B200 (sm_100a) kernels for the
 * segmentalist hot path (score every candidate segment embedding against every
 * mixture component, then segment with dynamic programming).
 *
 * The reference (kamperh/segmentalist) has no FFI: its seams are duck-typed
 * Python objects plus one Cython module.  Each entry point below names the
 * reference code it replaces (paths relative to the reference's segmentalist/
 * directory).  Conventions:
 *   - every pointer is a DEVICE pointer unless the name starts with h_;
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on
 *     it unless stated;
 *   - return value: 0 ok, <0 bad argument (SEGB_E_*), >0 a cudaError_t;
 *     segb_last_error() gives the text;
 *   - no CPU fallback exists: without a CUDA device every call fails.
 *
 * Layouts
 *   X            [n_emb, D] row-major embeddings (components.X), float32 or float64
 *   banded slot  (pos_off[u] + t - 1) * S + (l - 1)  <->  segment [t-l, t) of
 *                utterance u, i.e. packed-triangular entry t(t-1)/2 + (t-l) of
 *                utterances.py:59-65; S = band width (longest stored span)
 *   bounds       [n_pos] uint8, bounds[pos_off[u] + j] = boundaries[u, j]
 */
#ifndef SEGB200_H
#define SEGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEGB_E_ARG        (-1)
#define SEGB_E_UNSUPPORTED (-2)
#define SEGB_E_NODEVICE   (-3)

/* per-utterance DP status codes written to `status` */
#define SEGB_DP_OK          0
#define SEGB_DP_INFEASIBLE  1  /* back-tracking reached t == 0 (reference: undefined negative index) */
#define SEGB_DP_EMPTY_SLICE 2  /* n_slices_min trimmed a window to nothing (reference: UB) */
#define SEGB_DP_NAN         3

#define SEGB_DP_FFBS           0  /* unigram_acoustic_wordseg.py:653-756 forward_backward */
#define SEGB_DP_VITERBI_GMM    1  /* unigram_acoustic_wordseg.py:759-864 forward_backward_viterbi */
#define SEGB_DP_VITERBI_KMEANS 2  /* kmeans_acoustic_wordseg.py:449-555 forward_backward_kmeans_viterbi */
/* OR-ed into `mode`: the caller vouches that `scores` hold no NaN and no +inf (-inf = unusable segment is fine),
 * e.g. because they come from segb_kmeans_band_scores over finite embeddings.  The batched k-means Viterbi
 * kernel then skips its per-candidate NaN compares (status never becomes SEGB_DP_NAN); other paths ignore it. */
#define SEGB_DP_SCORES_FINITE  0x100

const char *segb_last_error(void);
int segb_version(void);
/* number of kernel launches issued by this library since load (bench "gpu_launches") */
int64_t segb_launch_count(void);

/* ------------------------------------------------------------------ corpus */

/* Device view of Utterances (utterances.py:91-105) in banded form. */
typedef struct {
    int32_t n_utt;
    int32_t S;                 /* band width */
    int32_t N_max;             /* longest utterance (landmarks) */
    int32_t n_slices_min;
    int32_t n_slices_max;      /* 0 = unlimited, as in the reference */
    int64_t n_pos;             /* sum of utterance lengths */
    const int64_t *pos_off;    /* [n_utt + 1] */
    const int32_t *seg_id;     /* [n_pos * S] embedding id or -1 (vec_ids) */
    const double *seg_dur;     /* [n_pos * S] frames, NaN = unusable (durations) */
    uint8_t *bounds;           /* [n_pos] current boundaries */
    int32_t *tok_id;           /* [n_pos] embedding id of the token ENDING at this landmark, -1 if none */
} segb_corpus;

/* ------------------------------------------------------------------ DP (A5-A7) */

/* Batched segmentation DP over n_utt independent utterances (thread per utterance with
 * TMA-staged score blocks for batches; warp per utterance for long spans and single
 * utterances).  Replaces the three fb_func implementations named above.
 * scores: banded float64.  uniforms: utterance u reads uniforms[pos_off[u] + i]
 * for its i-th back-sampled segment, unless u_counter != NULL, in which case
 * the draws are taken from uniforms[*u_counter ...] and *u_counter is advanced
 * (sequential Gibbs; n_utt must be 1).  alphas (optional) receives
 * log_alphas / gammas.  utt_first selects a sub-range of the corpus.          */
int segb_dp_banded(const segb_corpus *c, int32_t utt_first, int32_t n_utt, const double *scores,
                   int32_t mode, double log_p_continue, double anneal_temp,
                   const double *uniforms, int64_t *u_counter,
                   uint8_t *bounds_out, double *log_prob, double *alphas,
                   int32_t *n_draws, int32_t *status, void *stream);

/* ------------------------------------------------------------------ fixed-variance FBGMM (A1-A4, A9) */

/* Device view of GaussianComponentsFixedVar (gaussian_components_fixedvar.py:80-120)
 * plus the FBGMM scalars (fbgmm.py:58-63).  The *T tables are [D, K_max]
 * (component index fastest) so a warp scoring 32 components reads contiguous
 * memory; mu_N = mu_N_num / prec_N is cached because log_post_pred (:247)
 * recomputes that quotient for every item.                                     */
typedef struct {
    int32_t D, K_max;
    int32_t x_is_f64;          /* dtype of X */
    int64_t n_emb;
    const void *X;
    double *mu_N_numT;         /* [D, K_max] mu_N_numerators^T */
    double *prec_NT;           /* [D, K_max] precision_Ns^T */
    double *prec_predT;        /* [D, K_max] precision_preds^T */
    double *mu_NT;             /* [D, K_max] mu_N_numerators / precision_Ns */
    double *log_prod_prec_pred;/* [K_max] */
    int32_t *counts;           /* [K_max] */
    int32_t *assignments;      /* [n_emb] */
    int32_t *K;                /* [1] active components */
    int64_t *n_total;          /* [1] sum(counts) */
    const double *precision;   /* [D] 1/var */
    const double *mu_0;        /* [D] */
    const double *precision_0; /* [D] 1/var_0 */
    double alpha, lms;
    double sum_log_precision_0;/* sum_d log(precision_0[d]) (_cython_utils.sum_log, :228) */
    /* Component model.  0: fixed variance (fields as named).  1: diagonal covariance with a
     * normal-inverse-chi-squared prior (gaussian_components_diag.py:82-345, niw.py) -- the same
     * kernels, with the tables reinterpreted:
     *   mu_N_numT -> m_N_numerators^T        prec_NT -> S_N_partials^T
     *   prec_predT -> inv_vars^T             mu_NT   -> m_N = m_N_numerators / (k_0 + n_k)
     *   log_prod_prec_pred -> log_prod_vars  mu_0 -> m_0   precision_0 -> S_0   precision: unused
     * and the predictive a product of Student's t densities (:237-259, :347-360).            */
    int32_t model;
    int32_t v_0;               /* model 1: prior degrees of freedom (integer, niw.py:12-14) */
    double k_0;                /* model 1: prior pseudo-count */
} segb_fixedvar;
#define SEGB_MODEL_FIXEDVAR 0
#define SEGB_MODEL_DIAG     1

/* add_item (:153-170) / del_item (:172-188, incl. del_component :190-221) for a
 * list of items, applied strictly in list order by one thread block.
 * ks[i] > K is clamped to K as FBGMM does before add_item (fbgmm.py:459-460).
 * relabel_ids/relabel_n: the set of item ids whose assignment may equal the
 * moved component (the reference scans all of `assignments`; pass NULL/0 for
 * that behaviour, or the corpus tok_id table to scan only live tokens).      */
int segb_fixedvar_add_items(const segb_fixedvar *m, const int32_t *ids, const int32_t *ks, int32_t n,
                            void *stream);
int segb_fixedvar_del_items(const segb_fixedvar *m, const int32_t *ids, int32_t n,
                            const int32_t *relabel_ids, int64_t relabel_n, void *stream);

/* The constructor's `for k: for i in where(assignments == k): add_item(i, k)` (:111-120) into an EMPTY
 * model, in parallel: every component replays its own members' additions in index order (same rounded
 * operations as add_item, bit-identical statistics).  order [n_emb] = item ids stably sorted by
 * assignment, seg_off [K_max + 1] = start of component k's members (unassigned items first);
 * components 0..K_new-1 must all be non-empty.                                                    */
int segb_fixedvar_build(const segb_fixedvar *m, const int64_t *order, const int64_t *seg_off, int32_t K_new,
                        void *stream);

/* log_post_pred(i) (:242-253) for k < K and log_prior(i) (:224-231) for one item:
 * out[0..K_max): active slots get the posterior predictive, the rest the prior. */
int segb_fixedvar_log_pred_row(const segb_fixedvar *m, int32_t id, double *out, void *stream);

/* FBGMM.log_marg_i (fbgmm.py:256-285) for n items, optionally followed by the
 * duration scaling of get_vec_embed_log_probs (unigram_acoustic_wordseg.py:474-511):
 *   out[i] = log_marg_i(ids[i]) * durs[i]**time_power_term + wip
 * with -inf where ids[i] == -1 or durs[i] is NaN.  durs == NULL -> raw log_marg_i. */
int segb_fixedvar_log_marg(const segb_fixedvar *m, const int32_t *ids, const double *durs, int64_t n,
                           double time_power_term, double wip, double *out, void *stream);

/* gibbs_sample_inside_loop_i (fbgmm.py:422-463; mode 0) or map_assign_i
 * (:465-494; mode 1) for a list of items in order; each assignment sees the
 * statistics left by the previous one.  Draws come from uniforms[*u_counter..]. */
int segb_fixedvar_assign_items(const segb_fixedvar *m, const int32_t *ids, int32_t n, int32_t mode,
                               double anneal_temp, const double *uniforms, int64_t *u_counter,
                               int32_t *ks_out, void *stream);

/* UnigramAcousticWordseg.gibbs_sample_i (unigram_acoustic_wordseg.py:252-360)
 * for the utterances listed in h_order (HOST array), strictly sequential:
 * remove the utterance's tokens, score every candidate segment, run the DP
 * (fb_mode SEGB_DP_FFBS or SEGB_DP_VITERBI_GMM), assign the new tokens.
 * scratch_scores: [N_max * S] doubles.  log_probs: [n_order] (device) receives
 * the per-utterance log_prob; status [n_order].  No host synchronisation.     */
int segb_gibbs_sweep_fixedvar(const segb_fixedvar *m, const segb_corpus *c, const int32_t *h_order,
                              int32_t n_order, int32_t fb_mode, double time_power_term, double wip,
                              double anneal_temp, int32_t anneal_gibbs_am,
                              const double *uniforms, int64_t *u_counter, double *scratch_scores,
                              double *log_probs, int32_t *status, void *stream);

/* The same sweep as ONE cooperative kernel launch: CTA b owns a contiguous range of components
 * in shared memory, steps are separated by grid barriers, the small replicated state (counts, K,
 * draws consumed) advances identically on every CTA (csrc/fixedvar_gibbs.cu).  d_order is a DEVICE
 * array; work: segb_gibbs_work_bytes() bytes of device scratch.  Returns SEGB_E_UNSUPPORTED when
 * the model does not fit the per-CTA shared memory (use segb_gibbs_sweep_fixedvar then).      */
int64_t segb_gibbs_work_bytes(int32_t K_max, int32_t N_max, int32_t S);
/* Replicas (sequential Gibbs does not shard; independent chains do): limit the CTAs of the following
 * cooperative sweeps (0 = one per SM).  R chains with n_sm / R CTAs each, launched on R streams, run side by
 * side; each chain keeps its own sequential order.  Process-wide setting.                               */
int segb_gibbs_set_max_ctas(int32_t max_ctas);
int segb_gibbs_sweep_fixedvar_coop(const segb_fixedvar *m, const segb_corpus *c, const int32_t *d_order,
                                   int32_t n_order, int32_t fb_mode, double time_power_term, double wip,
                                   double anneal_temp, int32_t anneal_gibbs_am,
                                   const double *uniforms, int64_t *u_counter, void *work,
                                   double *log_probs, int32_t *status, void *stream);

/* FBGMM.gibbs_sample inner loop (fbgmm.py:357-400) over the items listed in d_items (DEVICE
 * array, in order): cache the item's component, del_item, draw a component from
 * lms*log(alpha/K_max + n_k) + log_post_pred / log_prior (one uniform per item, from
 * uniforms[*u_counter ...]), then add_item -- or restore the cached statistics when the item
 * returns to its component and no component died.  Same cooperative kernel and `work` buffer
 * (segb_gibbs_work_bytes(K_max, 1, 1) bytes suffice) as segb_gibbs_sweep_fixedvar_coop.        */
int segb_fbgmm_gibbs_items_coop(const segb_fixedvar *m, const int32_t *d_items, int32_t n_items,
                                double anneal_temp, const double *uniforms, int64_t *u_counter,
                                void *work, void *stream);

/* ------------------------------------------------------------------ bigram LM + bigram cluster sampling (8f rank 4) */

/* Device view of BigramSmoothLM (bigram_lms.py:18-113): the smoothed, interpolated
 * maximum-likelihood bigram LM over the K = K_max component labels.  bigram_counts[j, i] =
 * count of label i following label j (bigram_lms.py:36-38).  The counts are TIED to the
 * acoustic model's components: when del_component moves the last component into slot k,
 * the LM rows/columns move with it (gaussian_components_fixedvar.py:205-208, :218-221).   */
typedef struct {
    int32_t K;
    double intrp_lambda, a, b;
    int32_t *unigram_counts;   /* [K] */
    int32_t *bigram_counts;    /* [K, K] row-major (j_prev, i_cur) */
} segb_bigram_lm;

/* segb_fixedvar_del_items with the LM tie: a component that empties takes its LM row, column
 * and unigram count along when the last component moves into its slot.                      */
int segb_fixedvar_del_items_lm(const segb_fixedvar *m, const segb_bigram_lm *lm, const int32_t *ids, int32_t n,
                               const int32_t *relabel_ids, int64_t relabel_n, void *stream);
/* counts_from_utterance (bigram_lms.py:98-105; sign = +1) / remove_counts_from_utterance
 * (:107-113; sign = -1) for a transcript of n labels (device array).                        */
int segb_bigram_lm_update(const segb_bigram_lm *lm, const int32_t *transcript, int32_t n, int32_t sign,
                          void *stream);
/* log_prob_vec_i (:64-69; j_prev < 0) or log(prob_vec_given_j(j_prev)) (:78-91): out[0..K).  */
int segb_bigram_lm_log_prob_row(const segb_bigram_lm *lm, int32_t j_prev, double *out, void *stream);

/* BigramAcousticWordseg.gibbs_sample_i (bigram_acoustic_wordseg.py:386-543, fb_type="unigram")
 * for the utterances listed in h_order (HOST array), strictly sequential and without host
 * synchronisation: remove the utterance's transcript from the LM and its tokens from the
 * components (with the LM tie), score every candidate segment with log_marg_i_embed_unigram
 * (:314-330; m->alpha must hold the LM's `a`, the counts are tied so the unigram term equals
 * FBGMM.log_marg_i's), FFBS (skipped when assignments_only), then sample the new tokens'
 * components left to right under the bigram prior (gibbs_sample_inside_loop_i_embed,
 * :333-384: log_prob_vec_i for the first token, log prob_vec_given_j(previous label) after)
 * and add the new transcript to the LM.  Buffers as segb_gibbs_sweep_fixedvar.                */
int segb_gibbs_sweep_bigram(const segb_fixedvar *m, const segb_bigram_lm *lm, const segb_corpus *c,
                            const int32_t *h_order, int32_t n_order, int32_t assignments_only,
                            double time_power_term, double wip, double anneal_temp, int32_t anneal_gibbs_am,
                            const double *uniforms, int64_t *u_counter, double *scratch_scores,
                            double *log_probs, int32_t *status, void *stream);

/* The same bigram sweep as ONE cooperative launch (csrc/fixedvar_gibbs.cu, as
 * segb_gibbs_sweep_fixedvar_coop): the owners form their slots' prior from the LM row of the
 * previous label, CTA 0 keeps the LM tables (transcript out / in, rows and columns following a
 * component that moves).  d_order is a DEVICE array; work: segb_gibbs_work_bytes() bytes.
 * Full sweeps only (assignments_only: use segb_gibbs_sweep_bigram); SEGB_E_UNSUPPORTED when the
 * model does not fit the per-CTA shared memory.                                                */
int segb_gibbs_sweep_bigram_coop(const segb_fixedvar *m, const segb_bigram_lm *lm, const segb_corpus *c,
                                 const int32_t *d_order, int32_t n_order, double time_power_term, double wip,
                                 double anneal_temp, int32_t anneal_gibbs_am, const double *uniforms,
                                 int64_t *u_counter, void *work, double *log_probs, int32_t *status, void *stream);

/* ------------------------------------------------------------------ k-means (A10-A12) */

/* Device view of KMeansComponents (kmeans_components.py:18-91). `means` has X's
 * dtype; meansT is its [D, K_max] transpose used by the scoring kernels.      */
typedef struct {
    int32_t D, K_max;
    int32_t x_is_f64;
    int64_t n_emb;
    const void *X;
    double *mean_num;          /* [K_max, D] mean_numerators */
    void *means;               /* [K_max, D] X dtype */
    void *meansT;              /* [D, K_max] X dtype */
    const void *random_means;  /* [K_max, D] X dtype */
    int32_t *counts;           /* [K_max] */
    int32_t *assignments;      /* [n_emb] */
    int32_t *K;                /* [1] */
} segb_kmeans;

/* neg_sqrd_norm(i) (kmeans_components.py:225-226) over all K_max rows, in X's
 * dtype and NumPy's pairwise summation order (bit-exact).                     */
int segb_kmeans_neg_sqrd_norm_row(const segb_kmeans *m, int32_t id, void *out, void *stream);

/* max / first argmax of neg_sqrd_norm for n items (:228-232).  ids[i] == -1 ->
 * best_val = -inf, best_k = -1.  best_val has X's dtype.                      */
int segb_kmeans_best(const segb_kmeans *m, const int32_t *ids, int64_t n, void *best_val,
                     int32_t *best_k, void *stream);

/* add_item (:93-111, with the k > K clamp) / del_item (:113-132) /
 * clean_components (:263-266), list order, one thread block.                  */
int segb_kmeans_add_items(const segb_kmeans *m, const int32_t *ids, const int32_t *ks, int32_t n,
                          void *stream);
int segb_kmeans_del_items(const segb_kmeans *m, const int32_t *ids, int32_t n, void *stream);
/* The constructor's add loop (:79-81) into an EMPTY model for all components in parallel, every component
 * summing its members in index order (bit-identical statistics); order / seg_off as in segb_fixedvar_build,
 * components 0..K_new-1 all non-empty.                                                                 */
int segb_kmeans_build(const segb_kmeans *m, const int64_t *order, const int64_t *seg_off, int32_t K_new,
                      void *stream);
int segb_kmeans_clean(const segb_kmeans *m, const int32_t *relabel_ids, int64_t relabel_n, void *stream);
/* KMeans.fit M-step (kmeans.py:149-151): del_item(ids[i]); add_item(ids[i], ks[i]) in list order. */
int segb_kmeans_move_items(const segb_kmeans *m, const int32_t *ids, const int32_t *ks, int32_t n, void *stream);

/* get_vec_embed_neg_len_sqrd_norms (kmeans_acoustic_wordseg.py:334-351) from
 * per-embedding best values: scores[slot] = (double)best_val[seg_id[slot]] *
 * seg_dur[slot] + wip, -inf where absent.  Covers landmark positions
 * [pos_first, pos_first + n_positions) (host-known: pos_off of the utterance range). */
int segb_kmeans_band_scores(const segb_kmeans *m, const segb_corpus *c, int64_t pos_first, int64_t n_positions,
                            const void *best_val, double wip, double *scores, void *stream);

/* SegmentalKMeansWordseg.segment_i (kmeans_acoustic_wordseg.py:225-332) for the
 * utterances in h_order (HOST array), strictly sequential with online mean
 * updates.  totals [n_order] receives sum_neg_len_sqrd_norm per utterance.    */
int segb_kmeans_segment_sweep(const segb_kmeans *m, const segb_corpus *c, const int32_t *h_order,
                              int32_t n_order, double wip, double *scratch_scores,
                              float *scratch_best, int32_t *scratch_arg,
                              double *totals, int32_t *status, void *stream);

/* Frozen-state sweep pieces (new batch mode; KMeans.fit semantics kmeans.py:124-171).
 * segb_kmeans_collect: for utterances [utt_first, +n_utt) turn `bounds` into
 * tokens (tok_id), set assignments[id] = best_k[id], and accumulate
 * sum_x[k] += X[id] (float64) and cnt[k] += 1 for the chosen segments.  sum_x /
 * cnt are the buffers an NCCL allreduce combines across ranks.
 * segb_kmeans_set_means: mean_num <- sum_x, counts <- cnt,
 * means[k] = (X dtype)(sum_x[k] / cnt[k]) where cnt[k] > 0 (:110).            */
int segb_kmeans_collect(const segb_kmeans *m, const segb_corpus *c, int32_t utt_first, int32_t n_utt,
                        const int32_t *best_k, double *sum_x, int64_t *cnt, void *stream);
int segb_kmeans_set_means(const segb_kmeans *m, const double *sum_x, const int64_t *cnt, void *stream);

/* ------------------------------------------------------------------ tensor-core scoring (tcgen05 + TMA) */

/* Sizes of the pre-tiled fp16 operand images used by the tcgen05 filter GEMM. */
int64_t segb_mma_x_tiles_bytes(int64_t n_emb, int32_t D);
int64_t segb_mma_w_tiles_bytes(int32_t K_max, int32_t D);
int64_t segb_mma_cand_bytes(int64_t n_emb);

/* Ingest: write the fp16 UMMA-canonical tile image of X (done once; X is constant across
 * sweeps).  x_err [2 * n_emb]: per row (|x - fp16(x)|, |x|); x_max [2]: their maxima -- the
 * inputs of the rigorous candidate threshold.                                                */
int segb_mma_pack_x(const float *X, int64_t n_emb, int32_t D, void *x_tiles, float *x_err, float *x_max,
                    void *stream);
/* Per sweep: tile image of the K_max float32 means, with -|mu|^2/2 folded into padding columns
 * of the inner dimension.  w_err [2 * (K_max + 128)]: per component (|mu - fp16(mu)|,
 * |fp16(mu)|); w_max [2]: their maxima.                                                      */
int segb_mma_pack_means(const float *means, int32_t K_max, int32_t D, void *w_tiles, float *w_err, float *w_max,
                        void *stream);
/* Filter GEMM (tcgen05.mma kind::f16, fp32 accumulate in TMEM, operands staged by
 * cp.async.bulk): for every embedding the best / second / third 16-component chunk of
 * x.mu - |mu|^2/2 and, for the best two, which members lie within the error bound of the
 * chunk maximum.  cand: opaque 32-byte per-row records for segb_mma_refine.                  */
int segb_mma_filter(const void *x_tiles, const void *w_tiles, int64_t n_emb, int32_t K_max, int32_t D,
                    const float *x_max, const float *w_max, void *cand, void *stream);
/* Refine: re-score the surviving candidates (normally one per embedding) exactly -- float32,
 * NumPy pairwise order, same bits as segb_kmeans_best -> bit-exact max / first argmax.
 * work: segb_mma_refine_work_bytes() bytes of scratch.  n_fallback (device, required; reset by
 * the call) counts rows whose third-best chunk was still inside the bound: those get an
 * exhaustive exact scan (second kernel, one block per such row).                             */
int64_t segb_mma_refine_work_bytes(int64_t n_emb, int32_t K_max);
int segb_mma_refine(const segb_kmeans *m, const void *cand, const float *x_err, const float *w_max,
                    int64_t n_emb, void *work, float *best_val, int32_t *best_k, int64_t *n_fallback,
                    void *stream);
/* Same refine with a SECOND-LEVEL tensor pass in front of the exhaustive scan (models with many
 * near-duplicate components -- a diffuse k-means state, K_true << K_max -- leave more than three chunks
 * inside the bound for many rows): the undecided rows' fp16 images are gathered into a compact tile
 * image, the filter GEMM runs over them again with a bitmap epilogue (every component whose filter
 * score reaches best - tau of the first pass), and only the flagged components are re-scored exactly.
 * Same bits out (kmeans_components.py:225-232).  x_tiles / w_tiles: the images segb_mma_filter read.
 * work: segb_mma_refine2_work_bytes() bytes: the undecided list is worked off in rounds of n_emb/8 rows
 * (a smaller buffer shortens the rounds; below one 256-row round the call falls back to
 * segb_mma_refine).  n_fallback counts the rows the top-3 records could not decide; only rows with an
 * empty bitmap (NaN scores) still take the exhaustive scan.  max_rounds > 0 limits the rounds that are
 * launched (a caller that knows the previous sweep's count saves the empty launches); undecided rows
 * beyond them take the exhaustive scan -- still exact, only slower.  0 = as many as n_emb may need.     */
int64_t segb_mma_refine2_work_bytes(int64_t n_emb, int32_t K_max, int32_t D);
int segb_mma_refine2(const segb_kmeans *m, const void *x_tiles, const void *w_tiles, const void *cand,
                     const float *x_err, const float *w_max, int64_t n_emb, void *work, int64_t work_bytes,
                     int32_t max_rounds, float *best_val, int32_t *best_k, int64_t *n_fallback, void *stream);

/* e4m3 FIRST-LEVEL filter for the k-means scorer (kind::f8f6f4: twice the MMA rate, half the operand bytes).
 * Same contract as segb_mma_*: a rigorous per-row bound on the filter's error decides which components can be the
 * reference's argmax (kmeans_components.py:225-232); the survivors are re-scored exactly, so max / first argmax
 * stay bit-identical.  The e4m3 bound is ~100x looser than the fp16 one: it decides rows whose best component is
 * well separated (a trained model); the rest go through an fp16 second-level pass (segb_mma8_refine), then the
 * exhaustive scan.  `scale`: a power of two that brings the operands into e4m3's normal range -- callers pick
 * the largest one with scale * max|x_d| <= 448 and scale * max|x| <= 448 (means are averages of embeddings).
 *   x_tiles8 / w_tiles8: e4m3 tile images (segb_mma8_*_tiles_bytes), x_err8 [2 n_emb], x_max8 [2],
 *   w_err8 [4 (K_max + 128)], w_max8 [4] = (e_mu, n_mu, e_bias, bias_max) in the scaled space.
 * segb_mma8_refine: w_tiles16 / w_max16 from segb_mma_pack_means (the second level runs in fp16 over a compact image
 * converted on the fly from the fp32 rows: no resident fp16 image of X); work: segb_mma_refine2_work_bytes().
 * The record's m1 / m2 / m3 are PACKED KEYS (chunk id in the low 12 mantissa bits of the chunk maximum; filter_tau8
 * carries the 2^-11 relative term): segb_mma8_filter returns SEGB_E_UNSUPPORTED for K_max > 65536 -- use segb_mma_*. */
int64_t segb_mma8_x_tiles_bytes(int64_t n_emb, int32_t D);
int64_t segb_mma8_w_tiles_bytes(int32_t K_max, int32_t D);
int segb_mma8_pack_x(const float *X, int64_t n_emb, int32_t D, float scale, void *x_tiles8, float *x_err8,
                     float *x_max8, void *stream);
int segb_mma8_pack_means(const float *means, int32_t K_max, int32_t D, float scale, void *w_tiles8, float *w_err8,
                         float *w_max8, void *stream);
int segb_mma8_filter(const void *x_tiles8, const void *w_tiles8, int64_t n_emb, int32_t K_max, int32_t D,
                     const float *x_max8, const float *w_max8, void *cand, void *stream);
int segb_mma8_refine(const segb_kmeans *m, const void *cand, const float *x_err8, const float *w_max8, float scale,
                     const void *w_tiles16, const float *w_max16, int64_t n_emb, void *work, int64_t work_bytes,
                     int32_t max_rounds, float *best_val, int32_t *best_k, int64_t *n_fallback, void *stream);

/* ------------------------------------------------------------------ tensor-core log_marg_i (fixed variance) */

/* FBGMM.log_marg_i (fbgmm.py:256-285) for ALL n_emb embeddings against the frozen model, as one
 * FP32-accurate tcgen05 GEMM (fp16 hi/lo split of both operands, three passes; norms, the
 * count-weighted log prior and all constants folded into a fourth K step) with an online
 * logsumexp over the K_max slots fused into the epilogue.  Isotropic prior.var / prior.var_0
 * only (caller checks).  Accuracy ~1e-6 relative (north star: 1e-4).
 * segb_fvmma_pack_x: fp16 split tile image of X (once).  segb_fvmma_log_marg: packs the model
 * image into w_tiles (K_max * D work) and runs the GEMM; out[i] = log_marg_i(i), float32.     */
int64_t segb_fvmma_x_tiles_bytes(int64_t n_emb, int32_t D);
int64_t segb_fvmma_w_tiles_bytes(int32_t K_max, int32_t D);
int segb_fvmma_pack_x(const float *X, int64_t n_emb, int32_t D, void *x_tiles, void *stream);
int segb_fvmma_log_marg(const segb_fixedvar *m, const void *x_tiles, void *w_tiles, int64_t n_emb, float *out,
                        void *stream);

/* ------------------------------------------------------------------ tensor-core log_marg_i, filter-and-refine (fixed variance) */

/* FBGMM.log_marg_i (fbgmm.py:256-285) for ALL n_emb embeddings against the frozen model at ONE tensor pass:
 * the filter GEMM of segb_mma_filter (fp16 operands, fp32 TMEM accumulators) computes
 * s^_k ~ s_k = lms*pi_k + log_post_pred_k(x) (log_prior for the empty slots, one virtual row) and keeps, per
 * embedding, the 16-component chunks that can hold a component within T nats of the best one (rigorous
 * rounding bound added); segb_fvf_refine re-scores those exactly in float64 (delta form, like
 * gaussian_components_fixedvar.py:247-252) and takes the logsumexp over the exact scores.  What is dropped
 * carries at most K_max*exp(-T) of the sum (T = 25: 7e-8).  aniso != 0: prior.var / prior.var_0 are D-vectors
 * (gaussian_components_fixedvar.py:95-99, tests/test_gaussian_components_fixedvar.py:51-53): inner dimension
 * [x, x*x] in two chunks.  Rows with a flat posterior (third-best chunk inside the threshold) get an
 * exhaustive exact scan; n_fallback counts them.
 *   segb_fvf_pack_x      once: fp16 tile image of X (+ per-row rounding-error norms x_err [2 n_emb], x_max [2])
 *   segb_fvf_pack_model  per model state: fp16 tile image of the model in w_tiles, the exact row tables and
 *                        per-row error norms in `model` (segb_fvf_model_bytes), model-wide maxima in w_max [4]
 *   segb_fvf_filter      the GEMM; cand = segb_mma_cand_bytes(n_emb) bytes of per-row records
 *   segb_fvf_refine      log_marg [n_emb] float64, map_k [n_emb] (optional): the MAP slot of map_assign_i
 *                        (fbgmm.py:465-494; the first empty slot is K); rec_out (optional, 16 bytes per row): the
 *                        thresholded row records segb_fvf_choose_tokens draws from; work = segb_fvf_work_bytes(n_emb) */
int64_t segb_fvf_x_tiles_bytes(int64_t n_emb, int32_t D, int32_t aniso);
int64_t segb_fvf_w_tiles_bytes(int32_t K_max, int32_t D, int32_t aniso);
int64_t segb_fvf_model_bytes(int32_t K_max, int32_t D, int32_t aniso);
int64_t segb_fvf_work_bytes(int64_t n_emb);
int segb_fvf_pack_x(const float *X, int64_t n_emb, int32_t D, int32_t aniso, void *x_tiles, float *x_err,
                    float *x_max, void *stream);
int segb_fvf_pack_model(const segb_fixedvar *m, int32_t aniso, void *w_tiles, void *model, float *w_max,
                        void *stream);
int segb_fvf_filter(const void *x_tiles, const void *w_tiles, int64_t n_emb, int32_t K_max, int32_t D,
                    int32_t aniso, const float *x_max, const float *w_max, float T, void *cand, void *stream);
int segb_fvf_refine(const float *X, int64_t n_emb, int32_t D, int32_t K_max, int32_t aniso, const void *model,
                    const void *cand, const float *x_err, const float *w_max, float T, void *work,
                    double *log_marg, int32_t *map_k, void *rec_out, int64_t *n_fallback, void *stream);

/* e4m3 FIRST LEVEL of the same filter (isotropic variances; kind::f8f6f4).  Operands are scaled by powers of two (sx
 * for the embeddings, alpha for |x|^2 -- callers pick the largest ones with sx * max|x_d| <= 448, sx * max|x| <= 448 and
 * alpha * max|x|^2 <= 448; the weight scale is chosen on the device from the model); constants ride in two-term e4m3
 * splits and dead model rows are pushed out of reach by sixteen +-448 columns.  The bound (lse_bound8, measured rounding
 * errors) is ~40 nats at p ~ 500, so with T = 20 the pass decides rows whose best component is > 100 nats ahead of the
 * fourth-best chunk -- a trained model; other rows take the exhaustive exact scan (n_fallback): callers watch that
 * count and go back to segb_fvf_* when it is not small.  Call order: segb_fvf_pack_model(aniso = 0) [exact row tables,
 * w_max16] -> segb_fvf8_pack_model -> segb_fvf8_filter -> segb_fvf8_refine.  log_marg_i / MAP slot / row records as
 * segb_fvf_refine.  w_err8: segb_fvf8_w_err_bytes(); w_max8 [8].  Packed top-3 keys as in segb_mma8_filter
 * (lse_bound8 carries their term): K_max + 1 <= 65536.                                                                  */
int64_t segb_fvf8_x_tiles_bytes(int64_t n_emb, int32_t D);
int64_t segb_fvf8_w_tiles_bytes(int32_t K_max, int32_t D);
int64_t segb_fvf8_w_err_bytes(int32_t K_max);
int segb_fvf8_pack_x(const float *X, int64_t n_emb, int32_t D, float sx, float alpha, void *x_tiles8, float *x_err8,
                     float *x_max8, void *stream);
int segb_fvf8_pack_model(int32_t K_max, int32_t D, const void *model, const float *w_max16, float sx, float alpha,
                         void *w_tiles8, float *w_err8, float *w_max8, void *stream);
int segb_fvf8_filter(const void *x_tiles8, const void *w_tiles8, int64_t n_emb, int32_t K_max, int32_t D,
                     const float *x_max8, const float *w_max8, float sx, float alpha, float T, void *cand, void *stream);
int segb_fvf8_refine(const float *X, int64_t n_emb, int32_t D, int32_t K_max, const void *model, const void *cand,
                     const float *x_err8, const float *w_max8, float sx, float alpha, float T, void *work,
                     double *log_marg, int32_t *map_k, void *rec_out, int64_t *n_fallback, void *stream);

/* ------------------------------------------------------------------ fused scoring: fp32 embeddings in, results out */

/* The same two scorers as ONE kernel that reads the fp32 embeddings once (csrc/score_fused.cu): four aux warps
 * convert the next 256 rows to the fp16 operand layout directly in shared memory (no resident fp16 image of
 * X, no pack pass) and re-score the previous 256 rows' surviving candidates exactly while the tensor pipe
 * works on the current ones; the per-row filter records never leave shared memory.
 *   segb_fused_kmeans_best  = segb_mma_filter + segb_mma_refine (w_tiles / w_max from segb_mma_pack_means);
 *                             rows [0, n_emb) of m->X.  Same bits out.  D even, <= 142 (else the two-kernel path).
 *   segb_fused_fv_log_marg  = segb_fvf_filter + segb_fvf_refine for isotropic variances (w_tiles / model / w_max
 *                             from segb_fvf_pack_model with aniso = 0).
 * work: (n_emb + 64) * 4 bytes (list of the rows left to the exhaustive scan); n_fallback (device) is reset.   */
int segb_fused_kmeans_best(const segb_kmeans *m, const void *w_tiles, const float *w_max, int64_t n_emb,
                           void *work, float *best_val, int32_t *best_k, int64_t *n_fallback, void *stream);
int segb_fused_fv_log_marg(const float *X, int64_t n_emb, int32_t D, int32_t K_max, const void *w_tiles,
                           const void *model, const float *w_max, float T, void *work, double *log_marg,
                           int32_t *map_k, void *rec_out, int64_t *n_fallback, void *stream);

/* get_vec_embed_log_probs (unigram_acoustic_wordseg.py:474-511) from per-embedding log marginals:
 * scores[slot] = log_marg[seg_id[slot]] * seg_dur[slot]**time_power_term + wip, -inf for absent slots /
 * NaN durations, over landmark positions [pos_first, pos_first + n_positions).                        */
int segb_fixedvar_band_scores(const segb_corpus *c, int64_t pos_first, int64_t n_positions,
                              const double *log_marg, double time_power_term, double wip, double *scores,
                              void *stream);

/* Frozen-model component choice for the tokens of the current boundaries (tok_id): mode 1 = map_assign_i
 * (fbgmm.py:465-494; choice[id] = map_k[id]); mode 0 = gibbs_sample_inside_loop_i (:422-463, anneal_temp 1):
 * inverse-CDF draw with uniforms[pos] over the exact probabilities of the slots the filter kept, in slot
 * order, empty slots last.  choice[id] is the raw slot index (>= K: an empty slot); add_item's clamp is
 * applied afterwards in token order (segb_frozen_new_list / segb_frozen_clamp).                        */
int segb_fvf_choose_tokens(const float *X, int32_t D, int32_t K_max, int32_t K, int32_t aniso, const void *model,
                           const void *recs, const segb_corpus *c, int64_t pos_first, int64_t n_positions,
                           int32_t mode, const int32_t *map_k, const double *uniforms, int32_t *choice,
                           void *stream);

/* ------------------------------------------------------------------ frozen-state model update (new batch mode) */

/* tok_id from bounds for utterances [utt_first, +n_utt) (utterances.py:159-174 in banded form). */
int segb_tokens_from_bounds(const segb_corpus *c, int32_t utt_first, int32_t n_utt, void *stream);

/* add_item's `k > K -> K`, `k == K` opens a component (kmeans_components.py:103-106, fbgmm.py:459-460) for a
 * whole sweep, in token order, without host logic:
 *   segb_frozen_new_list: ordered compaction of the tokens (positions [pos_first, +n_positions)) whose
 *     choice[id] >= K_before into list_j (the choices) / list_id (the embedding ids), n_list[0] = count
 *     (entries beyond `cap` are dropped; segb_frozen_clamp reports that);
 *   [ranks exchange list_j / n_list with a fixed-size all-gather: rank order is global token order]
 *   segb_frozen_clamp: ONE serial pass (a single warp) over the concatenated lists [world][cap]; the
 *     resolved components of this rank's tokens are written back into choice, the new K into K_out.   */
int64_t segb_frozen_new_work_bytes(int64_t n_positions);
int segb_frozen_new_list(const segb_corpus *c, int64_t pos_first, int64_t n_positions, const int32_t *choice,
                         int32_t K_before, void *work, int32_t cap, int32_t *list_j, int32_t *list_id,
                         int32_t *n_list, void *stream);
int segb_frozen_clamp(const int32_t *lists, const int32_t *counts, int32_t world, int32_t cap, int32_t my_rank,
                      int32_t K_before, int32_t K_max, const int32_t *list_id, int32_t *out_k, int32_t *choice,
                      int32_t *K_out, int32_t *overflow, void *stream);

/* FBGMM frozen update.  collect: sum_x[k] += X[id], cnt[k] += 1 over the tokens (k = choice[id]) -- the
 * buffers one all-reduce combines across ranks.  update: FBGMM.setup_components with the new assignments
 * (fbgmm.py:96-137): labels made consecutive in order (new_label [K_max] out), statistics of every component
 * from (sum_x, cnt) in closed form (gaussian_components_fixedvar.py:153-170, :317-325; equal to the
 * constructor's sequential add_item sums up to float64 rounding), assignments of the tokens, K, n_total. */
int segb_fixedvar_frozen_collect(const segb_fixedvar *m, const segb_corpus *c, int64_t pos_first,
                                 int64_t n_positions, const int32_t *choice, double *sum_x, int64_t *cnt,
                                 void *stream);
int segb_fixedvar_frozen_update(const segb_fixedvar *m, const segb_corpus *c, int64_t pos_first,
                                int64_t n_positions, const int32_t *choice, const double *sum_x,
                                const int64_t *cnt, int32_t *new_label, void *stream);

/* clean_components (kmeans_components.py:263-266) after segb_kmeans_set_means: the swap-with-last deletions
 * of the components with cnt == 0 among the first *m->K are replayed on the counts, rows gathered, inactive
 * slots reset to their random rows, live tokens relabelled, *m->K updated.  No host round trip.        */
int64_t segb_kmeans_frozen_clean_work_bytes(int32_t K_max, int32_t D);
int segb_kmeans_frozen_clean(const segb_kmeans *m, const segb_corpus *c, int64_t pos_first, int64_t n_positions,
                             const int64_t *cnt, void *work, void *stream);

/* ------------------------------------------------------------------ per-iteration diagnostics */

/* Both calls take the items grouped by component: `order` [n_emb] = item ids stably sorted by
 * assignment (members keep their index order, like np.where), seg_off [K_max + 1] = start of
 * component k's members in `order` (unassigned items, -1, come first).
 *
 * GaussianComponentsFixedVar.log_marg_k (gaussian_components_fixedvar.py:261-283) for every
 * component: X[members].sum(axis=0) and np.square(X[members]).sum(axis=0) are formed in X's dtype
 * with NumPy's row-after-row order, the closed form in float64 with separately rounded operations
 * and NumPy's pairwise np.sum over the D terms.  out_k [K_max]: log_marg_k(k) for k < K, 0 above;
 * log_marg() (:285-296) is their sum in component order (host).
 * work: segb_fixedvar_log_marg_k_work_bytes() bytes of scratch.                                */
int64_t segb_fixedvar_log_marg_k_work_bytes(int32_t K_max, int32_t D);
int segb_fixedvar_log_marg_k(const segb_fixedvar *m, const int64_t *order, const int64_t *seg_off,
                             void *work, double *out_k, void *stream);

/* KMeansComponents.sum_neg_sqrd_norm (kmeans_components.py:234-247), per component:
 * out_k[k] = -np.sum(deltas*deltas), deltas = mean_numerators[k]/counts[k] - X[members] (float64,
 * NumPy's pairwise order over the flattened array); the objective is the sum over k < K (host). */
int segb_kmeans_sum_neg_sqrd_norm_k(const segb_kmeans *m, const int64_t *order, const int64_t *seg_off,
                                    double *out_k, void *stream);

/* GaussianComponentsDiag.log_marg_k (gaussian_components_diag.py:271-288) for every component, from the
 * device-resident sufficient statistics (no member gathering: the closed form needs only counts,
 * m_N_numerators and S_N_partials); the two np.log(..).sum() in NumPy's pairwise order.
 * out_k [K_max]: log_marg_k(k) for k < K, 0 above; log_marg() (:290-301) is their sum in component
 * order (host).  m->model must be SEGB_MODEL_DIAG.                                               */
int segb_diag_log_marg_k(const segb_fixedvar *m, double *out_k, void *stream);

/* ------------------------------------------------------------------ ingestion (host) */

/* The random boundary initialisation of Utterances.__init__ (utterances.py:136-157) for all utterances in
 * order, HOST arrays, no device work: per utterance `rand(N) < p` is redrawn until the segmentation carries
 * an embedding and respects the span limits.  uniforms: a block drawn in advance from np.random (consumed
 * sequentially = the reference's stream).  Returns the number consumed, or -1 if the block ran out.     */
int64_t segb_host_init_boundaries(const int64_t *lengths, int64_t n_utt, const int64_t *packed_off,
                                  const int64_t *ids, const int64_t *pos_off, const double *uniforms,
                                  int64_t n_uniforms, double p_boundary, int64_t n_slices_min,
                                  int64_t n_slices_max, uint8_t *bounds_out);

/* ------------------------------------------------------------------ development aids */

/* Per-phase clock totals of CTA 0 of the cooperative Gibbs sweep (tools/gibbs_phases.py).
 * enable != 0: allocate / zero 16 device counters; the following segb_gibbs_sweep_fixedvar_coop
 * launches accumulate into them.  out16 (HOST array, may be NULL) receives the current totals
 * before they are zeroed.  enable == 0 releases the counters.  Synchronous.                    */
int segb_debug_gibbs_prof(unsigned long long *out16, int enable);

/* Test aid: start value of the cooperative sweeps' grid-barrier arrival counter (default 0).  The counter
 * is 32 bits wide and wraps during long launches by design (every CTA compares its own target modulo
 * 2^32); a base just below 2^32 forces the wrap after a few barriers.  Applies to the following
 * segb_*_coop launches.                                                                          */
int segb_debug_gibbs_bar_base(uint32_t base);

#ifdef __cplusplus
}
#endif
#endif /* SEGB200_H */
