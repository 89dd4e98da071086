/*
 * oracle/seg_oracle_c.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the scalar/native pieces of the segmentalist hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product path
 * (segmentalist_b200/) never links or calls it.
 *
 * Every function names the reference code it restates (paths relative to
 * /root/reference).  Build: gcc -O2 -ffp-contract=off -shared -fPIC (see
 * oracle/Makefile).  -ffp-contract=off keeps a*b+c as two rounded operations,
 * which is what the reference's Cython/NumPy code executes.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define ORC_OK 0
#define ORC_ERR_INFEASIBLE 1   /* back-tracking reached t == 0 (reference: undefined negative index) */
#define ORC_ERR_EMPTY_SLICE 2  /* n_slices_min trimmed a window to nothing (reference: UB in Cython) */
#define ORC_ERR_NAN 3

/* segmentalist/_cython_utils.pyx:13-25 -- two-pass logsumexp, sequential sum. */
double orc_logsumexp(const double *a, int n)
{
    double max_a = a[0];
    double sum_exps = 0.0;
    int j;
    for (j = 1; j < n; ++j)
        if (a[j] > max_a) max_a = a[j];
    for (j = 0; j < n; ++j)
        sum_exps += exp(a[j] - max_a);
    return log(sum_exps) + max_a;
}

/* segmentalist/_cython_utils.pyx:52-59 */
double orc_sum_log(const double *y, int n)
{
    double s = log(y[0]);
    int i;
    for (i = 1; i < n; ++i) s += log(y[i]);
    return s;
}

/* segmentalist/_cython_utils.pyx:63-70 */
double orc_sum_square_a_times_b(const double *a, const double *b, int n)
{
    double s = 0.0;
    int i;
    for (i = 0; i < n; ++i) s += a[i] * a[i] * b[i];
    return s;
}

/* segmentalist/_cython_utils.pyx:75-89 and segmentalist/utils.py:10-21.
 * The uniform is supplied by the caller (the reference reads random.random()). */
int orc_draw(const double *p, int n, double u)
{
    int i;
    for (i = 0; i < n; ++i) {
        u = u - p[i];
        if (u < 0) return i;
    }
    return n - 1;
}

/* Window helper shared by the three DPs: the reference slices
 *   v[i:i+t][-S:cut] + a[:t][-S:cut]
 * (unigram_acoustic_wordseg.py:693-699,713-716).  With S == 0 the slice is the
 * whole prefix.  lo = first landmark index j in the window, hi = one past the
 * last after the n_slices_min cut.                                            */
static void window(int t, int S, int n_min, int *lo, int *hi_full, int *hi_cut)
{
    int w = (S == 0 || S > t) ? t : S;
    *lo = t - w;
    *hi_full = t;
    *hi_cut = (n_min > 1) ? t - (n_min - 1) : t;   /* python slice [-S : -(n_min-1)] */
    if (*hi_cut < *lo) *hi_cut = *lo;
}

static int all_neg_inf(const double *vec, const double *a, int base, int lo, int hi)
{
    int j;
    for (j = lo; j < hi; ++j)
        if (vec[base + j] + a[j] != -INFINITY) return 0;
    return 1;
}

/*
 * mode 0: forward_backward            (unigram_acoustic_wordseg.py:653-756)  FFBS
 * mode 1: forward_backward_viterbi    (unigram_acoustic_wordseg.py:759-864)
 * mode 2: forward_backward_kmeans_viterbi (kmeans_acoustic_wordseg.py:449-555)
 *
 * vec      packed-triangular scores, length N(N+1)/2, entry t(t-1)/2+j = segment [j,t)
 * uniforms consumed left to right, one per back-sampled segment (mode 0 only)
 * alphas   out, length N (log_alphas / gammas; alphas[0] = 0)
 * bounds   out, length N (0/1)
 * returns ORC_* ; *n_used = number of uniforms consumed.
 */
int orc_dp_packed(const double *vec, int N, int n_min, int S, int mode,
                  double log_p_continue, double anneal_temp,
                  const double *uniforms, int *n_used,
                  double *alphas, uint8_t *bounds, double *log_prob_out)
{
    int t, j, lo, hi_full, hi_cut, base, used = 0, status = ORC_OK;
    double total = 0.0;
    double c[4096], p[4096];

    for (j = 0; j < N; ++j) { bounds[j] = 0; alphas[j] = 1.0; }
    bounds[N - 1] = 1;
    alphas[0] = 0.0;

    /* forward pass */
    base = 0;
    for (t = 1; t < N; ++t) {
        window(t, S, n_min, &lo, &hi_full, &hi_cut);
        if (all_neg_inf(vec, alphas, base, lo, hi_full)) {
            alphas[t] = -INFINITY;
        } else {
            int n = hi_cut - lo;
            if (n <= 0) return ORC_ERR_EMPTY_SLICE;
            for (j = 0; j < n; ++j) c[j] = vec[base + lo + j] + alphas[lo + j];
            if (mode == 0) {
                alphas[t] = orc_logsumexp(c, n) + log_p_continue;
            } else {
                double m = c[0];
                for (j = 1; j < n; ++j) if (c[j] > m) m = c[j];   /* np.max */
                alphas[t] = m;
            }
        }
        base += t;
    }

    /* backward pass */
    t = N;
    for (;;) {
        int n, k, all_inf;
        base = (t - 1) * t / 2;
        window(t, S, n_min, &lo, &hi_full, &hi_cut);
        n = hi_cut - lo;
        if (n <= 0) return ORC_ERR_EMPTY_SLICE;
        for (j = 0; j < n; ++j) c[j] = vec[base + lo + j] + alphas[lo + j];
        all_inf = 1;
        for (j = 0; j < n; ++j) { if (isnan(c[j])) return ORC_ERR_NAN; if (c[j] != -INFINITY) all_inf = 0; }
        if (all_inf) {
            /* walk left until something is feasible; the re-computed window is
             * NOT cut by n_slices_min (unigram_acoustic_wordseg.py:723-730) */
            while (all_inf) {
                t = t - 1;
                if (t == 0) return ORC_ERR_INFEASIBLE;
                base = (t - 1) * t / 2;
                window(t, S, 0, &lo, &hi_full, &hi_cut);
                n = hi_full - lo;
                for (j = 0; j < n; ++j) {
                    c[j] = vec[base + lo + j] + alphas[lo + j];
                    if (c[j] != -INFINITY) all_inf = 0;
                }
            }
            bounds[t - 1] = 1;
        }
        if (mode == 2) {
            /* argmax of reversed raw scores, first max wins (kmeans_acoustic_wordseg.py:535-536) */
            int best = 0;
            for (j = 1; j < n; ++j) if (c[n - 1 - j] > c[n - 1 - best]) best = j;
            k = best + 1;
        } else {
            double lse = orc_logsumexp(c, n);
            if (mode == 0 && anneal_temp != 1.0) {
                /* unigram_acoustic_wordseg.py:731-736 */
                double q[4096], lse2;
                for (j = 0; j < n; ++j) q[j] = 1. / anneal_temp * (c[n - 1 - j] - lse);
                lse2 = orc_logsumexp(q, n);
                for (j = 0; j < n; ++j) p[j] = exp(q[j] - lse2);
            } else {
                for (j = 0; j < n; ++j) p[j] = exp(c[n - 1 - j] - lse);
            }
            if (mode == 0) {
                k = orc_draw(p, n, uniforms[used]) + 1;
                used++;
            } else {
                int best = 0;   /* np.argmax(p_k): first max; NaN handling not needed (guarded above) */
                for (j = 1; j < n; ++j) if (p[j] > p[best]) best = j;
                k = best + 1;
            }
        }
        if (n_min > 1) k += n_min - 1;
        if (t - k < 0) return ORC_ERR_EMPTY_SLICE;
        total += vec[base + t - k];
        if (t - k - 1 < 0) break;
        bounds[t - k - 1] = 1;
        t = t - k;
    }
    *n_used = used;
    *log_prob_out = total;
    return status;
}

/* NumPy float32 pairwise row sum, as executed by (deltas*deltas).sum(axis=1) on a
 * C-contiguous float32 [K, D] array (numpy/_core/src/umath/loops_utils.h.src
 * pairwise_sum: blocks of <= 128 use 8 running accumulators combined as
 * ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); longer inputs split at n/2 rounded down
 * to a multiple of 8).  Checked against numpy in tests/test_oracle_numpy_order.py. */
static float pairwise_f32(const float *a, int n)
{
    if (n < 8) {
        int i;
        float res = 0.f;
        for (i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        float r[8], res;
        int i, j;
        for (j = 0; j < 8; ++j) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (j = 0; j < 8; ++j) r[j] += a[i + j];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_f32(a, n2) + pairwise_f32(a + n2, n - n2);
    }
}

/* kmeans_components.py:225-226 for one item against all K_max float32 means:
 * out[k] = -sum_d (means[k,d]-x[d])^2, float32 arithmetic in NumPy's order.   */
void orc_kmeans_neg_sqrd_norm_f32(const float *means, const float *x, int K, int D, float *out)
{
    float tmp[4096];
    int k, d;
    for (k = 0; k < K; ++k) {
        const float *m = means + (size_t)k * D;
        for (d = 0; d < D; ++d) { float dl = m[d] - x[d]; tmp[d] = dl * dl; }
        out[k] = -pairwise_f32(tmp, D);
    }
}

/* max / first-argmax over orc_kmeans_neg_sqrd_norm_f32 for a batch of items
 * (kmeans_components.py:228-232).                                              */
void orc_kmeans_best_f32(const float *means, const float *X, const int64_t *ids, int n_ids,
                         int K, int D, float *best_val, int32_t *best_k)
{
    float tmp[4096];
    int i, k, d;
    for (i = 0; i < n_ids; ++i) {
        const float *x = X + (size_t)ids[i] * D;
        float bv = 0.f; int bk = -1;
        for (k = 0; k < K; ++k) {
            const float *m = means + (size_t)k * D;
            float v;
            for (d = 0; d < D; ++d) { float dl = m[d] - x[d]; tmp[d] = dl * dl; }
            v = -pairwise_f32(tmp, D);
            if (bk < 0 || v > bv) { bv = v; bk = k; }
        }
        best_val[i] = bv; best_k[i] = bk;
    }
}
