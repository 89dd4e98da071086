// Per-iteration diagnostics on the device.
//
// The reference recomputes, after every sweep, quantities that gather each component's members with
// `np.where(assignments == k)` -- O(N * K) on the host, and in a device build a full copy of X back
// to the host:
//   GaussianComponentsFixedVar.log_marg_k / log_marg   gaussian_components_fixedvar.py:261-296
//   KMeansComponents.sum_neg_sqrd_norm                 kmeans_components.py:234-247
// Here the caller supplies the items grouped by component (a stable sort of `assignments`, so members
// keep their index order like np.where) and the kernels reproduce NumPy's arithmetic: per-column
// sequential sums in X's dtype for `.sum(axis=0)`, float64 elementwise expressions with separately
// rounded operations, and NumPy's pairwise order for the final np.sum.  Results are per component;
// the host adds them in component order like the reference's loop.
#include "common.cuh"

namespace segb {

// s1[k,d] = X[members].sum(axis=0)[d], s2[k,d] = np.square(X[members]).sum(axis=0)[d] in X's dtype
// (NumPy adds the rows one after the other), widened to float64 on store.  Thread per (k, d).
template <typename T>
__global__ void fv_moments_kernel(segb_fixedvar m, const int64_t *order, const int64_t *seg_off, double *s1, double *s2) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)m.K_max * m.D) return;
    const int k = (int)(idx / m.D), d = (int)(idx % m.D);
    const T *X = (const T *)m.X;
    T a = T(0), b = T(0);
    for (int64_t j = seg_off[k]; j < seg_off[k + 1]; ++j) {
        const T x = X[order[j] * m.D + d];
        a = add_rn<T>(a, x);
        b = add_rn<T>(b, mul_rn<T>(x, x));
    }
    s1[idx] = (double)a;
    s2[idx] = (double)b;
}

// log_marg_k (:261-283): thread per component, np.sum over the D per-dimension terms.
__global__ void fv_log_marg_k_kernel(segb_fixedvar m, const int64_t *seg_off, const double *s1, const double *s2,
                                     double *out_k) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m.K_max) return;
    if (k >= *m.K) { out_k[k] = 0.0; return; }
    const double N = (double)m.counts[k];
    const double half_nm1 = (N - 1.0) / 2.0;
    const double b = __dmul_rn(__dmul_rn(0.5, N), log(2.0 * CUDART_PI));
    const bool f32 = !m.x_is_f64;
    out_k[k] = pairwise_sum<double>([&](int d) {
        const double pr = m.precision[d], p0 = m.precision_0[d], mu0 = m.mu_0[d];
        const double v1 = s1[(size_t)k * m.D + d], v2 = s2[(size_t)k * m.D + d];
        // np.square(X.sum(axis=0)) stays in X's dtype
        const double v1sq = f32 ? (double)__fmul_rn((float)v1, (float)v1) : __dmul_rn(v1, v1);
        const double den = __dadd_rn(__ddiv_rn(N, p0), __ddiv_rn(1.0, pr));
        const double mu0sq = __dmul_rn(mu0, mu0);
        const double a = __dmul_rn(half_nm1, log(pr));
        const double c = __dmul_rn(0.5, log(den));
        const double e = __dmul_rn(__dmul_rn(0.5, pr), v2);
        const double f = __dmul_rn(__dmul_rn(0.5, p0), mu0sq);
        const double g1 = __ddiv_rn(__dmul_rn(v1sq, pr), p0);
        const double g2 = __ddiv_rn(__dmul_rn(mu0sq, p0), pr);
        const double g3 = __dmul_rn(__dmul_rn(2.0, v1), mu0);
        const double h = __ddiv_rn(__dmul_rn(0.5, __dadd_rn(__dadd_rn(g1, g2), g3)), den);
        return __dadd_rn(__dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(a, b), c), e), f), h);
    }, m.D);
}

// out_k[k] = -np.sum(deltas * deltas), deltas = mean_numerators[k]/counts[k] - X[members]  (float64,
// NumPy's pairwise order over the flattened [n_k, D] array).  Thread per component.
template <typename T>
__global__ void km_objective_kernel(segb_kmeans m, const int64_t *order, const int64_t *seg_off, double *out_k) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m.K_max) return;
    if (k >= *m.K) { out_k[k] = 0.0; return; }
    const T *X = (const T *)m.X;
    const double cnt = (double)m.counts[k];
    const int64_t lo = seg_off[k];
    const int64_t n = (seg_off[k + 1] - lo) * m.D;
    const int D = m.D;
    const double s = pairwise_sum<double>([&](int i) {
        const int r = i / D, d = i % D;
        const double dl = __dsub_rn(__ddiv_rn(m.mean_num[(size_t)k * D + d], cnt), (double)X[order[lo + r] * D + d]);
        return __dmul_rn(dl, dl);
    }, (int)n);
    out_k[k] = -s;
}

// GaussianComponentsDiag.log_marg_k (gaussian_components_diag.py:271-288) from the device tables: thread per
// component, the two np.log(..).sum() in NumPy's pairwise order, separately rounded elementwise operations.
// Tables are [D, K_max]: mu_N_numT = m_N_numerators^T, prec_NT = S_N_partials^T; precision_0 = S_0.
__global__ void diag_log_marg_k_kernel(segb_fixedvar m, double *out_k) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m.K_max) return;
    if (k >= *m.K) { out_k[k] = 0.0; return; }
    const int D = m.D, KM = m.K_max;
    const double n = (double)m.counts[k];
    const double k_N = __dadd_rn(m.k_0, n);
    const double v_0 = (double)m.v_0, v_N = __dadd_rn(v_0, n);
    const double log_pi = log(CUDART_PI);
    const double sum_log_S0 = pairwise_sum<double>([&](int d) { return log(m.precision_0[d]); }, D);
    const double sum_log_SN = pairwise_sum<double>([&](int d) {
        const double m_N = __ddiv_rn(m.mu_N_numT[(size_t)d * KM + k], k_N);
        const double S_N = __dsub_rn(m.prec_NT[(size_t)d * KM + k], __dmul_rn(k_N, __dmul_rn(m_N, m_N)));
        return log(S_N);
    }, D);
    const double Dd = (double)D;
    // - n*D/2*log_pi + D/2*log(k_0) - D/2*log(k_N) + v_0/2*sum(log S_0) - v_N/2*sum(log S_N) + D*(gammaln(v_N/2) - gammaln(v_0/2))
    double r = -__dmul_rn(__ddiv_rn(__dmul_rn(n, Dd), 2.0), log_pi);
    r = __dadd_rn(r, __dmul_rn(__ddiv_rn(Dd, 2.0), log(m.k_0)));
    r = __dsub_rn(r, __dmul_rn(__ddiv_rn(Dd, 2.0), log(k_N)));
    r = __dadd_rn(r, __dmul_rn(__ddiv_rn(v_0, 2.0), sum_log_S0));
    r = __dsub_rn(r, __dmul_rn(__ddiv_rn(v_N, 2.0), sum_log_SN));
    r = __dadd_rn(r, __dmul_rn(Dd, __dsub_rn(lgamma(__ddiv_rn(v_N, 2.0)), lgamma(__ddiv_rn(v_0, 2.0)))));
    out_k[k] = r;
}

}  // namespace segb

using namespace segb;

extern "C" int segb_diag_log_marg_k(const segb_fixedvar *m, double *out_k, void *stream) {
    SEGB_CHECK_ARG(m && out_k, "null pointer");
    SEGB_CHECK_ARG(m->model == SEGB_MODEL_DIAG, "segb_diag_log_marg_k: diagonal-covariance components only");
    diag_log_marg_k_kernel<<<(m->K_max + 63) / 64, 64, 0, (cudaStream_t)stream>>>(*m, out_k);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t segb_fixedvar_log_marg_k_work_bytes(int32_t K_max, int32_t D) { return (int64_t)2 * K_max * D * 8; }

extern "C" int segb_fixedvar_log_marg_k(const segb_fixedvar *m, const int64_t *order, const int64_t *seg_off,
                                        void *work, double *out_k, void *stream) {
    SEGB_CHECK_ARG(m && order && seg_off && work && out_k, "null pointer");
    SEGB_CHECK_ARG(m->model == SEGB_MODEL_FIXEDVAR, "log_marg_k: fixed-variance components only");
    cudaStream_t st = (cudaStream_t)stream;
    double *s1 = (double *)work, *s2 = s1 + (size_t)m->K_max * m->D;
    const int64_t n = (int64_t)m->K_max * m->D;
    if (m->x_is_f64) fv_moments_kernel<double><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(*m, order, seg_off, s1, s2);
    else fv_moments_kernel<float><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(*m, order, seg_off, s1, s2);
    SEGB_LAUNCH_CHECK();
    fv_log_marg_k_kernel<<<(m->K_max + 63) / 64, 64, 0, st>>>(*m, seg_off, s1, s2, out_k);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_kmeans_sum_neg_sqrd_norm_k(const segb_kmeans *m, const int64_t *order, const int64_t *seg_off,
                                               double *out_k, void *stream) {
    SEGB_CHECK_ARG(m && order && seg_off && out_k, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (m->x_is_f64) km_objective_kernel<double><<<(m->K_max + 31) / 32, 32, 0, st>>>(*m, order, seg_off, out_k);
    else km_objective_kernel<float><<<(m->K_max + 31) / 32, 32, 0, st>>>(*m, order, seg_off, out_k);
    SEGB_LAUNCH_CHECK();
    return 0;
}
