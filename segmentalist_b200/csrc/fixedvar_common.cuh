// Pieces of the fixed-variance FBGMM kernels shared by fixedvar.cu (one launch per step) and
// fixedvar_gibbs.cu (persistent cooperative sweep).
#pragma once
#include "common.cuh"

namespace segb {

__device__ __forceinline__ double fv_x(const segb_fixedvar &m, int64_t id, int d) {
    return m.x_is_f64 ? ((const double *)m.X)[id * m.D + d] : (double)((const float *)m.X)[id * m.D + d];
}

// -0.5*D*log(2*pi) exactly as the reference forms it (:123): one product of three factors
__device__ __forceinline__ double fv_norm_const(int D) { return -0.5 * D * log(2. * 3.14159265358979323846); }

// x*x exactly as the reference caches it for the diagonal model: np.square(X) keeps X's dtype
// (gaussian_components_diag.py:122-123), so float32 embeddings are squared in float32.
__device__ __forceinline__ double fv_xsq(const segb_fixedvar &m, int64_t id, int d) {
    if (m.x_is_f64) { const double v = ((const double *)m.X)[id * m.D + d]; return __dmul_rn(v, v); }
    const float v = ((const float *)m.X)[id * m.D + d];
    return (double)__fmul_rn(v, v);
}

// Diagonal model: per-component constants of the Student's t predictive for n members
// (:237-259): D*(gammaln((v+1)/2) - gammaln(v/2) - log(v)/2 - log(pi)/2), (v+1)/2 and 1/v, v = v_0+n.
__device__ __forceinline__ void diag_consts(const segb_fixedvar &m, int n, double &cst, double &hv, double &iv) {
    const double v = (double)(m.v_0 + n);
    cst = m.D * (lgamma((v + 1.) / 2.) - lgamma(v / 2.) - 0.5 * log(v) - 0.5 * log(3.14159265358979323846));
    hv = (v + 1.) / 2.;
    iv = 1. / v;
}

// Shared-memory layout for the scoring kernels: xs[D] | red[40] | sk[K_max]
struct ScoreSmem {
    double *xs, *red, *sk;
    __device__ ScoreSmem(double *base, int D) : xs(base), red(base + D), sk(base + D + 40) {}
    static size_t bytes(int D, int K_max) { return sizeof(double) * ((size_t)D + 40 + (K_max > D ? K_max : D)); }
};

// Second half of gibbs_sample_inside_loop_i / map_assign_i (fbgmm.py:441-463 / :480-494): given
// the unnormalised log-probabilities of all K_max physical slots in s.sk (active slots, then the
// identical empty ones) and their maximum over this thread's slots in `mx`, normalise, anneal,
// and draw (mode 0, one uniform `u`) or take the first maximum (mode 1).  Block-uniform result;
// deterministic for a given blockDim, so replicas that run it on the same inputs agree.
static __device__ int fv_decide(ScoreSmem &s, int K, int KM, int mode, double anneal_temp, double u, double mx) {
    mx = block_max(mx, s.red);
    double sum = 0.0;
    for (int k = threadIdx.x; k < KM; k += blockDim.x) sum += exp(s.sk[k] - mx);
    sum = block_sum(sum, s.red);
    const double lse = log(sum) + mx;
    if (mode == 0 && anneal_temp != 1.0) {
        const double inv_t = 1. / anneal_temp;
        double mq = inv_t * (mx - lse);      // max of the scaled vector (inv_t > 0)
        double s2 = 0.0;
        for (int k = threadIdx.x; k < KM; k += blockDim.x) {
            const double q = inv_t * (s.sk[k] - lse);
            s.sk[k] = q;
            s2 += exp(q - mq);
        }
        s2 = block_sum(s2, s.red);
        const double lse2 = log(s2) + mq;
        for (int k = threadIdx.x; k < KM; k += blockDim.x) s.sk[k] = exp(s.sk[k] - lse2);
    } else {
        for (int k = threadIdx.x; k < KM; k += blockDim.x) s.sk[k] = exp(s.sk[k] - lse);
    }
    __syncthreads();
    int k_sel;
    if (mode == 1) {
        // np.argmax(prob_z): first maximum
        double pm = -1.0;
        for (int k = threadIdx.x; k < KM; k += blockDim.x) pm = fmax(pm, s.sk[k]);
        pm = block_max(pm, s.red);
        int best = 0x7fffffff;
        for (int k = threadIdx.x; k < KM; k += blockDim.x)
            if (s.sk[k] == pm) { best = k; break; }
        // block min via the double reducer (indices are exactly representable)
        const double bm = -block_max(-(double)best, s.red);
        k_sel = (int)bm;
    } else {
        // utils.draw (utils.py:10-21): u -= p[i] sequentially, first i with u < 0, else last.
        // Parallel form: chunked inclusive prefix sums; exact-serial fallback whenever some
        // partial sum comes within 1e-9 of u (where rounding order could change the answer).
        const int nt = blockDim.x;
        const int per = (KM + nt - 1) / nt;
        const int lo = min(threadIdx.x * per, KM), hi = min(lo + per, KM);
        double loc = 0.0;
        for (int k = lo; k < hi; ++k) loc += s.sk[k];
        // exclusive scan of `loc` across threads: warp scan + warp totals
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        double inc = loc;
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        __syncthreads();
        if (lane == 31) s.red[w] = inc;
        __syncthreads();
        double wbase = 0.0;
        for (int i = 0; i < w; ++i) wbase += s.red[i];
        double run = wbase + inc - loc;
        int first = 0x7fffffff;
        double margin = CUDART_INF;
        for (int k = lo; k < hi; ++k) {
            run += s.sk[k];
            const double r = u - run;
            margin = fmin(margin, fabs(r));
            if (r < 0 && first == 0x7fffffff) first = k;
        }
        const double gmargin = -block_max(-margin, s.red);
        const double gfirst = -block_max(-(double)first, s.red);
        k_sel = (gfirst > 2.0e9) ? KM - 1 : (int)gfirst;
        if (gmargin < 1e-9) {
            if (threadIdx.x == 0) {
                double uu = u;
                int r = KM - 1;
                for (int k = 0; k < KM; ++k) { uu = uu - s.sk[k]; if (uu < 0) { r = k; break; } }
                s.red[38] = (double)r;
            }
            __syncthreads();
            k_sel = (int)s.red[38];
        }
    }
    if (k_sel > K) k_sel = K;      // several empty slots at the end (fbgmm.py:459-460)
    return k_sel;
}

// One slot of log_prob_vec_i (:64-69) or log(prob_vec_given_j) (:78-91), NumPy's operation order,
// every operation separately rounded.
__device__ __forceinline__ double lm_log_prob(const segb_bigram_lm &lm, int j_prev, int k, double sum_a) {
    const double uk = __dadd_rn((double)__ldcg(lm.unigram_counts + k), __ddiv_rn(lm.a, (double)lm.K));
    if (j_prev < 0) return __dsub_rn(log(uk), log(sum_a));
    const double pv = __ddiv_rn(uk, sum_a);
    const double bj = __dadd_rn((double)__ldcg(lm.bigram_counts + (size_t)j_prev * lm.K + k), __ddiv_rn(lm.b, (double)lm.K));
    const double t2 = __ddiv_rn(__dmul_rn(__dsub_rn(1., lm.intrp_lambda), bj), __dadd_rn((double)__ldcg(lm.unigram_counts + j_prev), lm.b));
    return log(__dadd_rn(__dmul_rn(lm.intrp_lambda, pv), t2));
}


}  // namespace segb
