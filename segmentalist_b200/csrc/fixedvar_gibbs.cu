// Persistent cooperative Gibbs sweep for the fixed-variance FBGMM segmenter.
//
// Replaces UnigramAcousticWordseg.gibbs_sample_i (unigram_acoustic_wordseg.py:252-360) called
// for a whole list of utterances, i.e. per utterance, strictly in order:
//   remove its tokens (:270-273 -> GaussianComponentsFixedVar.del_item / del_component,
//   gaussian_components_fixedvar.py:172-221), score every candidate segment
//   (get_vec_embed_log_probs :474-511 -> FBGMM.log_marg_i fbgmm.py:256-285), run the DP
//   (forward_backward :653-756 or forward_backward_viterbi :759-864), assign the new tokens left
//   to right (gibbs_sample_inside_loop_i / map_assign_i fbgmm.py:422-494 -> add_item :153-170).
//
// Collapsed Gibbs is sequential -- every assignment sees the statistics left by the previous
// one -- so the sweep is latency-bound, and the version that launches four kernels per
// utterance (fixedvar.cu) spends its time pulling the [D, K_max] tables through ONE SM for every
// token.  Here ONE cooperative launch runs the whole sweep: CTA b owns the components k with
// k % G == b (so the active ones are spread evenly) and keeps their statistics in shared memory; every step is a short local
// computation followed by a grid barrier:
//   remove   owner-local updates; the small replicated state (counts, K, n_total) is advanced
//            identically by every CTA, so no communication unless a component dies (then the
//            last component's statistics move between owners through the global tables)
//   score    every CTA scores all candidate segments against its own components -> partial
//            (max, sum-exp) per segment -> barrier -> CTA s % G combines segment s -> barrier
//   DP       every CTA runs the same warp-level DP on the same scores (no broadcast needed)
//   assign   per token: owners publish the log-probabilities of their slots -> barrier -> every
//            CTA reads all K_max values and takes the same decision (fv_decide); the owner of
//            the chosen slot updates its statistics
// Global tables (the host-visible model) are written through on every update; everything another
// CTA may have written during the launch is read with ld.global.cg (L1 is not coherent across SMs).
// All arithmetic is float64; predictive sums use NumPy's pairwise order with separately rounded operations.
#include <string.h>
#include "dp_warp.cuh"
#include "fixedvar_common.cuh"

namespace segb {

constexpr int GB_THREADS = 256;

struct GibbsParams {
    segb_fixedvar m;
    segb_corpus c;
    const int32_t *order;          // device
    int32_t n_order, fb_mode, assign_mode, per, M_cap, xb;   // xb = candidate rows staged per batch
    int32_t item_mode;             // 1: `order` lists ITEMS (FBGMM.gibbs_sample), 0: utterances
    double tpt, wip, anneal_temp, assign_temp;
    const double *uniforms;
    int64_t *u_counter;
    double *log_probs;
    int32_t *status;
    // work area (global)
    unsigned *bar;                 // [0] arrival counter (pre-loaded with bar_base), [16..] per-CTA trace tags
    unsigned bar_base;             // value the host wrote into the counter before the launch
    double *part_m, *part_t;       // [G][M_cap] partial (max, sum-exp) of every candidate per CTA
    double *seg_prior;             // [M_cap] log_prior of every candidate
    double *scores;                // [M_cap] banded scores of the current utterance
    double *v;                     // [2][K_max] slot log-probabilities of the current token (double-buffered)
    unsigned long long *prof;      // optional [16] per-phase clock totals of CTA 0 (development aid), or NULL
    segb_bigram_lm lm;             // bigram sweeps (has_lm): the LM tied to the components (bigram_lms.py)
    int32_t has_lm;
};

// Grid barrier (all CTAs are co-resident: cooperative launch) on ONE monotonic counter: CTA-wide
// bar.sync, then thread 0 arrives with a release-add and polls with acquire loads until the counter
// reaches this CTA's next target -- one L2 round trip to arrive, no reset / generation write by the
// last arriver (the sense-reversing version cost ~4 us per barrier on the per-token critical path).
// Every CTA takes part in every barrier, so each keeps its own target in shared memory
// (g_bar_target: base + generation * grid size, advanced by the grid size per barrier) and compares
// it with the counter MODULO 2^32: the counter may wrap any number of times during a launch (a
// whole-model sweep over 30M items arrives 4.4e9 times) because all CTAs are always within one
// grid size (< 2^31) of each other.  The host writes the base into the counter before every launch;
// a protocol bug traps instead of hanging the GPU.
__shared__ unsigned g_bar_target;
__device__ __forceinline__ void grid_barrier_init(unsigned base) {
    if (threadIdx.x == 0) g_bar_target = base;
}
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned n_blocks, unsigned tag = 0, unsigned h = 0) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned *trace = bar + 16;                          // [n_blocks] last barrier tag of every CTA (diagnostics)
        trace[blockIdx.x] = tag;
        trace[160 + blockIdx.x] = h;
        unsigned old, cur;
        const unsigned target = g_bar_target + n_blocks;
        g_bar_target = target;
        asm volatile("atom.add.release.gpu.u32 %0, [%1], 1;" : "=r"(old) : "l"(bar) : "memory");
        unsigned spins = 0;
        for (;;) {
            asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(cur) : "l"(bar) : "memory");
            if ((int)(cur - target) >= 0) break;
            if (++spins > (1u << 23)) {
                volatile unsigned *vt = trace;
                printf("segb gibbs: grid barrier timed out: block %d at tag %x; tags of blocks 0..7: %x %x %x %x %x %x %x %x; state %x %x %x %x %x %x %x %x\n",
                       blockIdx.x, tag, vt[0], vt[1], vt[2], vt[3], vt[4], vt[5], vt[6], vt[7],
                       vt[160], vt[161], vt[162], vt[163], vt[164], vt[165], vt[166], vt[167]);
                __trap();
            }
        }
        (void)old;
    }
    __syncthreads();
}

struct GibbsSmem {
    double *mu, *pp, *num, *pN;    // [per][D] statistics of the owned components
    double *lpp;                   // [per] log_prod_precision_pred
    double *pl;                    // [per] log(alpha/K_max + count)  (refreshed when counts change)
    double *cst, *hv, *iv;         // [per] diagonal model: Student's t constants for the slot's count (diag_consts)
    double *xs;                    // [xb][D] staged embeddings
    double *vt;                    // [xb][per] scores of the staged batch against the owned components
    double *red;                   // [40]
    double *sk;                    // [K_max]  (ScoreSmem view: xs | red | sk is NOT contiguous here)
    double *sc;                    // [M_cap] scores of the utterance
    double *al;                    // [N_cap + 1]
    double *tmp;                   // [D]
    double *bk;                    // [3 D + 1] cached statistics of one component (whole-model sweeps)
    int32_t *counts;               // [K_max] replicated
    int32_t *sid;                  // [M_cap] embedding id of every banded slot of the current utterance
    int32_t *tok;                  // [N_cap] ids of the tokens being removed
    int32_t *tk;                   // [N_cap] their components (snapshot taken before anything is modified)
    uint8_t *bo;                   // [N_cap]
};

__host__ __device__ inline size_t gibbs_smem_bytes(int D, int K_max, int per, int xb, int M_cap, int N_cap) {
    size_t d = (size_t)4 * per * D + 5 * per + (size_t)xb * D + (size_t)xb * per + 40 + K_max + M_cap + (N_cap + 1) + D + (3 * D + 1);
    return d * 8 + (size_t)K_max * 4 + (size_t)M_cap * 4 + (size_t)N_cap * 8 + ((N_cap + 15) / 16) * 16 + 64;
}

// per-dimension term of the predictive sum: fixed variance ((mu-x)^2 * prec_pred, :247-252) or
// diagonal (log(1 + (m-x)^2 * inv_var / v), gaussian_components_diag.py:255-258)
// (branch-free variants: with the model test inside, every term is its own basic block and the eight
// independent accumulators of the pairwise sum are no longer interleaved -- ~100 clocks per term)
__device__ __forceinline__ double pred_term_fixed(double mu, double pp, double x) {
    const double dl = __dsub_rn(mu, x);
    return __dmul_rn(__dmul_rn(dl, dl), pp);
}
__device__ __forceinline__ double pred_term_diag(double mu, double pp, double x, double iv) {
    const double dl = __dsub_rn(mu, x);
    return log(__dadd_rn(1., __dmul_rn(__dmul_rn(__dmul_rn(dl, dl), pp), iv)));
}
__device__ __forceinline__ double pred_term(bool diag, double mu, double pp, double x, double iv) {
    const double dl = __dsub_rn(mu, x);
    const double q = __dmul_rn(__dmul_rn(dl, dl), pp);
    return diag ? log(__dadd_rn(1., __dmul_rn(q, iv))) : q;
}
// predictive log-density from the summed terms
__device__ __forceinline__ double pred_value(bool diag, double acc, double c0, double lpp, double cst, double hv) {
    return diag ? ((cst - 0.5 * lpp) - hv * acc) : ((c0 + 0.5 * lpp) - 0.5 * acc);
}

// gibbs_sample_inside_loop_i's draw (fbgmm.py:441-463, no annealing) on the per-token critical path of the
// cooperative sweep.  Same decision as fv_decide, restructured for latency: every thread owns a contiguous
// chunk of the K_max slot values (independent loads), ONE exponential per slot, the draw is taken in the
// unnormalised domain (first k with u * sum - prefix_k < 0, prefix over e_k = exp(v_k - max)), and the three
// block-wide exchanges (max; chunk sums = total + scan bases; first hit + margin) cost one barrier each.
// Whenever a prefix comes within 1e-9 (relative to the total) of the threshold -- where summation order
// could change the answer -- thread 0 repeats the reference's serial subtraction over exp(v_k - lse).
// red: >= 32 doubles of shared scratch.  Block-uniform, identical on every CTA.
constexpr int DECIDE_PER = 8;
__device__ int decide_fast(const double *vbuf, int K, int KM, double u, double *red, double inv_t) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = GB_THREADS >> 5;
    const int per = (KM + GB_THREADS - 1) / GB_THREADS;
    const int lo = min(tid * per, KM), hi = min(lo + per, KM);
    double v[DECIDE_PER];
#pragma unroll
    for (int i = 0; i < DECIDE_PER; ++i) v[i] = (lo + i < hi) ? __ldcg(vbuf + lo + i) : neg_inf();
    double mx = v[0];
#pragma unroll
    for (int i = 1; i < DECIDE_PER; ++i) mx = fmax(mx, v[i]);
    mx = warp_max(mx);
    if (lane == 0) red[w] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < nw; ++i) mx = fmax(mx, red[i]);
    double loc = 0.0;
#pragma unroll
    for (int i = 0; i < DECIDE_PER; ++i) { v[i] = (lo + i < hi) ? exp(inv_t * (v[i] - mx)) : 0.0; loc += v[i]; }
    double inc = loc;                       // inclusive scan of the chunk sums inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) red[8 + w] = inc;
    __syncthreads();
    double base = 0.0, sum = 0.0;
    for (int i = 0; i < nw; ++i) { const double t = red[8 + i]; if (i < w) base += t; sum += t; }
    const double target = u * sum;
    double run = base + inc - loc, margin = CUDART_INF;
    int first = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < DECIDE_PER; ++i) {
        if (lo + i < hi) {
            run += v[i];
            const double r = target - run;
            margin = fmin(margin, fabs(r));
            if (r < 0 && first == 0x7fffffff) first = lo + i;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        margin = fmin(margin, __shfl_xor_sync(FULL, margin, o));
        first = min(first, __shfl_xor_sync(FULL, first, o));
    }
    if (lane == 0) { red[16 + w] = margin; red[24 + w] = (double)first; }
    __syncthreads();
    margin = red[16];
    double gfirst = red[24];
    for (int i = 1; i < nw; ++i) { margin = fmin(margin, red[16 + i]); gfirst = fmin(gfirst, red[24 + i]); }
    int k_sel = (gfirst > 2.0e9) ? KM - 1 : (int)gfirst;
    if (margin < 1e-9 * sum && inv_t != 1.0) { __syncthreads(); return -1; }   // annealed: the caller repeats the draw with fv_decide
    if (margin < 1e-9 * sum) {
        if (tid == 0) {                     // utils.draw (utils.py:10-21) verbatim
            const double lse = log(sum) + mx;
            double uu = u;
            int r = KM - 1;
            for (int k = 0; k < KM; ++k) { uu = uu - exp(__ldcg(vbuf + k) - lse); if (uu < 0) { r = k; break; } }
            red[32] = (double)r;
        }
        __syncthreads();
        k_sel = (int)red[32];
    }
    __syncthreads();                        // red is reused by the caller
    if (k_sel > K) k_sel = K;               // several empty slots at the end (fbgmm.py:459-460)
    return k_sel;
}

// map_assign_i's choice (fbgmm.py:465-494: first maximum of the exp-normalised vector) on the same low-latency
// plan: the first maximum of the log-probabilities themselves, unless a second slot lies within 1e-9 of it
// (where rounding of the exponentials could merge or reorder them) -- then -1, and the caller runs fv_decide.
__device__ int decide_map_fast(const double *vbuf, int K, int KM, double *red) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = GB_THREADS >> 5;
    const int per = (KM + GB_THREADS - 1) / GB_THREADS;
    const int lo = min(tid * per, KM), hi = min(lo + per, KM);
    double v[DECIDE_PER];
#pragma unroll
    for (int i = 0; i < DECIDE_PER; ++i) v[i] = (lo + i < hi) ? __ldcg(vbuf + lo + i) : neg_inf();
    double mx = v[0];
#pragma unroll
    for (int i = 1; i < DECIDE_PER; ++i) mx = fmax(mx, v[i]);
    mx = warp_max(mx);
    if (lane == 0) red[w] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < nw; ++i) mx = fmax(mx, red[i]);
    int near = 0, first = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < DECIDE_PER; ++i) {
        if (lo + i < hi) {
            near += (v[i] >= mx - 1e-9) ? 1 : 0;
            if (v[i] == mx && first == 0x7fffffff) first = lo + i;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        near += __shfl_xor_sync(FULL, near, o);
        first = min(first, __shfl_xor_sync(FULL, first, o));
    }
    if (lane == 0) { red[8 + w] = (double)near; red[16 + w] = (double)first; }
    __syncthreads();
    double n_near = 0.0, gfirst = red[16];
    for (int i = 0; i < nw; ++i) { n_near += red[8 + i]; gfirst = fmin(gfirst, red[16 + i]); }
    __syncthreads();                        // red is reused by the caller
    if (!(mx == mx) || n_near != 1.0 || gfirst > 2.0e9) return -1;
    int k_sel = (int)gfirst;
    if (k_sel > K) k_sel = K;
    return k_sel;
}

// decide_fast for models with more than GB_THREADS * DECIDE_PER slots (K_max = 5000): the slot values pass
// through shared memory -- coalesced loads and the exponentials in slot-strided order, then every thread scans
// its contiguous chunk.  Same decision rule, same exact-serial fallback.  sk: K_max doubles of shared scratch.
__device__ int decide_fast_smem(const double *vbuf, int K, int KM, double u, double *red, double *sk, double inv_t) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = GB_THREADS >> 5;
    double mx = neg_inf();
    for (int k0 = tid; k0 < KM; k0 += 4 * GB_THREADS) {
        double t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int k = k0 + q * GB_THREADS; t[q] = (k < KM) ? __ldcg(vbuf + k) : neg_inf(); }
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int k = k0 + q * GB_THREADS; if (k < KM) { sk[k] = t[q]; mx = fmax(mx, t[q]); } }
    }
    mx = warp_max(mx);
    if (lane == 0) red[w] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < nw; ++i) mx = fmax(mx, red[i]);
    for (int k = tid; k < KM; k += GB_THREADS) sk[k] = exp(inv_t * (sk[k] - mx));
    __syncthreads();
    const int per = (KM + GB_THREADS - 1) / GB_THREADS;
    const int lo = min(tid * per, KM), hi = min(lo + per, KM);
    double loc = 0.0;
    for (int k = lo; k < hi; ++k) loc += sk[k];
    double inc = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) red[8 + w] = inc;
    __syncthreads();
    double base = 0.0, sum = 0.0;
    for (int i = 0; i < nw; ++i) { const double t = red[8 + i]; if (i < w) base += t; sum += t; }
    const double target = u * sum;
    double run = base + inc - loc, margin = CUDART_INF;
    int first = 0x7fffffff;
    for (int k = lo; k < hi; ++k) {
        run += sk[k];
        const double r = target - run;
        margin = fmin(margin, fabs(r));
        if (r < 0 && first == 0x7fffffff) first = k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        margin = fmin(margin, __shfl_xor_sync(FULL, margin, o));
        first = min(first, __shfl_xor_sync(FULL, first, o));
    }
    if (lane == 0) { red[16 + w] = margin; red[24 + w] = (double)first; }
    __syncthreads();
    margin = red[16];
    double gfirst = red[24];
    for (int i = 1; i < nw; ++i) { margin = fmin(margin, red[16 + i]); gfirst = fmin(gfirst, red[24 + i]); }
    int k_sel = (gfirst > 2.0e9) ? KM - 1 : (int)gfirst;
    if (margin < 1e-9 * sum && inv_t != 1.0) { __syncthreads(); return -1; }   // annealed: the caller repeats the draw with fv_decide
    if (margin < 1e-9 * sum) {
        if (tid == 0) {                     // utils.draw (utils.py:10-21) verbatim
            const double lse = log(sum) + mx;
            double uu = u;
            int r = KM - 1;
            for (int k = 0; k < KM; ++k) { uu = uu - exp(__ldcg(vbuf + k) - lse); if (uu < 0) { r = k; break; } }
            red[32] = (double)r;
        }
        __syncthreads();
        k_sel = (int)red[32];
    }
    __syncthreads();
    if (k_sel > K) k_sel = K;
    return k_sel;
}

// ITEM: whole-model sweep over items (FBGMM.gibbs_sample) instead of utterances.  HAS_LM: bigram sweeps.
// Separate instantiations keep one call site per step lambda, so each is inlined into its kernel (with both
// modes in one body the compiler left the lambdas out of line and their captures went to local memory).
template <bool ITEM, bool HAS_LM>
__global__ void __launch_bounds__(GB_THREADS, 1) fv_gibbs_kernel(GibbsParams p) {
    extern __shared__ __align__(16) unsigned char gsm[];
    const segb_fixedvar &m = p.m;
    const segb_corpus &c = p.c;
    const int D = m.D, KM = m.K_max, per = p.per, xb = p.xb, S = c.S;
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // component k lives on CTA k % G as local slot k / G: the ACTIVE components 0..K-1 are spread
    // evenly over the grid whatever K is (a contiguous split would leave most CTAs idle when
    // K << K_max); every CTA's active slots are its first n_act(K) local slots
    const int n_own = (b < KM) ? (KM - b + G - 1) / G : 0;
    auto own_of = [&](int k) { return (k % G) == b; };
    auto kl_of = [&](int k) { return k / G; };
    auto k_of = [&](int kl) { return b + kl * G; };
    auto n_act_of = [&](int K_) { return (K_ > b) ? min(n_own, (K_ - b + G - 1) / G) : 0; };

    grid_barrier_init(p.bar_base);
    GibbsSmem s;
    {
        double *q = reinterpret_cast<double *>(gsm);
        s.mu = q; q += (size_t)per * D;
        s.pp = q; q += (size_t)per * D;
        s.num = q; q += (size_t)per * D;
        s.pN = q; q += (size_t)per * D;
        s.lpp = q; q += per;
        s.pl = q; q += per;
        s.cst = q; q += per; s.hv = q; q += per; s.iv = q; q += per;
        s.xs = q; q += (size_t)xb * D;
        s.vt = q; q += (size_t)xb * per;
        s.red = q; q += 40;
        s.sk = q; q += KM;
        s.sc = q; q += p.M_cap;
        s.al = q; q += c.N_max + 1;
        s.tmp = q; q += D;
        s.bk = q; q += 3 * D + 1;
        s.counts = reinterpret_cast<int32_t *>(q);
        s.sid = s.counts + KM;
        s.tok = s.sid + p.M_cap;
        s.tk = s.tok + c.N_max;
        s.bo = reinterpret_cast<uint8_t *>(s.tk + c.N_max);
    }
    ScoreSmem ds(s.xs, D);          // only .red / .sk are used by fv_decide
    ds.red = s.red; ds.sk = s.sk;
    long long t_last = clock64();
    auto tick = [&](int ph) {
        if (p.prof && blockIdx.x == 0 && threadIdx.x == 0) {
            const long long now = clock64();
            p.prof[ph] += (unsigned long long)(now - t_last);
            t_last = now;
        }
    };

    // ---- load the owned statistics and the replicated state
    for (int i = tid; i < n_own * D; i += GB_THREADS) {
        const int kl = i / D, d = i % D;
        const size_t o = (size_t)d * KM + k_of(kl);
        s.mu[kl * D + d] = m.mu_NT[o]; s.pp[kl * D + d] = m.prec_predT[o];
        s.num[kl * D + d] = m.mu_N_numT[o]; s.pN[kl * D + d] = m.prec_NT[o];
    }
    for (int kl = tid; kl < n_own; kl += GB_THREADS) s.lpp[kl] = m.log_prod_prec_pred[k_of(kl)];
    for (int k = tid; k < KM; k += GB_THREADS) s.counts[k] = m.counts[k];
    int K = *m.K;
    long long n_total = *m.n_total;
    long long u_pos = p.u_counter ? *p.u_counter : 0;
    const double c0 = fv_norm_const(D);
    const double log_empty = log(m.alpha / KM + 0.);
    const bool diag = (m.model == SEGB_MODEL_DIAG);
    // diagonal model: prior predictive = Student's t with var_0 = (k_0+1)/(k_0 v_0) S_0 (:216-223)
    double d_cst0 = 0., d_hv0 = 0., d_iv0 = 0., d_f0 = 0., d_lpv0 = 0.;
    if (diag) {
        diag_consts(m, 0, d_cst0, d_hv0, d_iv0);
        d_f0 = (m.k_0 + 1.) / (m.k_0 * m.v_0);
        d_lpv0 = pairwise_sum_max256<double>([&](int d) { return log(d_f0 * m.precision_0[d]); }, D);
    }
    // log_prior(x) for the embedding staged at xrow, by one warp (same bits on every CTA)
    auto warp_log_prior = [&](const double *xrow) -> double {
        double acc = 0.0;
        for (int d = lane; d < D; d += 32) {
            const double dl = xrow[d] - m.mu_0[d];
            acc += diag ? log(1. + d_iv0 * (dl * dl) * (1. / (d_f0 * m.precision_0[d]))) : dl * dl * m.precision_0[d];
        }
        acc = warp_sum(acc);
        return diag ? (d_cst0 - 0.5 * d_lpv0 - d_hv0 * acc) : (c0 + 0.5 * m.sum_log_precision_0 - 0.5 * acc);
    };
    // count-dependent constants of owned slot kl
    auto slot_consts = [&](int kl) {
        const int n = s.counts[k_of(kl)];
        s.pl[kl] = log(m.alpha / KM + (double)n);
        if (diag) diag_consts(m, n, s.cst[kl], s.hv[kl], s.iv[kl]);
    };
    unsigned tok_parity = 0;
    __syncthreads();

    // refresh precision_pred, mu_N, log_prod_precision_pred of owned slot kl and write the
    // component through to the global tables (:317-325)
    auto refresh_and_publish = [&](int kl) {
        const int k = k_of(kl);
        const double k_N = m.k_0 + (double)s.counts[k], v_N = (double)(m.v_0 + s.counts[k]);
        const double f = diag ? __ddiv_rn(k_N + 1., __dmul_rn(k_N, v_N)) : 0.;
        for (int d = tid; d < D; d += GB_THREADS) {
            const double pNv = s.pN[kl * D + d];
            double ppv, muv;
            if (diag) {                                  // gaussian_components_diag.py:332-345
                muv = __ddiv_rn(s.num[kl * D + d], k_N);
                const double var = __dmul_rn(f, __dsub_rn(pNv, __dmul_rn(k_N, __dmul_rn(muv, muv))));
                ppv = __ddiv_rn(1., var);
                s.tmp[d] = log(var);
            } else {
                const double pr = m.precision[d];
                ppv = __ddiv_rn(__dmul_rn(pNv, pr), __dadd_rn(pNv, pr));
                muv = __ddiv_rn(s.num[kl * D + d], pNv);
                s.tmp[d] = log(ppv);
            }
            s.pp[kl * D + d] = ppv;
            s.mu[kl * D + d] = muv;
            const size_t o = (size_t)d * KM + k;
            m.mu_N_numT[o] = s.num[kl * D + d]; m.prec_NT[o] = pNv; m.prec_predT[o] = ppv; m.mu_NT[o] = muv;
        }
        __syncthreads();
        if (tid < 16) {
            const double l = pairwise_sum_lanes16<double>([&](int i) { return s.tmp[i]; }, D, 0xffffu, tid);   // D <= 256 (launcher)
            if (tid == 0) {
                s.lpp[kl] = l;
                m.log_prod_prec_pred[k] = l;
                m.counts[k] = s.counts[k];
            }
        }
        __syncthreads();
    };
    auto zero_slot = [&](int kl) {
        const int k = k_of(kl);
        for (int d = tid; d < D; d += GB_THREADS) {
            s.mu[kl * D + d] = 0.; s.pp[kl * D + d] = 0.; s.num[kl * D + d] = 0.; s.pN[kl * D + d] = 0.;
            const size_t o = (size_t)d * KM + k;
            m.mu_N_numT[o] = 0.; m.prec_NT[o] = 0.; m.prec_predT[o] = 0.; m.mu_NT[o] = 0.;
        }
        if (tid == 0) { s.lpp[kl] = 0.; m.log_prod_prec_pred[k] = 0.; m.counts[k] = 0; }
        __syncthreads();
    };

    // del_item (:172-188) of item `id`, currently in component k.  The replicated state (counts, K,
    // n_total) advances on every CTA; only the owner touches statistics.  Entries j_next.. of the
    // pending removal list s.tk[0..n_list) are relabelled if a component moves.
    auto remove_one = [&](int id, int k, unsigned tag, int j_next, int n_list, bool scan_all) {
        __syncthreads();
        const int cnt = s.counts[k] - 1;
        __syncthreads();
        if (tid == 0) s.counts[k] = cnt;
        n_total -= 1;
        const bool own = own_of(k);
        if (cnt > 0) {
            if (own) {
                const int kl = kl_of(k);
                for (int d = tid; d < D; d += GB_THREADS) {
                    if (diag) {                              // gaussian_components_diag.py:190-193
                        s.num[kl * D + d] = __dsub_rn(s.num[kl * D + d], fv_x(m, id, d));
                        s.pN[kl * D + d] = __dsub_rn(s.pN[kl * D + d], fv_xsq(m, id, d));
                    } else {
                        const double pr = m.precision[d];
                        s.num[kl * D + d] = __dsub_rn(s.num[kl * D + d], __dmul_rn(pr, fv_x(m, id, d)));
                        s.pN[kl * D + d] = __dsub_rn(s.pN[kl * D + d], pr);
                    }
                }
                __syncthreads();
                refresh_and_publish(kl);
            }
        } else {
            // del_component (:190-221): the last component moves into slot k
            const int last = K - 1;
            grid_barrier(p.bar, G, tag | 0x10);   // every owner's write-through is visible
            if (HAS_LM && b == 0) {
                // the tied LM counts follow the component (gaussian_components_fixedvar.py:205-208, :218-221):
                // unigram count, row `last` -> row k, then column `last` -> column k; the old slot is cleared
                const int KL = p.lm.K;
                int32_t *bi = p.lm.bigram_counts;
                if (k != last) {
                    if (tid == 0) p.lm.unigram_counts[k] = p.lm.unigram_counts[last];
                    for (int i = tid; i < KL; i += GB_THREADS) bi[(size_t)k * KL + i] = bi[(size_t)last * KL + i];
                    __syncthreads();
                    for (int j = tid; j < KL; j += GB_THREADS) bi[(size_t)j * KL + k] = bi[(size_t)j * KL + last];
                    __syncthreads();
                }
                if (tid == 0) p.lm.unigram_counts[last] = 0;
                for (int i = tid; i < KL; i += GB_THREADS) { bi[(size_t)last * KL + i] = 0; bi[(size_t)i * KL + last] = 0; }
                __syncthreads();
            }
            if (k != last) {
                if (own) {
                    const int kl = kl_of(k);
                    for (int d = tid; d < D; d += GB_THREADS) {
                        const size_t a = (size_t)d * KM + k, o = (size_t)d * KM + last;
                        const double nu = __ldcg(m.mu_N_numT + o), pNv = __ldcg(m.prec_NT + o), ppv = __ldcg(m.prec_predT + o),
                                     muv = __ldcg(m.mu_NT + o);
                        s.num[kl * D + d] = nu; s.pN[kl * D + d] = pNv; s.pp[kl * D + d] = ppv; s.mu[kl * D + d] = muv;
                        m.mu_N_numT[a] = nu; m.prec_NT[a] = pNv; m.prec_predT[a] = ppv; m.mu_NT[a] = muv;
                        m.mu_N_numT[o] = 0.; m.prec_NT[o] = 0.; m.prec_predT[o] = 0.; m.mu_NT[o] = 0.;
                    }
                    if (tid == 0) {
                        const double l = __ldcg(m.log_prod_prec_pred + last);
                        s.lpp[kl] = l;
                        m.log_prod_prec_pred[k] = l; m.log_prod_prec_pred[last] = 0.;
                        m.counts[k] = s.counts[last]; m.counts[last] = 0;
                    }
                }
                if (own_of(last)) {           // the old home forgets it (global zeroed by the new owner)
                    const int kl = kl_of(last);
                    for (int d = tid; d < D; d += GB_THREADS) {
                        s.mu[kl * D + d] = 0.; s.pp[kl * D + d] = 0.; s.num[kl * D + d] = 0.; s.pN[kl * D + d] = 0.;
                    }
                    if (tid == 0) s.lpp[kl] = 0.;
                }
                // relabel the moved component's members, split over the grid: the live tokens of the
                // corpus (segmenter sweeps) or every item (whole-model sweeps, as the reference does, :202)
                if (scan_all) {
                    for (int64_t i = (int64_t)b * GB_THREADS + tid; i < m.n_emb; i += (int64_t)G * GB_THREADS)
                        if (__ldcg(m.assignments + i) == last) m.assignments[i] = k;
                } else {
                    for (int64_t i = (int64_t)b * GB_THREADS + tid; i < c.n_pos; i += (int64_t)G * GB_THREADS) {
                        const int t_id = __ldcg(c.tok_id + i);
                        if (t_id >= 0 && __ldcg(m.assignments + t_id) == last) m.assignments[t_id] = k;
                    }
                }
                for (int jj = j_next + tid; jj < n_list; jj += GB_THREADS) if (s.tk[jj] == last) s.tk[jj] = k;
                __syncthreads();
                if (tid == 0) { s.counts[k] = s.counts[last]; s.counts[last] = 0; }
            } else if (own) {
                zero_slot(kl_of(k));
            }
            K = last;
            grid_barrier(p.bar, G, tag | 0x20);   // relabelled assignments are visible
        }
        __syncthreads();
    };

    // One assignment step (gibbs_sample_inside_loop_i / map_assign_i, fbgmm.py:422-494) for item `id`
    // whose embedding is staged in s.xs[0..D): owners publish their slots' log-probabilities, grid
    // barrier, every CTA takes the same decision, the owner of the chosen slot updates.
    // x_prior = log_prior(x); k_restore = slot whose cached statistics (s.bk) are put back when it is
    // chosen again, or -1.
    const double *x_tok = s.xs;      // embedding of the token being assigned (a row of s.xs)
    int lm_prev = -1;                // bigram sweeps: label of the utterance's previous token, LM normaliser
    double lm_sum_a = 0.0;
    auto assign_one = [&](int id, double x_prior, unsigned tag, int k_restore) -> int {
    double *vbuf = p.v + (size_t)tok_parity * KM;
    tok_parity ^= 1;
    const int na = n_act_of(K);
    {
        // one half-warp per owned slot (the predictive sum is 130 dependent-latency terms)
        const int hw = tid >> 4, jl = tid & 15;
        const unsigned hmask = 0xffffu << (lane & 16);
        for (int kl = hw; kl < n_own; kl += GB_THREADS / 16) {
            double val;
            if (kl < na) {
                const double *mu = s.mu + kl * D, *pp = s.pp + kl * D;
                const double iv = diag ? s.iv[kl] : 0.;
                auto term_f = [&](int d) { return pred_term_fixed(mu[d], pp[d], x_tok[d]); };
                auto term_d = [&](int d) { return pred_term_diag(mu[d], pp[d], x_tok[d], iv); };
                const double acc = diag ? pairwise_sum_lanes16<double>(term_d, D, hmask, jl)     // D <= 256 (launcher)
                                        : pairwise_sum_lanes16<double>(term_f, D, hmask, jl);
                // prior term: lms*log(alpha/K_max + n_k) (fbgmm.py:436), or under a bigram LM the row of the
                // previous label, lms*log P(k | j_prev) (bigram_acoustic_wordseg.py:347-355)
                const double prior = HAS_LM ? __dmul_rn(lm_log_prob(p.lm, lm_prev, k_of(kl), lm_sum_a), m.lms)
                                              : ((p.assign_mode == 0) ? m.lms * s.pl[kl] : s.pl[kl]);
                val = prior + pred_value(diag, acc, c0, s.lpp[kl], diag ? s.cst[kl] : 0., diag ? s.hv[kl] : 0.);
            } else {
                val = HAS_LM ? __dadd_rn(__dmul_rn(lm_log_prob(p.lm, lm_prev, k_of(kl), lm_sum_a), m.lms), x_prior)
                               : ((p.assign_mode == 0) ? m.lms : 1.0) * log_empty + x_prior;
            }
            if (jl == 0) vbuf[k_of(kl)] = val;
        }
    }
    tick(4);
    grid_barrier(p.bar, G, tag | 0x50, (unsigned)(K | ((unsigned)n_total << 8) | ((unsigned)u_pos << 20)));
    tick(5);
    const double uu = (p.assign_mode == 0) ? p.uniforms[u_pos] : 0.0;
    if (p.assign_mode == 0) u_pos += 1;
    int k_sel;
    // sampling (also annealed: p_k ~ exp((v_k - max) / T), fbgmm.py:446-449) and MAP assignment take the
    // low-latency paths; near-ties and the rare annealed draws within 1e-9 of a CDF step use fv_decide
    k_sel = -1;
    if (p.assign_mode == 0)
        k_sel = (KM <= GB_THREADS * DECIDE_PER) ? decide_fast(vbuf, K, KM, uu, s.red, 1. / p.assign_temp)
                                                : decide_fast_smem(vbuf, K, KM, uu, s.red, s.sk, 1. / p.assign_temp);
    else if (KM <= GB_THREADS * DECIDE_PER)
        k_sel = decide_map_fast(vbuf, K, KM, s.red);
    if (k_sel < 0) {
        double mx = neg_inf();
        for (int k = tid; k < KM; k += GB_THREADS) {
            const double val = __ldcg(vbuf + k);
            s.sk[k] = val;
            mx = fmax(mx, val);
        }
        k_sel = fv_decide(ds, K, KM, p.assign_mode, p.assign_temp, uu, mx);
    }
    tick(6);
    // add_item (:153-170) -- or, in whole-model sweeps, put the cached statistics back when the
    // item returns to its old component and no component died in between (fbgmm.py:397-400)
    const bool fresh = (k_sel == K);
    const bool restore = (k_sel == k_restore);
    __syncthreads();
    if (tid == 0) s.counts[k_sel] += 1;
    if (fresh) K += 1;
    n_total += 1;
    if (own_of(k_sel)) {
        const int kl = kl_of(k_sel);
        for (int d = tid; d < D; d += GB_THREADS) {
            if (restore) { s.num[kl * D + d] = s.bk[d]; s.pN[kl * D + d] = s.bk[D + d]; continue; }
            double nu = s.num[kl * D + d], pNv = s.pN[kl * D + d];
            if (diag) {                                  // gaussian_components_diag.py:162-177
                if (fresh) {
                    nu = __dmul_rn(m.k_0, m.mu_0[d]);
                    pNv = __dadd_rn(m.precision_0[d], __dmul_rn(m.k_0, __dmul_rn(m.mu_0[d], m.mu_0[d])));
                }
                s.num[kl * D + d] = __dadd_rn(nu, x_tok[d]);
                s.pN[kl * D + d] = __dadd_rn(pNv, fv_xsq(m, id, d));
                continue;
            }
            if (fresh) { nu = __dmul_rn(m.precision_0[d], m.mu_0[d]); pNv = m.precision_0[d]; }
            s.num[kl * D + d] = __dadd_rn(nu, __dmul_rn(m.precision[d], x_tok[d]));
            s.pN[kl * D + d] = __dadd_rn(pNv, m.precision[d]);
        }
        __syncthreads();
        if (tid == 0) {
            m.assignments[id] = k_sel;
            slot_consts(kl);
        }
        refresh_and_publish(kl);
        if (restore) {       // cached precision_pred / log_prod_precision_pred go back bit for bit
            for (int d = tid; d < D; d += GB_THREADS) {
                s.pp[kl * D + d] = s.bk[2 * D + d];
                m.prec_predT[(size_t)d * KM + k_sel] = s.bk[2 * D + d];
            }
            if (tid == 0) { s.lpp[kl] = s.bk[3 * D]; m.log_prod_prec_pred[k_sel] = s.bk[3 * D]; }
        }
    }
    __syncthreads();
    tick(7);
    return k_sel;
    };

    // ================= whole-model sweep: FBGMM.gibbs_sample over a list of items (fbgmm.py:357-400)
    if (ITEM) {
        for (int it = 0; it < p.n_order; ++it) {
            const int id = p.order[it];
            const int k_old = __ldcg(m.assignments + id);            // uniform over the grid
            const int K_old = K;
            __syncthreads();
            if (k_old >= 0) {
                if (own_of(k_old)) {                    // cache_component_stats (:128-141)
                    const int kl = kl_of(k_old);
                    for (int d = tid; d < D; d += GB_THREADS) {
                        s.bk[d] = s.num[kl * D + d]; s.bk[D + d] = s.pN[kl * D + d]; s.bk[2 * D + d] = s.pp[kl * D + d];
                    }
                    if (tid == 0) s.bk[3 * D] = s.lpp[kl];
                }
                __syncthreads();
                remove_one(id, k_old, (unsigned)it << 8, 0, 0, true);
            }
            for (int d = tid; d < D; d += GB_THREADS) s.xs[d] = fv_x(m, id, d);
            __syncthreads();
            if (warp == 0) {                                            // log_prior(x), same bits on every CTA
                const double lp = warp_log_prior(s.xs);
                if (lane == 0) s.red[39] = lp;
            }
            for (int kl = tid; kl < n_own; kl += GB_THREADS) slot_consts(kl);
            __syncthreads();
            const double x_prior = s.red[39];
            assign_one(id, x_prior, (unsigned)it << 8, (k_old >= 0 && K == K_old) ? k_old : -1);
        }
        if (b == 0 && tid == 0) {
            *m.K = K;
            *m.n_total = n_total;
            if (p.u_counter) *p.u_counter = u_pos;
        }
        return;
    }

    for (int it = 0; it < p.n_order; ++it) {
        const int u = p.order[it];
        const int64_t off = c.pos_off[u];
        const int N = (int)(c.pos_off[u + 1] - off);
        const int n_slots = N * S;

        // the same utterance twice in a row: its new tokens must be visible before they are read
        if (it > 0 && p.order[it - 1] == u) grid_barrier(p.bar, G, (it << 8) | 1);

        tick(8);
        // ================= remove the utterance's current tokens (:270-273)
        // snapshot of (token, component) before anything is modified: every CTA must see the same list
        for (int j = tid; j < N; j += GB_THREADS) {
            const int id = __ldcg(c.tok_id + off + j);
            s.tok[j] = id;
            s.tk[j] = (id >= 0) ? __ldcg(m.assignments + id) : -1;
        }
        for (int j = tid; j < n_slots; j += GB_THREADS) s.sid[j] = c.seg_id[off * S + j];
        // one CTA pulls the NEXT utterance's embeddings and slot table into L2 while this one is sampled
        if (it + 1 < p.n_order && b == (it % G)) {
            const int un = p.order[it + 1];
            const int64_t offn = c.pos_off[un];
            const int n_sn = (int)(c.pos_off[un + 1] - offn) * S;
            const size_t row_bytes = (size_t)D * (m.x_is_f64 ? 8 : 4);
            for (int j = tid; j < n_sn; j += GB_THREADS) {
                const int id = c.seg_id[offn * S + j];
                if (id < 0) continue;
                const char *row = (const char *)m.X + (size_t)id * row_bytes;
                for (size_t o = 0; o < row_bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + o));
            }
        }
        __syncthreads();
        if (HAS_LM && b == 0 && tid == 0) {
            // remove_counts_from_utterance (bigram_lms.py:107-113) over the old transcript, before any
            // component can move (bigram_acoustic_wordseg.py:410-417); CTA 0 owns the LM tables
            int jp = -1;
            for (int j = 0; j < N; ++j) {
                if (s.tok[j] < 0 || s.tk[j] < 0) continue;
                const int i = s.tk[j];
                p.lm.unigram_counts[i] -= 1;
                if (jp >= 0) p.lm.bigram_counts[(size_t)jp * p.lm.K + i] -= 1;
                jp = i;
            }
        }
        __syncthreads();
        for (int j = 0; j < N; ++j) {
            const int id = s.tok[j];
            const int k = s.tk[j];
            if (id < 0 || k < 0) continue;
            remove_one(id, k, (it << 8) | (j << 16), j + 1, N, false);
        }

        tick(0);
        // ================= score every candidate segment (:474-511, fbgmm.py:256-285)
        const double log_norm = log((double)n_total + m.alpha);
        for (int kl = tid; kl < n_own; kl += GB_THREADS) slot_consts(kl);
        __syncthreads();
        const int n_act = n_act_of(K);           // owned ACTIVE components
        // Embeddings of a batch are staged as float64, warp per row, lanes over dimensions.  The loads of
        // batch i+1 are issued right after batch i has been stored, so their latency hides behind the pair
        // evaluation of batch i (registers as the prefetch buffer).
        constexpr int NW = GB_THREADS / 32, RPW = 4, DCH = 5;     // rows per warp / 32-wide chunks of D kept in registers
        const bool reg_stage = (D <= 32 * DCH) && !m.x_is_f64 && (xb <= NW * RPW);
        float v[RPW][DCH];
        auto load_batch = [&](int s0, int nb) {
            const float *Xf = (const float *)m.X;
#pragma unroll
            for (int q = 0; q < RPW; ++q) {
                const int bb = warp + q * NW;
                const int id = (bb < nb) ? s.sid[s0 + bb] : -1;
                const float *row = Xf + (size_t)(id < 0 ? 0 : id) * D;
#pragma unroll
                for (int cch = 0; cch < DCH; ++cch) {
                    const int d = lane + 32 * cch;
                    v[q][cch] = (id >= 0 && d < D) ? row[d] : 0.f;
                }
            }
        };
        if (reg_stage) load_batch(0, min(xb, n_slots));
        for (int s0 = 0; s0 < n_slots; s0 += xb) {
            const int nb = min(xb, n_slots - s0);
            if (reg_stage) {
#pragma unroll
                for (int q = 0; q < RPW; ++q) {
                    const int bb = warp + q * NW;
                    if (bb < nb) {
#pragma unroll
                        for (int cch = 0; cch < DCH; ++cch) {
                            const int d = lane + 32 * cch;
                            if (d < D) s.xs[bb * D + d] = (double)v[q][cch];
                        }
                    }
                }
            } else {
                for (int bb = warp; bb < nb; bb += NW) {
                    const int id = s.sid[s0 + bb];
                    for (int d = lane; d < D; d += 32) s.xs[bb * D + d] = (id >= 0) ? fv_x(m, id, d) : 0.0;
                }
            }
            __syncthreads();
            tick(11);
            if (reg_stage && s0 + xb < n_slots) load_batch(s0 + xb, min(xb, n_slots - s0 - xb));
            if (diag) {
                // Student's t terms cost a float64 log each: a half-warp per (segment, component) pair (lane =
                // NumPy accumulator) keeps all 256 threads busy where thread-per-pair left most of them idle
                const int jl = tid & 15;
                const unsigned hmask = 0xffffu << (lane & 16);
                for (int i = tid >> 4; i < nb * n_act; i += GB_THREADS / 16) {
                    const int bb = i / n_act, kl = i % n_act;
                    const double *mu = s.mu + kl * D, *pp = s.pp + kl * D, *xr = s.xs + bb * D;
                    const double iv = s.iv[kl];
                    const double acc = pairwise_sum_lanes16<double>(
                        [&](int d) { return pred_term_diag(mu[d], pp[d], xr[d], iv); }, D, hmask, jl);
                    if (jl == 0)
                        s.vt[bb * per + kl] = m.lms * (s.pl[kl] - log_norm) + pred_value(true, acc, c0, s.lpp[kl], s.cst[kl], s.hv[kl]);
                }
            } else {
                for (int i = tid; i < nb * n_act; i += GB_THREADS) {
                    const int bb = i / n_act, kl = i % n_act;
                    const double *mu = s.mu + kl * D, *pp = s.pp + kl * D, *xr = s.xs + bb * D;
                    const double acc = pairwise_sum_max256<double>([&](int d) { return pred_term_fixed(mu[d], pp[d], xr[d]); }, D);
                    s.vt[bb * per + kl] = m.lms * (s.pl[kl] - log_norm) + pred_value(false, acc, c0, s.lpp[kl], 0., 0.);
                }
            }
            __syncthreads();
            tick(12);
            for (int bb = tid; bb < nb; bb += GB_THREADS) {
                double mx = neg_inf(), t = 0.0;
                for (int kl = 0; kl < n_act; ++kl) mx = fmax(mx, s.vt[bb * per + kl]);
                for (int kl = 0; kl < n_act; ++kl) t += exp(s.vt[bb * per + kl] - mx);
                p.part_m[(size_t)b * p.M_cap + s0 + bb] = mx;
                p.part_t[(size_t)b * p.M_cap + s0 + bb] = t;
            }
            // log_prior of the candidates this CTA will combine (slot % G == b), one warp each
            for (int bb = warp; bb < nb; bb += GB_THREADS / 32) {
                if ((s0 + bb) % G != b || s.sid[s0 + bb] < 0) continue;
                const double lp = warp_log_prior(s.xs + bb * D);
                if (lane == 0) p.seg_prior[s0 + bb] = lp;
            }
            __syncthreads();
            tick(13);
        }
        tick(1);
        grid_barrier(p.bar, G, (it << 8) | 0x30);
        tick(9);
        // every CTA has taken its snapshot: the removed tokens can now be marked unassigned
        // (not earlier: a slower CTA would read an already cleared token list)
        if (b == G - 1) for (int j = tid; j < N; j += GB_THREADS) if (s.tok[j] >= 0 && s.tk[j] >= 0) m.assignments[s.tok[j]] = -1;
        if (b == 0) for (int j = tid; j < N; j += GB_THREADS) c.tok_id[off + j] = -1;
        // combine: CTA (slot % G) owns the slot; warp 0, lanes over CTAs
        for (int slot = b; slot < n_slots; slot += G) {
            if (warp == 0) {
                const int id = s.sid[slot];
                const double du = c.seg_dur[off * S + slot];
                double out = neg_inf();
                if (id >= 0 && du == du) {                       // uniform over the warp
                    const int n_empty = KM - K;
                    const double e = m.lms * (log_empty - log_norm) + __ldcg(p.seg_prior + slot);
                    double gm = (n_empty > 0) ? e : neg_inf();
                    // the G partial (max, sum) pairs: all loads of a lane issued together (G <= 160 -> <= 5 per lane)
                    double pm[5], pt[5];
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        const int bb = lane + 32 * q;
                        pm[q] = (bb < G) ? __ldcg(p.part_m + (size_t)bb * p.M_cap + slot) : neg_inf();
                        pt[q] = (bb < G) ? __ldcg(p.part_t + (size_t)bb * p.M_cap + slot) : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < 5; ++q) gm = fmax(gm, pm[q]);
                    gm = warp_max(gm);
                    double t = 0.0;
#pragma unroll
                    for (int q = 0; q < 5; ++q)
                        if (pm[q] > neg_inf()) t += pt[q] * exp(pm[q] - gm);
                    t = warp_sum(t);
                    if (n_empty > 0) t += n_empty * exp(e - gm);
                    double val = log(t) + gm;
                    val *= (p.tpt == 1.0) ? du : pow(du, p.tpt);
                    out = val + p.wip;
                }
                if (lane == 0) p.scores[slot] = out;
            }
        }
        tick(2);
        grid_barrier(p.bar, G, (it << 8) | 0x40);
        tick(10);

        // ================= DP: every CTA runs it on the same scores (:653-864)
        for (int i = tid; i < n_slots; i += GB_THREADS) s.sc[i] = __ldcg(p.scores + i);
        __syncthreads();
        double total = 0.0;
        int dp_status = SEGB_DP_OK, used = 0;
        if (warp == 0 && N > 0) {
            DpParams dp;
            dp.S = S; dp.n_min = c.n_slices_min; dp.n_max = c.n_slices_max; dp.mode = p.fb_mode;
            dp.log_p_continue = 0.0; dp.anneal_temp = p.anneal_temp; dp.uniforms = p.uniforms;
            dp_warp_body(dp, s.sc, s.bo, N, s.al, nullptr, u_pos, total, dp_status, used);
            if (lane == 0) { s.red[36] = total; s.red[37] = (double)dp_status; s.red[38] = (double)used; }
        }
        __syncthreads();
        if (N > 0) { total = s.red[36]; dp_status = (int)s.red[37]; used = (int)s.red[38]; }
        u_pos += used;
        if (b == 0) {
            if (tid == 0) {
                p.log_probs[it] = (dp_status == SEGB_DP_OK) ? total : CUDART_NAN;
                p.status[it] = dp_status;
            }
            for (int j = tid; j < N; j += GB_THREADS) c.bounds[off + j] = s.bo[j];
        }
        __syncthreads();
        tick(3);
        if (dp_status != SEGB_DP_OK) continue;                   // uniform over the grid

        // ================= assign the new tokens left to right (:339-349)
        // the token list (embedding id, slot) is formed once and the tokens' embeddings are staged together
        // (one memory latency per utterance instead of one per token); s.tok / s.tk are free again here
        if (tid == 0) {
            int n_t = 0, j_prev = 0;
            for (int j = 0; j < N; ++j) {
                if (!s.bo[j]) continue;
                const int t = j + 1, l = t - j_prev;
                j_prev = j + 1;
                const int slot = (t - 1) * S + (l - 1);
                const int id = (l <= S) ? s.sid[slot] : -1;
                if (b == 0) c.tok_id[off + j] = id;
                if (id < 0) continue;                             // back-tracking leftovers are skipped (:340-342)
                s.tok[n_t] = id; s.tk[n_t] = slot; ++n_t;
            }
            s.red[35] = (double)n_t;
        }
        __syncthreads();
        const int n_new = (int)s.red[35];
        lm_prev = -1;
        lm_sum_a = __dadd_rn((double)n_total, p.lm.a);        // sum_ints(unigram_counts) + a: the counts are tied
        for (int r0 = 0; r0 < n_new; r0 += xb) {
            const int nr = min(xb, n_new - r0);
            for (int r = warp; r < nr; r += GB_THREADS / 32) {
                const int id = s.tok[r0 + r];
                for (int d = lane; d < D; d += 32) s.xs[r * D + d] = fv_x(m, id, d);
            }
            // log_prior of the tokens' segments (computed in the scoring phase): fetched with the embeddings,
            // not one L2 round trip per token on the critical path (s.al is free after the DP)
            for (int r = tid; r < nr; r += GB_THREADS) s.al[r] = __ldcg(p.seg_prior + s.tk[r0 + r]);
            __syncthreads();
            for (int r = 0; r < nr; ++r) {
                x_tok = s.xs + r * D;
                const int k_new = assign_one(s.tok[r0 + r], s.al[r], (it << 8) | ((r0 + r) << 16), -1);
                if (HAS_LM) {
                    lm_prev = k_new;
                    if (tid == 0) s.tk[r0 + r] = k_new;           // the slot index is no longer needed
                }
            }
        }
        x_tok = s.xs;
        if (HAS_LM) {
            __syncthreads();
            if (b == 0 && tid == 0) {                           // counts_from_utterance over the new transcript (:499)
                int jp = -1;
                for (int r = 0; r < n_new; ++r) {
                    const int i = s.tk[r];
                    p.lm.unigram_counts[i] += 1;
                    if (jp >= 0) p.lm.bigram_counts[(size_t)jp * p.lm.K + i] += 1;
                    jp = i;
                }
            }
            lm_prev = -1;
        }
    }
    if (b == 0 && tid == 0) {
        *m.K = K;
        *m.n_total = n_total;
        if (p.u_counter) *p.u_counter = u_pos;
    }
}

// CTAs of one cooperative sweep: one per SM, at most one per component; segb_gibbs_set_max_ctas() lowers the
// limit so that several independent chains (replicas) can share the GPU, each on its own stream
static int g_max_ctas = 0;
static inline int gibbs_grid(int K_max, int n_sm) {
    int g = K_max < n_sm ? K_max : n_sm;
    if (g_max_ctas > 0 && g_max_ctas < g) g = g_max_ctas;
    return g;
}

// development aid: per-phase clock totals of CTA 0 (enabled by segb_debug_gibbs_prof(…, 1))
static unsigned long long *g_prof = nullptr;
// test aid: the value the arrival counter starts from (default 0); a value just below 2^32 makes the
// counter wrap after a few barriers (tests/test_gpu_parity.py::test_gibbs_barrier_counter_wrap)
static unsigned g_bar_base = 0;
static int init_barrier(unsigned *bar, unsigned *base_out, cudaStream_t st) {
    SEGB_CUDA(cudaMemsetAsync(bar, 0, 2048, st));
    *base_out = g_bar_base;
    if (g_bar_base) SEGB_CUDA(cudaMemcpyAsync(bar, &g_bar_base, sizeof(unsigned), cudaMemcpyHostToDevice, st));
    return 0;
}

}  // namespace segb

using namespace segb;

extern "C" int64_t segb_gibbs_work_bytes(int32_t K_max, int32_t N_max, int32_t S) {
    const int64_t M_cap = (int64_t)N_max * S, G = 160;
    return 2048 + 8 * (2 * G * M_cap + 2 * M_cap + 2 * (int64_t)K_max) + 256;
}

static int launch_gibbs_coop(const segb_fixedvar *m, const segb_bigram_lm *lm, const segb_corpus *c,
                             const int32_t *d_order, int32_t n_order, int32_t fb_mode, double time_power_term,
                             double wip, double anneal_temp, int32_t anneal_gibbs_am, const double *uniforms,
                             int64_t *u_counter, void *work, double *log_probs, int32_t *status, void *stream) {
    SEGB_CHECK_ARG(m && c && d_order && work && log_probs && status, "null pointer");
    SEGB_CHECK_ARG(fb_mode == SEGB_DP_FFBS || fb_mode == SEGB_DP_VITERBI_GMM, "fb_mode");
    SEGB_CHECK_ARG(fb_mode == SEGB_DP_VITERBI_GMM || (uniforms && u_counter), "FFBS needs uniforms");
    SEGB_CHECK_ARG(c->tok_id && c->bounds, "corpus needs bounds and tok_id");
    if (n_order == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int n_sm = 0, coop = 0;
    { const int rc = device_info(nullptr, &n_sm, &coop); if (rc) return rc; }
    if (!coop || n_sm > 160) { set_error("cooperative launch unavailable"); return SEGB_E_UNSUPPORTED; }
    if (m->D > 256) { set_error("the cooperative Gibbs sweep covers D <= 256"); return SEGB_E_UNSUPPORTED; }
    GibbsParams p;
    p.m = *m; p.c = *c; p.order = d_order; p.n_order = n_order; p.fb_mode = fb_mode;
    p.assign_mode = (fb_mode == SEGB_DP_FFBS) ? 0 : 1;
    p.item_mode = 0;
    p.has_lm = lm ? 1 : 0;
    if (lm) p.lm = *lm; else memset(&p.lm, 0, sizeof(p.lm));
    p.tpt = time_power_term; p.wip = wip; p.anneal_temp = anneal_temp;
    p.assign_temp = anneal_gibbs_am ? anneal_temp : 1.0;
    p.uniforms = uniforms; p.u_counter = u_counter; p.log_probs = log_probs; p.status = status;
    const int G = gibbs_grid(m->K_max, n_sm);
    p.per = (m->K_max + G - 1) / G;
    p.M_cap = c->N_max * c->S;
    p.xb = (p.per > 12) ? 8 : 32;
    const size_t smem = gibbs_smem_bytes(m->D, m->K_max, p.per, p.xb, p.M_cap, c->N_max);
    if (smem > 227 * 1024 - 16) { set_error("model too large for the persistent Gibbs sweep (%zu bytes of shared memory)", smem); return SEGB_E_UNSUPPORTED; }
    unsigned char *w = (unsigned char *)work;
    p.bar = (unsigned *)w; w += 2048;                 // count, generation, pad, per-CTA trace tags
    p.part_m = (double *)w; w += 8 * (size_t)G * p.M_cap;
    p.part_t = (double *)w; w += 8 * (size_t)G * p.M_cap;
    p.seg_prior = (double *)w; w += 8 * (size_t)p.M_cap;
    p.scores = (double *)w; w += 8 * (size_t)p.M_cap;
    p.v = (double *)w;
    p.prof = g_prof;
    { const int rc = init_barrier(p.bar, &p.bar_base, st); if (rc) return rc; }
    const void *kern = lm ? (const void *)fv_gibbs_kernel<false, true> : (const void *)fv_gibbs_kernel<false, false>;
    SEGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {&p};
    SEGB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(G), dim3(GB_THREADS), args, smem, st));
    count_launch();
    return 0;
}

extern "C" int segb_gibbs_sweep_fixedvar_coop(const segb_fixedvar *m, const segb_corpus *c, const int32_t *d_order,
                                              int32_t n_order, int32_t fb_mode, double time_power_term, double wip,
                                              double anneal_temp, int32_t anneal_gibbs_am, const double *uniforms,
                                              int64_t *u_counter, void *work, double *log_probs, int32_t *status,
                                              void *stream) {
    return launch_gibbs_coop(m, nullptr, c, d_order, n_order, fb_mode, time_power_term, wip, anneal_temp, anneal_gibbs_am,
                             uniforms, u_counter, work, log_probs, status, stream);
}

extern "C" int segb_gibbs_sweep_bigram_coop(const segb_fixedvar *m, const segb_bigram_lm *lm, const segb_corpus *c,
                                            const int32_t *d_order, int32_t n_order, double time_power_term, double wip,
                                            double anneal_temp, int32_t anneal_gibbs_am, const double *uniforms,
                                            int64_t *u_counter, void *work, double *log_probs, int32_t *status,
                                            void *stream) {
    SEGB_CHECK_ARG(lm && m && lm->K == m->K_max && lm->unigram_counts && lm->bigram_counts, "the LM covers the K_max component labels");
    SEGB_CHECK_ARG(m->model == SEGB_MODEL_FIXEDVAR, "bigram sampling: fixed-variance components");
    return launch_gibbs_coop(m, lm, c, d_order, n_order, SEGB_DP_FFBS, time_power_term, wip, anneal_temp, anneal_gibbs_am,
                             uniforms, u_counter, work, log_probs, status, stream);
}

extern "C" int segb_fbgmm_gibbs_items_coop(const segb_fixedvar *m, const int32_t *d_items, int32_t n_items,
                                           double anneal_temp, const double *uniforms, int64_t *u_counter,
                                           void *work, void *stream) {
    SEGB_CHECK_ARG(m && d_items && uniforms && u_counter && work, "null pointer");
    if (n_items == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int coop = 0, n_sm = 0;
    { const int rc = device_info(nullptr, &n_sm, &coop); if (rc) return rc; }
    if (!coop || n_sm > 160) { set_error("cooperative launch unavailable"); return SEGB_E_UNSUPPORTED; }
    if (m->D > 256) { set_error("the cooperative Gibbs sweep covers D <= 256"); return SEGB_E_UNSUPPORTED; }
    GibbsParams p;
    memset(&p, 0, sizeof(p));
    p.m = *m;
    p.c.N_max = 1; p.c.S = 1;                        // no corpus: only the shared-memory carve-up reads these
    p.order = d_items; p.n_order = n_items; p.item_mode = 1;
    p.fb_mode = SEGB_DP_FFBS; p.assign_mode = 0;
    p.tpt = 1.0; p.wip = 0.0; p.anneal_temp = anneal_temp; p.assign_temp = anneal_temp;
    p.uniforms = uniforms; p.u_counter = u_counter;
    const int G = gibbs_grid(m->K_max, n_sm);
    p.per = (m->K_max + G - 1) / G;
    p.M_cap = 1;
    p.xb = 1;
    const size_t smem = gibbs_smem_bytes(m->D, m->K_max, p.per, p.xb, p.M_cap, 1);
    if (smem > 227 * 1024 - 16) { set_error("model too large for the persistent Gibbs sweep (%zu bytes of shared memory)", smem); return SEGB_E_UNSUPPORTED; }
    unsigned char *w = (unsigned char *)work;
    p.bar = (unsigned *)w; w += 2048;
    p.part_m = p.part_t = p.seg_prior = p.scores = nullptr;
    p.v = (double *)w;
    { const int rc = init_barrier(p.bar, &p.bar_base, st); if (rc) return rc; }
    SEGB_CUDA(cudaFuncSetAttribute(fv_gibbs_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {&p};
    SEGB_CUDA(cudaLaunchCooperativeKernel((const void *)fv_gibbs_kernel<true, false>, dim3(G), dim3(GB_THREADS), args, smem, st));
    count_launch();
    return 0;
}

extern "C" int segb_gibbs_set_max_ctas(int32_t max_ctas) {
    g_max_ctas = max_ctas > 0 ? max_ctas : 0;
    return 0;
}

extern "C" int segb_debug_gibbs_bar_base(uint32_t base) {
    g_bar_base = base;
    return 0;
}

// Development aid (tools/gibbs_phases.py): enable != 0 allocates / zeroes the phase counters and makes the
// next cooperative segmenter sweeps accumulate CTA 0's clocks per phase; out16 (host) receives the totals.
extern "C" int segb_debug_gibbs_prof(unsigned long long *out16, int enable) {
    if (enable && !g_prof) SEGB_CUDA(cudaMalloc(&g_prof, 16 * sizeof(unsigned long long)));
    if (out16 && g_prof) SEGB_CUDA(cudaMemcpy(out16, g_prof, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (g_prof) SEGB_CUDA(cudaMemset(g_prof, 0, 16 * sizeof(unsigned long long)));
    if (!enable && g_prof) { cudaFree(g_prof); g_prof = nullptr; }
    return 0;
}
