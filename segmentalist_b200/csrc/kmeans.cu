// K-means components: exact scoring (NumPy float32 order), online updates, the
// sequential segmental k-means sweep and the frozen-state sweep helpers.
//
// Replaces (reference paths relative to segmentalist/):
//   KMeansComponents.add_item/del_item/del_component/clean_components  kmeans_components.py:93-166,263-266
//   KMeansComponents.neg_sqrd_norm / max_ / argmax_neg_sqrd_norm_i    kmeans_components.py:225-232
//   SegmentalKMeansWordseg.get_vec_embed_neg_len_sqrd_norms / segment_i
//                                                                     kmeans_acoustic_wordseg.py:225-351
//   KMeans.fit E-step                                                 kmeans.py:124-146
//
// `means` has X's dtype (float32 in practice): the quotient mean_numerators/counts
// is formed in float64 and rounded on store, distances are accumulated in X's
// dtype in NumPy's pairwise order -- reproduced here operation by operation so
// that max / argmax decisions are bit-exact.
#include "common.cuh"

namespace segb {

int launch_dp_local(const segb_corpus *c, int32_t utt, const double *local_scores, int32_t mode,
                    double log_p_continue, double anneal_temp, const double *uniforms, int64_t *u_counter,
                    double *log_prob, int32_t *status, cudaStream_t stream);

template <typename T> __device__ __forceinline__ T neg_inf_t();
template <> __device__ __forceinline__ float neg_inf_t<float>() { return -CUDART_INF_F; }
template <> __device__ __forceinline__ double neg_inf_t<double>() { return -CUDART_INF; }

template <typename T> struct KM {
    static __device__ __forceinline__ const T *X(const segb_kmeans &m) { return (const T *)m.X; }
    static __device__ __forceinline__ T *means(const segb_kmeans &m) { return (T *)m.means; }
    static __device__ __forceinline__ T *meansT(const segb_kmeans &m) { return (T *)m.meansT; }
    static __device__ __forceinline__ const T *rnd(const segb_kmeans &m) { return (const T *)m.random_means; }
};

// means[k, :] = mean_numerators[k, :] / counts[k]   (kmeans_components.py:110)
template <typename T> __device__ void km_store_mean(const segb_kmeans &m, int k, int cnt) {
    for (int d = threadIdx.x; d < m.D; d += blockDim.x) {
        const T v = (T)__ddiv_rn(m.mean_num[(size_t)k * m.D + d], (double)cnt);
        KM<T>::means(m)[(size_t)k * m.D + d] = v;
        KM<T>::meansT(m)[(size_t)d * m.K_max + k] = v;
    }
}

template <typename T> __device__ void km_add_item(const segb_kmeans &m, int id, int k) {
    __syncthreads();
    const int K = *m.K;
    if (k > K) k = K;                           // :103-104
    const int cnt = m.counts[k] + 1;
    __syncthreads();
    for (int d = threadIdx.x; d < m.D; d += blockDim.x) {
        const size_t o = (size_t)k * m.D + d;
        m.mean_num[o] = __dadd_rn(m.mean_num[o], (double)KM<T>::X(m)[(size_t)id * m.D + d]);
    }
    if (threadIdx.x == 0) {
        if (k == K) *m.K = K + 1;
        m.counts[k] = cnt;
        m.assignments[id] = k;
    }
    __syncthreads();
    km_store_mean<T>(m, k, cnt);
    __syncthreads();
}

template <typename T> __device__ void km_del_item(const segb_kmeans &m, int id) {
    __syncthreads();
    const int k = m.assignments[id];
    if (k == -1) return;
    const int cnt = m.counts[k] - 1;
    __syncthreads();
    for (int d = threadIdx.x; d < m.D; d += blockDim.x) {
        const size_t o = (size_t)k * m.D + d;
        m.mean_num[o] = __dsub_rn(m.mean_num[o], (double)KM<T>::X(m)[(size_t)id * m.D + d]);
    }
    if (threadIdx.x == 0) { m.counts[k] = cnt; m.assignments[id] = -1; }
    __syncthreads();
    if (cnt != 0) km_store_mean<T>(m, k, cnt);   // an emptied component keeps its stale mean (:131-132)
    __syncthreads();
}

template <typename T>
__device__ void km_del_component(const segb_kmeans &m, int k, const int32_t *relabel_ids, int64_t relabel_n) {
    __syncthreads();
    const int last = *m.K - 1;
    const int cnt_last = m.counts[last];
    __syncthreads();
    if (k != last) {
        for (int d = threadIdx.x; d < m.D; d += blockDim.x)
            m.mean_num[(size_t)k * m.D + d] = m.mean_num[(size_t)last * m.D + d];
        __syncthreads();
        km_store_mean<T>(m, k, cnt_last);       // :160 recomputes the quotient from the moved numerators
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
    __syncthreads();
    for (int d = threadIdx.x; d < m.D; d += blockDim.x) {
        m.mean_num[(size_t)last * m.D + d] = 0.;
        const T v = KM<T>::rnd(m)[(size_t)last * m.D + d];     // :166
        KM<T>::means(m)[(size_t)last * m.D + d] = v;
        KM<T>::meansT(m)[(size_t)d * m.K_max + last] = v;
    }
    if (threadIdx.x == 0) {
        if (k != last) m.counts[k] = cnt_last;
        m.counts[last] = 0;
        *m.K = last;
    }
    __syncthreads();
}

template <typename T>
__device__ void km_clean(const segb_kmeans &m, const int32_t *relabel_ids, int64_t relabel_n) {
    __syncthreads();
    const int K0 = *m.K;
    for (int k = K0 - 1; k >= 0; --k) {         // np.where(counts[:K] == 0)[0][::-1]
        __syncthreads();
        if (m.counts[k] == 0) km_del_component<T>(m, k, relabel_ids, relabel_n);
    }
}

// -sum_d (means[k,d] - x[d])^2 in T arithmetic, NumPy pairwise order (:225-226)
template <typename T>
__device__ __forceinline__ T km_neg_dist(const T *meansT_k, int KM_, const T *xs, int D) {
    auto f = [&](int d) -> T {
        const T dl = sub_rn<T>(meansT_k[(size_t)d * KM_], xs[d]);
        return mul_rn<T>(dl, dl);
    };
    T s;
    if (D <= 128) s = pairwise_block<T>(f, 0, D);
    else if (D <= 256) {
        int n2 = D / 2; n2 -= n2 % 8;
        s = add_rn<T>(pairwise_block<T>(f, 0, n2), pairwise_block<T>(f, n2, D - n2));
    } else s = pairwise_sum<T>(f, D);
    return -s;
}

// Block-wide (max value, first index) over per-thread candidates. red: >= 40 doubles.
__device__ __forceinline__ void block_argmax(double v, int k, double *red, double &vmax, int &kmax) {
    vmax = block_max(v, red);
    const double cand = (v == vmax) ? (double)k : 4.0e9;
    kmax = (int)(-block_max(-cand, red));
}

// max / first argmax over all K_max slots for the item staged in xs (block-cooperative)
template <typename T>
__device__ void km_best_of(const segb_kmeans &m, const T *xs, double *red, T &best_v, int &best_k) {
    const int KM_ = m.K_max;
    T bv = neg_inf_t<T>();
    int bk = 0x7fffffff;
    // NumPy: max() of an array with a NaN is NaN and argmax() is the index of the FIRST NaN (kmeans_components.py:228-232
    // on an embedding with a NaN element).  A NaN score therefore competes as +inf (scores are <= 0, so +inf is free).
    for (int k = threadIdx.x; k < KM_; k += blockDim.x) {
        T v = km_neg_dist<T>(KM<T>::meansT(m) + k, KM_, xs, m.D);
        if (v != v) v = -neg_inf_t<T>();
        if (v > bv || bk == 0x7fffffff) { bv = v; bk = k; }
    }
    double vmax; int kmax;
    block_argmax((double)bv, bk, red, vmax, kmax);
    best_v = vmax == (double)CUDART_INF ? (T)CUDART_NAN : (T)vmax;
    best_k = kmax;
}

template <typename T>
__global__ void __launch_bounds__(256) km_best_kernel(segb_kmeans m, const int32_t *ids, int64_t n, T *best_val,
                                                      int32_t *best_k) {
    extern __shared__ double smem[];
    double *red = smem;
    T *xs = (T *)(smem + 40);
    for (int64_t it = blockIdx.x; it < n; it += gridDim.x) {
        const int id = ids ? ids[it] : (int)it;
        if (id < 0) { if (threadIdx.x == 0) { best_val[it] = neg_inf_t<T>(); best_k[it] = -1; } continue; }
        __syncthreads();
        for (int d = threadIdx.x; d < m.D; d += blockDim.x) xs[d] = KM<T>::X(m)[(size_t)id * m.D + d];
        __syncthreads();
        T bv; int bk;
        km_best_of<T>(m, xs, red, bv, bk);
        if (threadIdx.x == 0) { best_val[it] = bv; best_k[it] = bk; }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) km_row_kernel(segb_kmeans m, int id, T *out) {
    extern __shared__ double smem[];
    T *xs = (T *)(smem + 40);
    for (int d = threadIdx.x; d < m.D; d += blockDim.x) xs[d] = KM<T>::X(m)[(size_t)id * m.D + d];
    __syncthreads();
    for (int k = threadIdx.x; k < m.K_max; k += blockDim.x)
        out[k] = km_neg_dist<T>(KM<T>::meansT(m) + k, m.K_max, xs, m.D);
}

template <typename T>
__global__ void __launch_bounds__(256) km_add_list_kernel(segb_kmeans m, const int32_t *ids, const int32_t *ks, int n) {
    for (int i = 0; i < n; ++i) km_add_item<T>(m, ids[i], ks[i]);
}
template <typename T>
__global__ void __launch_bounds__(256) km_del_list_kernel(segb_kmeans m, const int32_t *ids, int n) {
    for (int i = 0; i < n; ++i) if (ids[i] >= 0) km_del_item<T>(m, ids[i]);
}
template <typename T>
__global__ void __launch_bounds__(256) km_move_list_kernel(segb_kmeans m, const int32_t *ids, const int32_t *ks, int n) {
    for (int i = 0; i < n; ++i) { km_del_item<T>(m, ids[i]); km_add_item<T>(m, ids[i], ks[i]); }
}
template <typename T>
__global__ void __launch_bounds__(256) km_clean_kernel(segb_kmeans m, const int32_t *relabel_ids, int64_t relabel_n) {
    km_clean<T>(m, relabel_ids, relabel_n);
}

// scores[slot] = float64(best_val[seg_id]) * dur + wip   (kmeans_acoustic_wordseg.py:334-351)
template <typename T>
__global__ void km_band_scores_kernel(segb_corpus c, int64_t slot_lo, int64_t n_slots, const T *best_val, double wip,
                                      double *scores) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    const int64_t slot = slot_lo + i;
    const int id = c.seg_id[slot];
    const double du = c.seg_dur[slot];
    double v = neg_inf();
    if (id >= 0 && du == du) v = __dadd_rn(__dmul_rn((double)best_val[id], du), wip);
    scores[slot] = v;
}

// ---- sequential segment_i pieces

template <typename T>
__global__ void __launch_bounds__(256) km_score_utt_kernel(segb_kmeans m, segb_corpus c, int u, double wip,
                                                           double *local_scores, int32_t *local_arg) {
    extern __shared__ double smem[];
    double *red = smem;
    T *xs = (T *)(smem + 40);
    const int64_t off = c.pos_off[u];
    const int N = (int)(c.pos_off[u + 1] - off);
    const int n_slots = N * c.S;
    for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
        const int id = c.seg_id[off * c.S + slot];
        const double du = c.seg_dur[off * c.S + slot];
        if (id < 0) { if (threadIdx.x == 0) { local_scores[slot] = neg_inf(); local_arg[slot] = -1; } continue; }
        __syncthreads();
        for (int d = threadIdx.x; d < m.D; d += blockDim.x) xs[d] = KM<T>::X(m)[(size_t)id * m.D + d];
        __syncthreads();
        T bv; int bk;
        km_best_of<T>(m, xs, red, bv, bk);
        if (threadIdx.x == 0) {
            local_scores[slot] = (du == du) ? __dadd_rn(__dmul_rn((double)bv, du), wip) : neg_inf();
            local_arg[slot] = bk;
        }
    }
}

// del old tokens, add new tokens with their frozen-means argmax, clean (:312-320)
template <typename T>
__global__ void __launch_bounds__(256) km_update_utt_kernel(segb_kmeans m, segb_corpus c, int u,
                                                            const int32_t *local_arg, int32_t *dp_status) {
    if (*dp_status != SEGB_DP_OK) return;
    const int64_t off = c.pos_off[u];
    const int N = (int)(c.pos_off[u + 1] - off);
    for (int j = 0; j < N; ++j) {
        const int id = c.tok_id[off + j];
        if (id >= 0) km_del_item<T>(m, id);
    }
    __syncthreads();
    int j_prev = 0;
    bool bad = false;
    for (int j = 0; j < N; ++j) {
        const bool b = c.bounds[off + j];
        int id = -1, k = -1;
        if (b) {
            const int t = j + 1, l = t - j_prev;
            j_prev = j + 1;
            if (l <= c.S) { id = c.seg_id[(off + t - 1) * c.S + (l - 1)]; k = local_arg[(t - 1) * c.S + (l - 1)]; }
            if (id < 0) bad = true;     // the reference's add_item asserts on a -1 embedding
        }
        __syncthreads();
        if (threadIdx.x == 0) c.tok_id[off + j] = id;
        if (id >= 0) km_add_item<T>(m, id, k);
    }
    km_clean<T>(m, c.tok_id, c.n_pos);
    if (bad && threadIdx.x == 0) *dp_status = SEGB_DP_INFEASIBLE;
}

// ---- frozen-state sweep pieces: one warp per utterance

// The walk over an utterance's boundaries used to be one position at a time with a chain of four dependent loads per
// token (boundary flag -> segment id -> winner -> embedding row; ncu: ~29 warps stalled on loads per issue).  Now a
// lane owns a position: flags, segment ids and winners of 32 positions are fetched side by side, the token rows are
// pulled into L2 as soon as their ids are known, and the warp then adds them one token at a time.
template <typename T>
__global__ void __launch_bounds__(128) km_collect_kernel(segb_kmeans m, segb_corpus c, int utt_first, int n_utt,
                                                         const int32_t *best_k, double *sum_x,
                                                         unsigned long long *cnt) {
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= n_utt) return;
    const int u = utt_first + w;
    const int64_t off = c.pos_off[u];
    const int N = (int)(c.pos_off[u + 1] - off);
    const unsigned FULLM = 0xffffffffu;
    int j_prev = 0;                                   // start of the open segment, carried across 32-position pieces
    for (int base = 0; base < N; base += 32) {
        const int j = base + lane;
        const bool b = j < N && c.bounds[off + j];
        const unsigned bm = __ballot_sync(FULLM, b);
        int id = -1, k = -1;
        if (b) {
            const unsigned below = bm & ((1u << lane) - 1u);
            const int start = below ? base + (31 - __clz(below)) + 1 : j_prev;
            const int l = j + 1 - start;
            if (l <= c.S) id = c.seg_id[(off + j) * c.S + (l - 1)];
            if (id >= 0) {
                const char *row = reinterpret_cast<const char *>(KM<T>::X(m) + (size_t)id * m.D);
                for (int o = 0; o < (int)sizeof(T) * m.D; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + o));
                k = best_k[id];
            }
        }
        if (j < N) c.tok_id[off + j] = id;
        if (id >= 0) { m.assignments[id] = k; atomicAdd(&cnt[k], 1ull); }
        if (bm) j_prev = base + (31 - __clz(bm)) + 1;
        unsigned tm = __ballot_sync(FULLM, id >= 0);
        while (tm) {                                  // one token at a time, the whole warp on its row
            const int src = __ffs(tm) - 1;
            tm &= tm - 1;
            const int tid = __shfl_sync(FULLM, id, src), tk = __shfl_sync(FULLM, k, src);
            for (int d = lane; d < m.D; d += 32)
                atomicAdd(&sum_x[(size_t)tk * m.D + d], (double)KM<T>::X(m)[(size_t)tid * m.D + d]);
        }
    }
}

template <typename T>
__global__ void km_set_means_kernel(segb_kmeans m, const double *sum_x, const long long *cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)m.K_max * m.D) return;
    const int k = (int)(i / m.D), d = (int)(i % m.D);
    const long long n = cnt[k];
    m.mean_num[i] = sum_x[i];
    if (d == 0) m.counts[k] = (int)n;
    if (n > 0) {
        const T v = (T)__ddiv_rn(sum_x[i], (double)n);
        KM<T>::means(m)[i] = v;
        KM<T>::meansT(m)[(size_t)d * m.K_max + k] = v;
    }
}

// The constructor's add loop (kmeans_components.py:79-81: add_item for every assigned item, component by
// component in index order) for all components at once: block k sums its own members in index order with
// the same separately rounded float64 additions as km_add_item, so mean_numerators, counts and means come
// out bit-identical to the sequential loop.  order / seg_off as in segb_fixedvar_build.
template <typename T>
__global__ void __launch_bounds__(128) km_build_kernel(segb_kmeans m, const int64_t *order, const int64_t *seg_off, int K_new) {
    const int k = blockIdx.x;
    if (k >= K_new) return;
    const int64_t lo = seg_off[k], hi = seg_off[k + 1];
    const int cnt = (int)(hi - lo);
    for (int d = threadIdx.x; d < m.D; d += blockDim.x) {
        double acc = 0.0;
        for (int64_t i = lo; i < hi; ++i) acc = __dadd_rn(acc, (double)KM<T>::X(m)[(size_t)order[i] * m.D + d]);
        const size_t o = (size_t)k * m.D + d;
        m.mean_num[o] = acc;
        if (cnt > 0) {
            const T v = (T)__ddiv_rn(acc, (double)cnt);
            KM<T>::means(m)[o] = v;
            KM<T>::meansT(m)[(size_t)d * m.K_max + k] = v;
        }
    }
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) m.assignments[order[i]] = k;
    if (threadIdx.x == 0) { m.counts[k] = cnt; if (k == 0) *m.K = K_new; }
}

static size_t km_smem(const segb_kmeans *m) { return sizeof(double) * (40 + (size_t)m->D + 8); }

}  // namespace segb

using namespace segb;

#define KM_DISPATCH(m, CALL_F32, CALL_F64) \
    do { if ((m)->x_is_f64) { CALL_F64; } else { CALL_F32; } } while (0)

extern "C" int segb_kmeans_neg_sqrd_norm_row(const segb_kmeans *m, int32_t id, void *out, void *stream) {
    SEGB_CHECK_ARG(m && out && id >= 0 && id < m->n_emb, "item id");
    cudaStream_t st = (cudaStream_t)stream;
    KM_DISPATCH(m, (km_row_kernel<float><<<1, 256, km_smem(m), st>>>(*m, id, (float *)out)),
                (km_row_kernel<double><<<1, 256, km_smem(m), st>>>(*m, id, (double *)out)));
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_kmeans_best(const segb_kmeans *m, const int32_t *ids, int64_t n, void *best_val,
                                int32_t *best_k, void *stream) {
    SEGB_CHECK_ARG(m && best_val && best_k && n >= 0, "null pointer");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (int)(n < 148 * 16 ? n : 148 * 16);
    KM_DISPATCH(m, (km_best_kernel<float><<<blocks, 256, km_smem(m), st>>>(*m, ids, n, (float *)best_val, best_k)),
                (km_best_kernel<double><<<blocks, 256, km_smem(m), st>>>(*m, ids, n, (double *)best_val, best_k)));
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_kmeans_add_items(const segb_kmeans *m, const int32_t *ids, const int32_t *ks, int32_t n,
                                     void *stream) {
    SEGB_CHECK_ARG(m && ids && ks && n >= 0, "null pointer");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    KM_DISPATCH(m, (km_add_list_kernel<float><<<1, 256, 0, st>>>(*m, ids, ks, n)),
                (km_add_list_kernel<double><<<1, 256, 0, st>>>(*m, ids, ks, n)));
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_kmeans_build(const segb_kmeans *m, const int64_t *order, const int64_t *seg_off, int32_t K_new,
                                 void *stream) {
    SEGB_CHECK_ARG(m && order && seg_off && K_new >= 0 && K_new <= m->K_max, "null pointer");
    if (K_new == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    KM_DISPATCH(m, (km_build_kernel<float><<<K_new, 128, 0, st>>>(*m, order, seg_off, K_new)),
                (km_build_kernel<double><<<K_new, 128, 0, st>>>(*m, order, seg_off, K_new)));
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_kmeans_del_items(const segb_kmeans *m, const int32_t *ids, int32_t n, void *stream) {
    SEGB_CHECK_ARG(m && ids && n >= 0, "null pointer");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    KM_DISPATCH(m, (km_del_list_kernel<float><<<1, 256, 0, st>>>(*m, ids, n)),
                (km_del_list_kernel<double><<<1, 256, 0, st>>>(*m, ids, n)));
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_kmeans_move_items(const segb_kmeans *m, const int32_t *ids, const int32_t *ks, int32_t n,
                                      void *stream) {
    SEGB_CHECK_ARG(m && ids && ks && n >= 0, "null pointer");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    KM_DISPATCH(m, (km_move_list_kernel<float><<<1, 256, 0, st>>>(*m, ids, ks, n)),
                (km_move_list_kernel<double><<<1, 256, 0, st>>>(*m, ids, ks, n)));
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_kmeans_clean(const segb_kmeans *m, const int32_t *relabel_ids, int64_t relabel_n, void *stream) {
    SEGB_CHECK_ARG(m, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    KM_DISPATCH(m, (km_clean_kernel<float><<<1, 256, 0, st>>>(*m, relabel_ids, relabel_n)),
                (km_clean_kernel<double><<<1, 256, 0, st>>>(*m, relabel_ids, relabel_n)));
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_kmeans_band_scores(const segb_kmeans *m, const segb_corpus *c, int64_t pos_first,
                                       int64_t n_positions, const void *best_val, double wip, double *scores,
                                       void *stream) {
    SEGB_CHECK_ARG(m && c && best_val && scores, "null pointer");
    SEGB_CHECK_ARG(n_positions >= 0 && pos_first >= 0 && pos_first + n_positions <= c->n_pos, "position range");
    if (n_positions == 0) return 0;
    const int64_t slot_lo = pos_first * c->S, n_slots = n_positions * c->S;
    const int threads = 256;
    const int64_t blocks = (n_slots + threads - 1) / threads;
    cudaStream_t st = (cudaStream_t)stream;
    KM_DISPATCH(m, (km_band_scores_kernel<float><<<(unsigned)blocks, threads, 0, st>>>(*c, slot_lo, n_slots, (const float *)best_val, wip, scores)),
                (km_band_scores_kernel<double><<<(unsigned)blocks, threads, 0, st>>>(*c, slot_lo, n_slots, (const double *)best_val, wip, scores)));
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_kmeans_segment_sweep(const segb_kmeans *m, const segb_corpus *c, const int32_t *h_order,
                                         int32_t n_order, double wip, double *scratch_scores,
                                         float *scratch_best, int32_t *scratch_arg, double *totals,
                                         int32_t *status, void *stream) {
    (void)scratch_best;
    SEGB_CHECK_ARG(m && c && h_order && scratch_scores && scratch_arg && totals && status, "null pointer");
    SEGB_CHECK_ARG(c->tok_id && c->bounds, "corpus needs bounds and tok_id");
    cudaStream_t st = (cudaStream_t)stream;
    const int score_blocks = c->N_max * c->S;
    for (int i = 0; i < n_order; ++i) {
        const int u = h_order[i];
        SEGB_CHECK_ARG(u >= 0 && u < c->n_utt, "utterance index");
        KM_DISPATCH(m, (km_score_utt_kernel<float><<<score_blocks, 256, km_smem(m), st>>>(*m, *c, u, wip, scratch_scores, scratch_arg)),
                    (km_score_utt_kernel<double><<<score_blocks, 256, km_smem(m), st>>>(*m, *c, u, wip, scratch_scores, scratch_arg)));
        SEGB_LAUNCH_CHECK();
        int r = launch_dp_local(c, u, scratch_scores, SEGB_DP_VITERBI_KMEANS, 0.0, 1.0, nullptr, nullptr,
                                totals + i, status + i, st);
        if (r) return r;
        KM_DISPATCH(m, (km_update_utt_kernel<float><<<1, 256, 0, st>>>(*m, *c, u, scratch_arg, status + i)),
                    (km_update_utt_kernel<double><<<1, 256, 0, st>>>(*m, *c, u, scratch_arg, status + i)));
        SEGB_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int segb_kmeans_collect(const segb_kmeans *m, const segb_corpus *c, int32_t utt_first, int32_t n_utt,
                                   const int32_t *best_k, double *sum_x, int64_t *cnt, void *stream) {
    SEGB_CHECK_ARG(m && c && best_k && sum_x && cnt, "null pointer");
    SEGB_CHECK_ARG(n_utt >= 0 && utt_first >= 0 && utt_first + n_utt <= c->n_utt, "utterance range");
    if (n_utt == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (n_utt + 3) / 4;
    KM_DISPATCH(m, (km_collect_kernel<float><<<blocks, 128, 0, st>>>(*m, *c, utt_first, n_utt, best_k, sum_x, (unsigned long long *)cnt)),
                (km_collect_kernel<double><<<blocks, 128, 0, st>>>(*m, *c, utt_first, n_utt, best_k, sum_x, (unsigned long long *)cnt)));
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_kmeans_set_means(const segb_kmeans *m, const double *sum_x, const int64_t *cnt, void *stream) {
    SEGB_CHECK_ARG(m && sum_x && cnt, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = (int64_t)m->K_max * m->D;
    const int threads = 256;
    KM_DISPATCH(m, (km_set_means_kernel<float><<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(*m, sum_x, (const long long *)cnt)),
                (km_set_means_kernel<double><<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(*m, sum_x, (const long long *)cnt)));
    SEGB_LAUNCH_CHECK();
    return 0;
}

