// Frozen-state sweep: the model-update half that is common to the k-means and the FBGMM batch modes.
//
// A frozen sweep scores and segments every utterance against the same model, then applies the
// reference's update rules to the chosen tokens as a batch (the reference's own frozen step is
// KMeans.fit, kmeans.py:124-171).  Two of those rules are sequential in token order:
//   * add_item's / FBGMM's clamp of a choice beyond the active components: `if k > K: k = K`, and a
//     choice of slot K opens a new component (kmeans_components.py:103-106, fbgmm.py:459-460,
//     gaussian_components_fixedvar.py:162-165);
//   * clean_components' swap-with-last deletions (kmeans_components.py:149-166, :263-266).
// Round 1 resolved both on the host (Python loops, all_gather_object).  Here they are device kernels over
// fixed-size buffers: an ordered compaction of the affected tokens (normally few), ONE serial pass over
// that list by a single warp, and parallel application.  Across ranks the lists are exchanged with a
// fixed-size all-gather (rank order = global token order) and every rank replays the same serial pass.
#include "common.cuh"

namespace segb {

// ---- tokens of the current boundaries: tok_id[p] = embedding id of the token ending at landmark p
__global__ void tokens_from_bounds_kernel(segb_corpus c, int utt_first, int n_utt) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_utt) return;
    const int u = utt_first + w;
    const int64_t off = c.pos_off[u];
    const int N = (int)(c.pos_off[u + 1] - off);
    int j_prev = 0;
    for (int j = 0; j < N; ++j) {
        int id = -1;
        if (c.bounds[off + j]) {
            const int l = j + 1 - j_prev;
            j_prev = j + 1;
            if (l <= c.S) id = c.seg_id[(off + j) * c.S + (l - 1)];
        }
        c.tok_id[off + j] = id;
    }
}

// ---- ordered compaction of the tokens whose choice lies beyond the active components
constexpr int CMP_THREADS = 256, CMP_PER_THREAD = 16, CMP_BLOCK = CMP_THREADS * CMP_PER_THREAD;

__device__ __forceinline__ bool is_new(const segb_corpus &c, const int32_t *choice, int64_t pos, int K_before, int &id,
                                       int &j) {
    id = c.tok_id[pos];
    if (id < 0) return false;
    j = choice[id];
    return j >= K_before;
}

__global__ void __launch_bounds__(CMP_THREADS) new_count_kernel(segb_corpus c, int64_t pos_first, int64_t n_positions,
                                                                const int32_t *choice, int K_before, int32_t *block_cnt) {
    __shared__ int red[CMP_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * CMP_BLOCK + (int64_t)threadIdx.x * CMP_PER_THREAD;
    int n = 0;
    for (int i = 0; i < CMP_PER_THREAD; ++i) {
        const int64_t p = base + i;
        int id, j;
        if (p < n_positions && is_new(c, choice, pos_first + p, K_before, id, j)) ++n;
    }
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(FULL, n, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < CMP_THREADS / 32; ++i) t += red[i];
        block_cnt[blockIdx.x] = t;
    }
}

// exclusive scan of the block counts (single CTA) -> block_off, total -> n_list[0]
__global__ void __launch_bounds__(1024) new_scan_kernel(const int32_t *block_cnt, int n_blocks, int32_t *block_off,
                                                        int32_t *n_list) {
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int lo = 0; lo < n_blocks; lo += 1024) {
        const int i = lo + threadIdx.x;
        const int v = i < n_blocks ? block_cnt[i] : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if ((threadIdx.x & 31) >= o) inc += t; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
        __syncthreads();
        int wb = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wb += wsum[w];
        const int c0 = carry;
        if (i < n_blocks) block_off[i] = c0 + wb + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c0 + wb + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_list[0] = carry;
}

__global__ void __launch_bounds__(CMP_THREADS) new_write_kernel(segb_corpus c, int64_t pos_first, int64_t n_positions,
                                                                const int32_t *choice, int K_before,
                                                                const int32_t *block_off, int32_t cap, int32_t *list_j,
                                                                int32_t *list_id) {
    __shared__ int wsum[CMP_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * CMP_BLOCK + (int64_t)threadIdx.x * CMP_PER_THREAD;
    int ids[CMP_PER_THREAD], js[CMP_PER_THREAD], n = 0;
    for (int i = 0; i < CMP_PER_THREAD; ++i) {
        const int64_t p = base + i;
        int id, j;
        if (p < n_positions && is_new(c, choice, pos_first + p, K_before, id, j)) { ids[n] = id; js[n] = j; ++n; }
    }
    int inc = n;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if ((threadIdx.x & 31) >= o) inc += t; }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    int wb = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wb += wsum[w];
    int dst = block_off[blockIdx.x] + wb + inc - n;
    for (int i = 0; i < n; ++i, ++dst)
        if (dst < cap) { list_j[dst] = js[i]; list_id[dst] = ids[i]; }
}

// The serial rule over the concatenated lists of all ranks (rank order = global token order):
//   k = min(j, K); if k == K: K += 1            (add_item, kmeans_components.py:103-106)
// One warp: lanes fetch 32 entries at a time, lane 0 walks them.  Writes the resolved components of rank
// `my_rank`'s entries to out_k and the final K to K_out.  lists: [world][cap], counts: [world] (entries
// beyond cap were dropped by the writer: reported through overflow[0]).
__global__ void __launch_bounds__(32) clamp_serial_kernel(const int32_t *lists, const int32_t *counts, int world, int cap,
                                                          int my_rank, int K_before, int K_max, int32_t *out_k,
                                                          int32_t *K_out, int32_t *overflow) {
    __shared__ int buf[32];
    const int lane = threadIdx.x;
    int K = K_before;
    bool over = false;
    for (int r = 0; r < world; ++r) {
        int n = counts[r];
        if (n > cap) { over = true; n = cap; }
        const int32_t *lst = lists + (size_t)r * cap;
        for (int lo = 0; lo < n; lo += 32) {
            const int i = lo + lane;
            const int j = i < n ? lst[i] : 0;
            __syncwarp();
            buf[lane] = j;
            __syncwarp();
            if (lane == 0) {
                const int m = min(32, n - lo);
                for (int q = 0; q < m; ++q) {
                    int k = buf[q];
                    if (k > K) k = K;
                    if (k == K && K < K_max) ++K;
                    buf[q] = k;
                }
            }
            __syncwarp();
            if (r == my_rank && i < n) out_k[i] = buf[lane];
        }
    }
    if (lane == 0) { *K_out = K; if (overflow) overflow[0] = over ? 1 : 0; }
}

__global__ void clamp_apply_kernel(const int32_t *list_id, const int32_t *out_k, const int32_t *n_list, int cap,
                                   int32_t *choice) {
    const int n = min(n_list[0], cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) choice[list_id[i]] = out_k[i];
}

// ---- FBGMM frozen update: statistics from (sum_x, cnt) in closed form

// per-token accumulation: sum_x[k] += X[id] (float64 atomics), cnt[k] += 1, k = choice[id]
__global__ void __launch_bounds__(256) fv_collect_kernel(const float *X, int D, segb_corpus c, int64_t pos_first,
                                                         int64_t n_positions, const int32_t *choice, double *sum_x,
                                                         unsigned long long *cnt) {
    // a lane per position: token ids and choices of 32 positions are fetched side by side and the token rows pulled
    // into L2 as soon as their ids are known (a warp per position walked the chain id -> choice -> row serially);
    // then the warp adds the rows one token at a time
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = w * 32; base < n_positions; base += nw * 32) {
        const int64_t p = base + lane;
        const int id = p < n_positions ? c.tok_id[pos_first + p] : -1;
        int k = -1;
        if (id >= 0) {
            const char *row = reinterpret_cast<const char *>(X + (size_t)id * D);
            for (int o = 0; o < 4 * D; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + o));
            k = choice[id];
            atomicAdd(&cnt[k], 1ull);
        }
        unsigned tm = __ballot_sync(0xffffffffu, id >= 0);
        while (tm) {
            const int src = __ffs(tm) - 1;
            tm &= tm - 1;
            const int tid = __shfl_sync(0xffffffffu, id, src), tk = __shfl_sync(0xffffffffu, k, src);
            for (int d = lane; d < D; d += 32) atomicAdd(&sum_x[(size_t)tk * D + d], (double)X[(size_t)tid * D + d]);
        }
    }
}

// order-preserving relabelling (FBGMM.setup_components -> make_consecutive, fbgmm.py:124-128): new
// label of component k = number of non-empty components before it; K_new = number of non-empty ones.
__global__ void __launch_bounds__(1024) fv_compact_labels_kernel(const long long *cnt, int K_max, int32_t *new_label,
                                                                 int32_t *K_new, long long *n_total) {
    __shared__ int wsum[32];
    __shared__ int carry;
    __shared__ long long tot;
    if (threadIdx.x == 0) { carry = 0; tot = 0; }
    __syncthreads();
    for (int lo = 0; lo < K_max; lo += 1024) {
        const int k = lo + threadIdx.x;
        const long long n = k < K_max ? cnt[k] : 0;
        const int v = n > 0 ? 1 : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if ((threadIdx.x & 31) >= o) inc += t; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
        __syncthreads();
        int wb = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wb += wsum[w];
        const int c0 = carry;
        if (k < K_max) new_label[k] = v ? c0 + wb + inc - v : -1;
        if (n > 0) atomicAdd((unsigned long long *)&tot, (unsigned long long)n);
        __syncthreads();
        if (threadIdx.x == 1023) carry = c0 + wb + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) { *K_new = carry; *n_total = tot; }
}

// Statistics of every slot from the reduced sums (one block per OLD label k):
//   mu_N_num = precision_0*mu_0 + precision*sum_x      precision_N = precision_0 + n*precision
//   precision_pred = precision_N*precision/(precision_N + precision)   mu_N = mu_N_num/precision_N
//   log_prod_precision_pred = sum_d log precision_pred (NumPy pairwise order)
// (gaussian_components_fixedvar.py:153-170, :317-325 in closed form: the sums are order-free, so N ranks
// and one rank produce the same bits).  Slots >= K_new are zeroed.
__global__ void __launch_bounds__(128) fv_set_stats_kernel(segb_fixedvar m, const double *sum_x, const long long *cnt,
                                                           const int32_t *new_label, const int32_t *K_new) {
    extern __shared__ double tmp[];
    const int D = m.D, KM = m.K_max;
    const int k = blockIdx.x;
    const long long n = cnt[k];
    const int Kn = *K_new;
    // zero the slots nobody writes: slot s >= K_new, handled by block s
    if (k >= Kn) {
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const size_t o = (size_t)d * KM + k;
            m.mu_N_numT[o] = 0.; m.prec_NT[o] = 0.; m.prec_predT[o] = 0.; m.mu_NT[o] = 0.;
        }
        if (threadIdx.x == 0) { m.log_prod_prec_pred[k] = 0.; m.counts[k] = 0; }
    }
    if (n <= 0) return;
    const int s = new_label[k];
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const size_t o = (size_t)d * KM + s;
        const double pr = m.precision[d], p0 = m.precision_0[d];
        const double num = __dadd_rn(__dmul_rn(p0, m.mu_0[d]), __dmul_rn(pr, sum_x[(size_t)k * D + d]));
        const double pN = __dadd_rn(p0, __dmul_rn((double)n, pr));
        const double pp = __ddiv_rn(__dmul_rn(pN, pr), __dadd_rn(pN, pr));
        m.mu_N_numT[o] = num;
        m.prec_NT[o] = pN;
        m.prec_predT[o] = pp;
        m.mu_NT[o] = __ddiv_rn(num, pN);
        tmp[d] = log(pp);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        m.log_prod_prec_pred[s] = pairwise_sum<double>([&](int i) { return tmp[i]; }, D);
        m.counts[s] = (int)n;
    }
}

// assignments of the sweep's tokens: assignments[id] = new_label[choice[id]]
__global__ void fv_relabel_kernel(segb_corpus c, int64_t pos_first, int64_t n_positions, const int32_t *choice,
                                  const int32_t *new_label, int32_t *assignments) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_positions) return;
    const int id = c.tok_id[pos_first + p];
    if (id >= 0) assignments[id] = new_label[choice[id]];
}

// ---- k-means clean_components replayed on the counts (single thread), then applied in parallel
// slot_src[s] = the component that ends up in slot s after the swap-with-last deletions of the
// emptied components in descending order (kmeans_components.py:263-266 -> :149-166).
__global__ void __launch_bounds__(1024) km_compaction_plan_kernel(const long long *cnt, const int32_t *K_old_p,
                                                                  int32_t *slot_src, int32_t *inv, int32_t *K_new,
                                                                  int32_t *empties) {
    // Parallel part: identity map and the list of emptied components in DESCENDING order (block scan over
    // the reversed index range).  Serial part (thread 0): the swap-with-last replay over that list only --
    // normally empty or a handful of entries, so the kernel costs microseconds, not K dependent loads.
    __shared__ int wsum[32];
    __shared__ int carry;
    const int K_old = *K_old_p;
    if (threadIdx.x == 0) carry = 0;
    for (int k = threadIdx.x; k < K_old; k += blockDim.x) slot_src[k] = k;
    __syncthreads();
    for (int lo = 0; lo < K_old; lo += 1024) {
        const int i = lo + threadIdx.x;                       // position in the reversed range
        const int k = K_old - 1 - i;
        const int v = (i < K_old && cnt[k] == 0) ? 1 : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if ((threadIdx.x & 31) >= o) inc += t; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
        __syncthreads();
        int wb = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wb += wsum[w];
        const int c0 = carry;
        if (v) empties[c0 + wb + inc - 1] = k;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c0 + wb + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int n_empty = carry;
        int K = K_old;
        for (int i = 0; i < n_empty; ++i) {
            const int k = empties[i];
            --K;
            if (k != K) slot_src[k] = slot_src[K];
        }
        K_new[0] = K;
        K_new[1] = K_old;
    }
    __syncthreads();
    const int K = K_new[0];
    for (int k = threadIdx.x; k < K_old; k += blockDim.x) inv[k] = -1;
    __syncthreads();
    for (int s = threadIdx.x; s < K; s += blockDim.x) inv[slot_src[s]] = s;      // old label -> new slot
}

template <typename T>
__global__ void km_compaction_apply_kernel(segb_kmeans m, const int32_t *slot_src, const int32_t *K_new_p,
                                           const double *num_in, const long long *cnt_in, const T *means_in) {
    // rows were snapshotted (num_in / cnt_in / means_in) before this launch; one block per slot
    const int s = blockIdx.x, D = m.D, KM = m.K_max;
    const int K = K_new_p[0], K_old = K_new_p[1];
    T *means = (T *)m.means, *meansT = (T *)m.meansT;
    const T *rnd = (const T *)m.random_means;
    if (s < K) {
        const int src = slot_src[s];
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            m.mean_num[(size_t)s * D + d] = num_in[(size_t)src * D + d];
            const T v = means_in[(size_t)src * D + d];
            means[(size_t)s * D + d] = v;
            meansT[(size_t)d * KM + s] = v;
        }
        if (threadIdx.x == 0) m.counts[s] = (int)cnt_in[src];
    } else if (s < K_old) {
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            m.mean_num[(size_t)s * D + d] = 0.;
            const T v = rnd[(size_t)s * D + d];             // inactive slots hold random data rows again (:166)
            means[(size_t)s * D + d] = v;
            meansT[(size_t)d * KM + s] = v;
        }
        if (threadIdx.x == 0) m.counts[s] = 0;
    }
    if (s == 0 && threadIdx.x == 0) *m.K = K;
}

__global__ void km_relabel_kernel(segb_corpus c, int64_t pos_first, int64_t n_positions, const int32_t *inv,
                                  const int32_t *K_new_p, int32_t *assignments) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_positions) return;
    const int K_old = K_new_p[1];
    const int id = c.tok_id[pos_first + p];
    if (id < 0) return;
    const int a = assignments[id];
    if (a >= 0 && a < K_old) assignments[id] = inv[a];
}

}  // namespace segb

using namespace segb;

extern "C" int segb_tokens_from_bounds(const segb_corpus *c, int32_t utt_first, int32_t n_utt, void *stream) {
    SEGB_CHECK_ARG(c && c->tok_id && c->bounds, "corpus needs bounds and tok_id");
    SEGB_CHECK_ARG(n_utt >= 0 && utt_first >= 0 && utt_first + n_utt <= c->n_utt, "utterance range");
    if (n_utt == 0) return 0;
    tokens_from_bounds_kernel<<<(n_utt + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*c, utt_first, n_utt);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t segb_frozen_new_work_bytes(int64_t n_positions) {
    const int64_t nb = (n_positions + CMP_BLOCK - 1) / CMP_BLOCK;
    return (2 * nb + 64) * (int64_t)sizeof(int32_t);
}

extern "C" int segb_frozen_new_list(const segb_corpus *c, int64_t pos_first, int64_t n_positions, const int32_t *choice,
                                    int32_t K_before, void *work, int32_t cap, int32_t *list_j, int32_t *list_id,
                                    int32_t *n_list, void *stream) {
    SEGB_CHECK_ARG(c && choice && work && list_j && list_id && n_list && cap > 0, "null pointer");
    SEGB_CHECK_ARG(pos_first >= 0 && n_positions >= 0 && pos_first + n_positions <= c->n_pos, "position range");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_positions == 0) { SEGB_CUDA(cudaMemsetAsync(n_list, 0, sizeof(int32_t), st)); return 0; }
    const int nb = (int)((n_positions + CMP_BLOCK - 1) / CMP_BLOCK);
    int32_t *block_cnt = (int32_t *)work, *block_off = block_cnt + nb;
    new_count_kernel<<<nb, CMP_THREADS, 0, st>>>(*c, pos_first, n_positions, choice, K_before, block_cnt);
    SEGB_LAUNCH_CHECK();
    new_scan_kernel<<<1, 1024, 0, st>>>(block_cnt, nb, block_off, n_list);
    SEGB_LAUNCH_CHECK();
    new_write_kernel<<<nb, CMP_THREADS, 0, st>>>(*c, pos_first, n_positions, choice, K_before, block_off, cap, list_j, list_id);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_frozen_clamp(const int32_t *lists, const int32_t *counts, int32_t world, int32_t cap,
                                 int32_t my_rank, int32_t K_before, int32_t K_max, const int32_t *list_id,
                                 int32_t *out_k, int32_t *choice, int32_t *K_out, int32_t *overflow, void *stream) {
    SEGB_CHECK_ARG(lists && counts && list_id && out_k && choice && K_out, "null pointer");
    SEGB_CHECK_ARG(world >= 1 && my_rank >= 0 && my_rank < world && cap > 0, "rank layout");
    cudaStream_t st = (cudaStream_t)stream;
    clamp_serial_kernel<<<1, 32, 0, st>>>(lists, counts, world, cap, my_rank, K_before, K_max, out_k, K_out, overflow);
    SEGB_LAUNCH_CHECK();
    clamp_apply_kernel<<<148, 256, 0, st>>>(list_id, out_k, counts + my_rank, cap, choice);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_fixedvar_frozen_collect(const segb_fixedvar *m, const segb_corpus *c, int64_t pos_first,
                                            int64_t n_positions, const int32_t *choice, double *sum_x, int64_t *cnt,
                                            void *stream) {
    SEGB_CHECK_ARG(m && c && choice && sum_x && cnt, "null pointer");
    SEGB_CHECK_ARG(!m->x_is_f64, "frozen FBGMM sweep: float32 embeddings");
    SEGB_CHECK_ARG(pos_first >= 0 && n_positions >= 0 && pos_first + n_positions <= c->n_pos, "position range");
    if (n_positions == 0) return 0;
    int64_t blocks = (n_positions + 255) / 256;         // a lane per position
    if (blocks > 148 * 16) blocks = 148 * 16;
    fv_collect_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float *)m->X, m->D, *c, pos_first,
                                                                          n_positions, choice, sum_x,
                                                                          (unsigned long long *)cnt);
    SEGB_LAUNCH_CHECK();
    return 0;
}

extern "C" int segb_fixedvar_frozen_update(const segb_fixedvar *m, const segb_corpus *c, int64_t pos_first,
                                           int64_t n_positions, const int32_t *choice, const double *sum_x,
                                           const int64_t *cnt, int32_t *new_label, void *stream) {
    SEGB_CHECK_ARG(m && c && choice && sum_x && cnt && new_label, "null pointer");
    SEGB_CHECK_ARG(m->model == SEGB_MODEL_FIXEDVAR, "fixed-variance model");
    cudaStream_t st = (cudaStream_t)stream;
    fv_compact_labels_kernel<<<1, 1024, 0, st>>>((const long long *)cnt, m->K_max, new_label, m->K, (long long *)m->n_total);
    SEGB_LAUNCH_CHECK();
    fv_set_stats_kernel<<<m->K_max, 128, sizeof(double) * m->D, st>>>(*m, sum_x, (const long long *)cnt, new_label, m->K);
    SEGB_LAUNCH_CHECK();
    if (n_positions > 0) {
        fv_relabel_kernel<<<(unsigned)((n_positions + 255) / 256), 256, 0, st>>>(*c, pos_first, n_positions, choice,
                                                                                 new_label, m->assignments);
        SEGB_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int segb_kmeans_frozen_clean(const segb_kmeans *m, const segb_corpus *c, int64_t pos_first,
                                        int64_t n_positions, const int64_t *cnt, void *work, void *stream) {
    // work: segb_kmeans_frozen_clean_work_bytes() bytes: slot_src [K_max] | inv [K_max] | empties [K_max] | K_new, K_old | snapshots
    SEGB_CHECK_ARG(m && c && cnt && work, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int KM = m->K_max, D = m->D;
    int32_t *slot_src = (int32_t *)work, *inv = slot_src + KM, *empties = inv + KM, *K_new = empties + KM;
    unsigned char *q = (unsigned char *)(K_new + 2);
    q += (16 - ((uintptr_t)q & 15)) & 15;
    double *num_in = (double *)q; q += sizeof(double) * (size_t)KM * D;
    long long *cnt_in = (long long *)q; q += sizeof(long long) * (size_t)KM;
    void *means_in = q;
    const size_t esz = m->x_is_f64 ? 8 : 4;
    SEGB_CUDA(cudaMemcpyAsync(num_in, m->mean_num, sizeof(double) * (size_t)KM * D, cudaMemcpyDeviceToDevice, st));
    SEGB_CUDA(cudaMemcpyAsync(cnt_in, cnt, sizeof(long long) * (size_t)KM, cudaMemcpyDeviceToDevice, st));
    SEGB_CUDA(cudaMemcpyAsync(means_in, m->means, esz * (size_t)KM * D, cudaMemcpyDeviceToDevice, st));
    km_compaction_plan_kernel<<<1, 1024, 0, st>>>((const long long *)cnt, m->K, slot_src, inv, K_new, empties);
    SEGB_LAUNCH_CHECK();
    if (m->x_is_f64)
        km_compaction_apply_kernel<double><<<KM, 128, 0, st>>>(*m, slot_src, K_new, num_in, cnt_in, (const double *)means_in);
    else
        km_compaction_apply_kernel<float><<<KM, 128, 0, st>>>(*m, slot_src, K_new, num_in, cnt_in, (const float *)means_in);
    SEGB_LAUNCH_CHECK();
    if (n_positions > 0) {
        km_relabel_kernel<<<(unsigned)((n_positions + 255) / 256), 256, 0, st>>>(*c, pos_first, n_positions, inv, K_new,
                                                                                 m->assignments);
        SEGB_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int64_t segb_kmeans_frozen_clean_work_bytes(int32_t K_max, int32_t D) {
    return (int64_t)(3 * K_max + 2) * 4 + 16 + (int64_t)K_max * D * 8 + (int64_t)K_max * 8 + (int64_t)K_max * D * 8;
}
