// Host-side ingestion helpers (no device code): the parts of corpus set-up that are sequential by
// construction in the reference and too slow as Python loops at 200k utterances.
#include <cmath>
#include "common.cuh"

// Random boundary initialisation of Utterances.__init__ (utterances.py:136-157), utterance by utterance:
//   repeat: boundaries[u, 0:N] = (np.random.rand(N) < p); boundaries[u, N-1] = True
//   until some chosen segment carries an embedding AND (max span <= n_slices_max and min span >=
//   n_slices_min, or N <= n_slices_min).
// `uniforms` is a block drawn in advance from np.random (rand(N) consumes N doubles in order, so a block
// consumed sequentially reproduces the reference's stream).  ids: the packed-triangular vec_ids of all
// utterances back to back, packed_off [U + 1] their starts; pos_off [U + 1] the landmark offsets.
// bounds_out [sum N].  Returns the number of uniforms consumed, or -1 if the block ran out (the caller
// draws a larger one from the same generator state and calls again).
extern "C" int64_t segb_host_init_boundaries(const int64_t *lengths, int64_t n_utt, const int64_t *packed_off,
                                             const int64_t *ids, const int64_t *pos_off, const double *uniforms,
                                             int64_t n_uniforms, double p_boundary, int64_t n_slices_min,
                                             int64_t n_slices_max, uint8_t *bounds_out) {
    int64_t used = 0;
    for (int64_t u = 0; u < n_utt; ++u) {
        const int64_t N = lengths[u];
        uint8_t *b = bounds_out + pos_off[u];
        const int64_t *vid = ids + packed_off[u];
        const int64_t n_packed = packed_off[u + 1] - packed_off[u];
        for (;;) {
            if (used + N > n_uniforms) return -1;
            for (int64_t j = 0; j < N; ++j) b[j] = uniforms[used + j] < p_boundary ? 1 : 0;
            used += N;
            b[N - 1] = 1;
            bool any_embed = false;
            int64_t j_prev = 0, span_max = 0, span_min = N + 1;
            for (int64_t j = 0; j < N; ++j) {
                if (!b[j]) continue;
                const int64_t k = (j + 1) * j / 2 + j_prev;        // packed slot of the segment [j_prev, j + 1)
                if (k < n_packed && vid[k] != -1) any_embed = true;
                const int64_t span = j + 1 - j_prev;
                if (span > span_max) span_max = span;
                if (span < span_min) span_min = span;
                j_prev = j + 1;
            }
            if (!any_embed) continue;                               // don't allow all disregarded embeddings
            if ((span_max <= n_slices_max && span_min >= n_slices_min) || N <= n_slices_min) break;
        }
    }
    return used;
}
