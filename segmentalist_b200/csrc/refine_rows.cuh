// Exact re-scoring of one embedding's surviving candidates by EIGHT lanes, shared by the standalone
// refine kernels (kmeans_mma.cu, fixedvar_filter.cu) and the fused score kernel (score_fused.cu).
#pragma once
#include "mma_common.cuh"

namespace segb {
namespace mma {

// ---- k-means: float32, NumPy's pairwise order (kmeans_components.py:225-226), bit-exact.
// NumPy's pairwise sum keeps 8 running accumulators per block of <= 128 terms.  Lane (g, c) of the
// 8-lane group owns accumulators 2c and 2c+1 of block g and moves them with 8-byte loads;
// (r[2c] + r[2c+1]) locally, then xor-1 and xor-2 shuffles inside the block's four lanes reproduce
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); the second four lanes take the second block when D > 128.
// Needs an even D (8-byte aligned rows) and at most MAXS accumulator steps per block.
struct Row8Geom {
    int g, c, lo_g, n_g, n8_g, steps, gbase;
    unsigned gmask;
    bool two_blocks;
    __device__ __forceinline__ Row8Geom(int D, int lane) {
        const int j = lane & 7;
        g = j >> 2; c = j & 3;
        gmask = 0xffu << (lane & 24);
        gbase = lane & 24;
        int n2 = 0;
        if (D > 128) { n2 = D / 2; n2 -= n2 % 8; }
        lo_g = g == 0 ? 0 : n2;
        n_g = (D > 128) ? (g == 0 ? n2 : D - n2) : (g == 0 ? D : 0);
        n8_g = n_g >= 8 ? n_g - (n_g % 8) : 0;
        steps = n8_g / 8;
        two_blocks = D > 128;
    }
};

// accumulator steps of the longer NumPy block (host and device)
__host__ __device__ inline int row8_steps_max(int D) {
    int n2 = 0;
    if (D > 128) { n2 = D / 2; n2 -= n2 % 8; }
    const int longest = (D > 128) ? (D - n2 > n2 ? D - n2 : n2) : D;
    return longest / 8;
}
// the 8-lane scheme covers NumPy's pairwise structure up to one split with blocks <= 128
__host__ __device__ inline bool row8_supported(int D) {
    int n2 = D / 2; n2 -= n2 % 8;
    return (D % 2 == 0) && (D <= 128 || (D <= 256 && D - n2 <= 128));
}

// code (refine_decide): -1 the best chunk i1 suffices, >= 0 also visit chunk i2.  masks: bits 0-15
// members of chunk i1 to score, bits 16-31 members of chunk i2.  Result on all 8 lanes.
// this lane's elements of the embedding row (independent of the filter record: issue them first)
template <int MAXS>
__device__ __forceinline__ void km_load_x8(const float *xr, const Row8Geom &q, float2 *xv) {
    const float2 *xr2 = reinterpret_cast<const float2 *>(xr + q.lo_g + 2 * q.c);
#pragma unroll
    for (int i = 0; i < MAXS; ++i) xv[i] = (i < q.steps) ? xr2[i * 4] : make_float2(0.f, 0.f);
}
// The block's n % 8 trailing elements (an even count <= 6 for even D: 8-byte loads), requested WITH the row's other
// loads: read one at a time inside km_exact_one8's tail loop they were a second DRAM round trip on the row's
// dependent chain (the row's last sector is touched by no other load).
constexpr int ROW8_TAIL = 3;
__device__ __forceinline__ void km_load_xtail8(const float *xr, const Row8Geom &q, float2 *xt) {
    const float2 *t2 = reinterpret_cast<const float2 *>(xr + q.lo_g + q.n8_g);
    const int n_t = q.n_g - q.n8_g;
#pragma unroll
    for (int t = 0; t < ROW8_TAIL; ++t) xt[t] = (2 * t < n_t) ? t2[t] : make_float2(0.f, 0.f);
}

// exact score of ONE component for the row held in xv (result on all 8 lanes of the group)
template <int MAXS>
__device__ __forceinline__ float km_exact_one8(const float *means, int D, const float *xr, const float2 *xv, int k,
                                               const Row8Geom &q, const float2 *xt = nullptr) {
    const float *mu = means + (size_t)k * D;
    const float2 *mu2 = reinterpret_cast<const float2 *>(mu + q.lo_g + 2 * q.c);
    float2 mt[ROW8_TAIL];
    if (xt) {                                                        // the mean row's trailing elements, with its other loads
        const float2 *t2 = reinterpret_cast<const float2 *>(mu + q.lo_g + q.n8_g);
        const int n_t = q.n_g - q.n8_g;
#pragma unroll
        for (int t = 0; t < ROW8_TAIL; ++t) mt[t] = (2 * t < n_t) ? t2[t] : make_float2(0.f, 0.f);
    }
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXS; ++i) {
        if (i < q.steps) {
            const float2 mv = mu2[i * 4];
            const float d0 = __fsub_rn(mv.x, xv[i].x), d1 = __fsub_rn(mv.y, xv[i].y);
            const float p0 = __fmul_rn(d0, d0), p1 = __fmul_rn(d1, d1);
            a0 = (i == 0) ? p0 : __fadd_rn(a0, p0);
            a1 = (i == 0) ? p1 : __fadd_rn(a1, p1);
        }
    }
    float acc = __fadd_rn(a0, a1);                                   // r[2c] + r[2c+1]
    acc = __fadd_rn(acc, __shfl_xor_sync(q.gmask, acc, 1));
    acc = __fadd_rn(acc, __shfl_xor_sync(q.gmask, acc, 2));
    if (q.n8_g == 0) acc = 0.f;
    if (xt) {                                                        // the block's n % 8 trailing terms, in order
        const int n_t = q.n_g - q.n8_g;
#pragma unroll
        for (int t = 0; t < ROW8_TAIL; ++t)
            if (2 * t < n_t) {
                const float d0 = __fsub_rn(mt[t].x, xt[t].x), d1 = __fsub_rn(mt[t].y, xt[t].y);
                acc = __fadd_rn(acc, __fmul_rn(d0, d0));
                acc = __fadd_rn(acc, __fmul_rn(d1, d1));
            }
    } else {
        for (int d = q.lo_g + q.n8_g; d < q.lo_g + q.n_g; ++d) {
            const float dl = __fsub_rn(mu[d], xr[d]);
            acc = __fadd_rn(acc, __fmul_rn(dl, dl));
        }
    }
    float tot = __shfl_sync(q.gmask, acc, q.gbase);
    if (q.two_blocks) tot = __fadd_rn(tot, __shfl_sync(q.gmask, acc, q.gbase + 4));
    return -tot;
}

template <int MAXS>
__device__ __forceinline__ void km_exact_row8(const float *means, int KM, int D, const float *xr, const float2 *xv,
                                              int i1, int i2, uint32_t masks, int code, const Row8Geom &q, float &bv,
                                              int &bk, const float2 *xt = nullptr) {
    bv = -CUDART_INF_F;
    bk = 0x7fffffff;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        uint32_t mk = pass == 0 ? (masks & 0xffffu) : (code >= 0 ? (masks >> 16) : 0u);
        const int chunk = pass == 0 ? i1 : i2;
        while (mk) {
            const int bit = __ffs(mk) - 1;
            mk &= mk - 1;
            const int k = chunk * CHUNK + bit;
            if (k >= KM) continue;
            const float v = km_exact_one8<MAXS>(means, D, xr, xv, k, q, xt);
            if (v > bv || (v == bv && k < bk)) { bv = v; bk = k; }
        }
    }
}

template <int MAXS>
__device__ __forceinline__ void km_exact_row8(const float *means, int KM, int D, const float *xr, int i1, int i2,
                                              uint32_t masks, int code, const Row8Geom &q, float &bv, int &bk) {
    float2 xv[MAXS];
    km_load_x8<MAXS>(xr, q, xv);
    km_exact_row8<MAXS>(means, KM, D, xr, xv, i1, i2, masks, code, q, bv, bk);
}

}  // namespace mma
}  // namespace segb
