// Pieces of the FBGMM log_marg_i refine shared by fixedvar_filter.cu (standalone refine) and
// score_fused.cu (refine fused behind the filter GEMM).
#pragma once
#include "mma_common.cuh"

namespace segb {
namespace fvf {

using namespace segb::mma;

// padded inner dimension of ONE chunk of the FBGMM filter operands, and the number of chunks
__host__ __device__ inline int kp_of(int D, int aniso) { return ((aniso ? D + 3 : D + 6) + 15) / 16 * 16; }
__host__ __device__ inline int nch_of(int aniso) { return aniso ? 2 : 1; }

// Exact per-component tables read by the refine (row-major so one component is one contiguous row) and,
// transposed (component index fastest), by the exhaustive scan:
//   mu [Kr][D] | P [Kr][D] (anisotropic only) | muT [D][Kr] | PT [D][Kr] (anisotropic only)
//   | cst_lse [Kr] | cst_map [Kr] | pk [Kr]                                              Kr = K_max + 1
// cst_lse = lms*pi_k - D/2 log 2pi + 1/2 sum_d log P_kd (+ log(K_max - K) on the virtual row);
// cst_map = log(alpha/K_max + n_k) - D/2 log 2pi + 1/2 sum_d log P_kd   (map_assign_i: no lms, fbgmm.py:475-479).
struct ModelRows {
    const double *mu, *P, *muT, *PT, *cst_lse, *cst_map, *pk;
};
__host__ __device__ inline int64_t model_doubles(int K_max, int D, int aniso) {
    const int64_t Kr = K_max + 1;
    return 2 * Kr * D * (aniso ? 2 : 1) + 3 * Kr;
}
__host__ __device__ inline ModelRows model_view(const double *base, int K_max, int D, int aniso) {
    const int64_t Kr = K_max + 1, n = Kr * D;
    ModelRows r;
    r.mu = base;
    r.P = aniso ? base + n : nullptr;
    r.muT = base + n * (aniso ? 2 : 1);
    r.PT = aniso ? r.muT + n : nullptr;
    const double *q = base + 2 * n * (aniso ? 2 : 1);
    r.cst_lse = q; r.cst_map = q + Kr; r.pk = q + 2 * Kr;
    return r;
}

// Exact float64 quadratic form of component row k for the embedding whose elements d = j, j+8, ... this
// lane holds: sum_d P_kd (mu_kd - x_d)^2 over the lane's elements (isotropic: the factor p_k is applied
// by the caller).  Eight lanes per embedding; the caller reduces over them.
template <bool ANISO, int DC = 0>
__device__ __forceinline__ double quad_part(const ModelRows &t, int k, const float *xr, int D, int j) {
    const double *mu = t.mu + (size_t)k * D;
    double acc = 0.0;
    if (DC > 0 && (DC & 1) == 0) {
        // dimension known at compile time: the same two chains, fully unrolled (all loads of the row in flight at once)
        double acc1 = 0.0;
        const double *P = ANISO ? t.P + (size_t)k * DC : nullptr;
        const float *xj = xr + 2 * j;
        const double *mj = t.mu + (size_t)k * DC + 2 * j;
#pragma unroll
        for (int i = 0; i < (DC + 15) / 16; ++i) {
            if (16 * i + 14 < DC || 2 * j + 16 * i < DC) {
                const float2 xv = *reinterpret_cast<const float2 *>(xj + 16 * i);
                const double2 mv = *reinterpret_cast<const double2 *>(mj + 16 * i);
                const double d0 = mv.x - (double)xv.x, d1 = mv.y - (double)xv.y;
                if (ANISO) {
                    const double2 pv = *reinterpret_cast<const double2 *>(P + 2 * j + 16 * i);
                    acc = fma(d0 * d0, pv.x, acc); acc1 = fma(d1 * d1, pv.y, acc1);
                } else { acc = fma(d0, d0, acc); acc1 = fma(d1, d1, acc1); }
            }
        }
        return acc + acc1;
    }
    if ((D & 1) == 0) {
        // even D: rows are 8-byte (x) / 16-byte (tables) aligned -- two elements per load, two chains
        double acc1 = 0.0;
        const double *P = ANISO ? t.P + (size_t)k * D : nullptr;
#pragma unroll 3
        for (int d = 2 * j; d < D; d += 16) {
            const float2 xv = *reinterpret_cast<const float2 *>(xr + d);
            const double2 mv = *reinterpret_cast<const double2 *>(mu + d);
            const double d0 = mv.x - (double)xv.x, d1 = mv.y - (double)xv.y;
            if (ANISO) {
                const double2 pv = *reinterpret_cast<const double2 *>(P + d);
                acc = fma(d0 * d0, pv.x, acc); acc1 = fma(d1 * d1, pv.y, acc1);
            } else { acc = fma(d0, d0, acc); acc1 = fma(d1, d1, acc1); }
        }
        return acc + acc1;
    }
    if (ANISO) {
        const double *P = t.P + (size_t)k * D;
#pragma unroll 4
        for (int d = j; d < D; d += 8) { const double dl = mu[d] - (double)xr[d]; acc = fma(dl * dl, P[d], acc); }
    } else {
#pragma unroll 4
        for (int d = j; d < D; d += 8) { const double dl = mu[d] - (double)xr[d]; acc = fma(dl, dl, acc); }
    }
    return acc;
}

// Running logsumexp + MAP argmax over the exact scores fed one at a time.
struct LseAcc {
    double m, s, best;
    int bk;
    __device__ __forceinline__ void init() { m = -CUDART_INF; s = 0.0; best = -CUDART_INF; bk = 0x7fffffff; }
    __device__ __forceinline__ void add(double v, double vmap, int k) {
        if (m == -CUDART_INF) { m = v; s = 1.0; }                  // first score: no exponential
        else if (v > m) { s = s * exp(m - v) + 1.0; m = v; }
        else s += exp(v - m);
        if (vmap > best || (vmap == best && k < bk)) { best = vmap; bk = k; }
    }
    __device__ __forceinline__ double lse() const { return s == 1.0 ? m : m + log(s); }
};


// The exact logsumexp (and MAP slot) of one embedding over the candidates the filter kept, by EIGHT lanes.
// code (refine_decide): -1 the best chunk i1 suffices, >= 0 also visit chunk i2; masks as in Cand.
template <bool ANISO, int DC = 0>
__device__ __forceinline__ LseAcc fv_exact_row8(const ModelRows &t, int Kr, int D, const float *xr, int i1, int i2,
                                                uint32_t masks, int code, int j, unsigned gmask) {
    LseAcc acc;
    acc.init();
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        // best chunk: the members within the threshold of the chunk maximum; second chunk: everything
        uint32_t mk = pass == 0 ? (masks & 0xffffu) : (code >= 0 ? (masks >> 16) : 0u);
        const int chunk = pass == 0 ? i1 : i2;
        while (mk) {
            const int bit = __ffs(mk) - 1;
            mk &= mk - 1;
            const int k = chunk * CHUNK + bit;
            if (k >= Kr) continue;
            double q = quad_part<ANISO, DC>(t, k, xr, D, j);
            q += __shfl_xor_sync(gmask, q, 1);
            q += __shfl_xor_sync(gmask, q, 2);
            q += __shfl_xor_sync(gmask, q, 4);
            const double pred = -0.5 * (ANISO ? q : t.pk[k] * q);
            const double c_lse = t.cst_lse[k];
            if (c_lse == -CUDART_INF) continue;                 // dead row
            acc.add(c_lse + pred, t.cst_map[k] + pred, k);
        }
    }
    return acc;
}

}  // namespace fvf
}  // namespace segb
