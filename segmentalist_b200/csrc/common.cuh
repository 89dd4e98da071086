// Shared helpers for libsegb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/segb200.h"

namespace segb {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
// current device id, its SM count and whether it supports cooperative launches (cached per device)
int device_info(int *dev_out, int *n_sm_out, int *coop_out);

#define SEGB_CHECK_ARG(cond, msg)                         \
    do {                                                  \
        if (!(cond)) {                                    \
            segb::set_error("bad argument: %s", msg);     \
            return SEGB_E_ARG;                            \
        }                                                 \
    } while (0)

#define SEGB_CUDA(expr)                                                            \
    do {                                                                           \
        cudaError_t e__ = (expr);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            segb::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));      \
            return (int)e__;                                                       \
        }                                                                          \
    } while (0)

#define SEGB_LAUNCH_CHECK()                                                        \
    do {                                                                           \
        segb::count_launch();                                                      \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) {                                                  \
            segb::set_error("kernel launch failed: %s", cudaGetErrorString(e__));  \
            return (int)e__;                                                       \
        }                                                                          \
    } while (0)

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double neg_inf() { return -CUDART_INF; }

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Block-wide reductions over blockDim.x threads (multiple of 32, <= 1024).
// `red` is shared scratch of >= 33 doubles.  Result broadcast to every thread.
__device__ __forceinline__ double block_max(double v, double *red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double x = lane < nw ? red[lane] : neg_inf();
        x = warp_max(x);
        if (lane == 0) red[32] = x;
    }
    __syncthreads();
    return red[32];
}
__device__ __forceinline__ double block_sum(double v, double *red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double x = lane < nw ? red[lane] : 0.0;
        x = warp_sum(x);
        if (lane == 0) red[32] = x;
    }
    __syncthreads();
    return red[32];
}

// NumPy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src,
// @TYPE@_pairwise_sum) applied to n values produced on the fly by `f(i)`.
// Separate rounded add per operation (no FMA contraction) so the result is
// bit-identical to `arr.sum(axis=-1)` on a contiguous row.
template <typename T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <typename T> __device__ __forceinline__ T sub_rn(T a, T b);
template <> __device__ __forceinline__ float sub_rn<float>(float a, float b) { return __fsub_rn(a, b); }
template <> __device__ __forceinline__ double sub_rn<double>(double a, double b) { return __dsub_rn(a, b); }
template <typename T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }

template <typename T, typename F>
__device__ T pairwise_block(F f, int lo, int n) {   // n <= 128
    if (n < 8) {
        T res = T(0);
        for (int i = 0; i < n; ++i) res = add_rn<T>(res, f(lo + i));
        return res;
    }
    T r0 = f(lo), r1 = f(lo + 1), r2 = f(lo + 2), r3 = f(lo + 3);
    T r4 = f(lo + 4), r5 = f(lo + 5), r6 = f(lo + 6), r7 = f(lo + 7);
    int i = 8;
    const int n8 = n - (n % 8);
    for (; i < n8; i += 8) {
        r0 = add_rn<T>(r0, f(lo + i));     r1 = add_rn<T>(r1, f(lo + i + 1));
        r2 = add_rn<T>(r2, f(lo + i + 2)); r3 = add_rn<T>(r3, f(lo + i + 3));
        r4 = add_rn<T>(r4, f(lo + i + 4)); r5 = add_rn<T>(r5, f(lo + i + 5));
        r6 = add_rn<T>(r6, f(lo + i + 6)); r7 = add_rn<T>(r7, f(lo + i + 7));
    }
    T res = add_rn<T>(add_rn<T>(add_rn<T>(r0, r1), add_rn<T>(r2, r3)),
                      add_rn<T>(add_rn<T>(r4, r5), add_rn<T>(r6, r7)));
    for (; i < n; ++i) res = add_rn<T>(res, f(lo + i));
    return res;
}

// Iterative form of the recursion (n > 128 splits at n/2 rounded down to a
// multiple of 8): an explicit stack of (lo, n, state) avoids device recursion.
template <typename T, typename F>
__device__ T pairwise_sum(F f, int n) {
    if (n <= 128) return pairwise_block<T>(f, 0, n);
    // depth <= 24 is enough for n < 2^31
    int lo_s[24], n_s[24];
    T acc_s[24];
    unsigned char st_s[24];   // 0 = fresh, 1 = left done
    int sp = 0;
    lo_s[0] = 0; n_s[0] = n; st_s[0] = 0;
    T ret = T(0);
    while (sp >= 0) {
        const int lo = lo_s[sp], nn = n_s[sp];
        if (nn <= 128) { ret = pairwise_block<T>(f, lo, nn); --sp; continue; }
        int n2 = nn / 2; n2 -= n2 % 8;
        if (st_s[sp] == 0) {            // descend left
            st_s[sp] = 1;
            ++sp; lo_s[sp] = lo; n_s[sp] = n2; st_s[sp] = 0;
        } else if (st_s[sp] == 1) {     // left returned in `ret`; descend right
            acc_s[sp] = ret; st_s[sp] = 2;
            ++sp; lo_s[sp] = lo + n2; n_s[sp] = nn - n2; st_s[sp] = 0;
        } else {                        // right returned
            ret = add_rn<T>(acc_s[sp], ret); --sp;
        }
    }
    return ret;
}

// The same pairwise sum (n <= 256 terms) computed by 16 consecutive lanes: NumPy's block sum
// keeps 8 running accumulators per block of <= 128 terms, so lane q of an 8-lane group owns
// accumulator q, the xor-butterfly (1, 2, 4) reproduces ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) on the
// group's first lane, which then adds the n % 8 leftovers one by one; the second 8-lane group
// takes the second block when n > 128.  Bit-identical to pairwise_sum<T>(f, n).  `hmask` = the
// 16 participating lanes of the warp, `j` = lane index within them; result on all 16 lanes.
// pairwise_sum for the sizes that occur per embedding (n <= 256): at most ONE split, so no recursion
// stack (the general routine keeps its stack in local memory).  Bit-identical to pairwise_sum<T>(f, n).
template <typename T, typename F>
__device__ __forceinline__ T pairwise_sum_le256(F f, int n) {
    if (n <= 128) return pairwise_block<T>(f, 0, n);
    if (n <= 256) {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return add_rn<T>(pairwise_block<T>(f, 0, n2), pairwise_block<T>(f, n2, n - n2));
    }
    return pairwise_sum<T>(f, n);
}
// The same without the general fallback: for kernels whose launcher guarantees n <= 256 (keeps the
// recursion stack of pairwise_sum out of their frames).
template <typename T, typename F>
__device__ __forceinline__ T pairwise_sum_max256(F f, int n) {
    if (n <= 128) return pairwise_block<T>(f, 0, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return add_rn<T>(pairwise_block<T>(f, 0, n2), pairwise_block<T>(f, n2, n - n2));
}
template <typename T, typename F>
__device__ __forceinline__ T pairwise_sum_lanes16(F f, int n, unsigned hmask, int j) {
    const int g = j >> 3, q = j & 7;
    const int hbase = (threadIdx.x & 31) & 16;
    int n2 = 0;
    if (n > 128) { n2 = n / 2; n2 -= n2 % 8; }
    const int lo = g == 0 ? 0 : n2;
    const int ng = (n > 128) ? (g == 0 ? n2 : n - n2) : (g == 0 ? n : 0);
    T res = T(0);
    if (ng >= 8) {
        const int n8 = ng - (ng % 8);
        T acc = f(lo + q);
        for (int i = 8; i < n8; i += 8) acc = add_rn<T>(acc, f(lo + i + q));
        acc = add_rn<T>(acc, __shfl_xor_sync(hmask, acc, 1));
        acc = add_rn<T>(acc, __shfl_xor_sync(hmask, acc, 2));
        acc = add_rn<T>(acc, __shfl_xor_sync(hmask, acc, 4));
        res = acc;
        for (int i = n8; i < ng; ++i) res = add_rn<T>(res, f(lo + i));
    } else {
        // keep the shuffles convergent for the whole half-warp
        T acc = T(0);
        acc = add_rn<T>(acc, __shfl_xor_sync(hmask, acc, 1));
        acc = add_rn<T>(acc, __shfl_xor_sync(hmask, acc, 2));
        acc = add_rn<T>(acc, __shfl_xor_sync(hmask, acc, 4));
        for (int i = 0; i < ng; ++i) res = add_rn<T>(res, f(lo + i));
    }
    T tot = __shfl_sync(hmask, res, hbase);
    if (n > 128) tot = add_rn<T>(tot, __shfl_sync(hmask, res, hbase + 8));
    return tot;
}

}  // namespace segb
