"""
ctypes binding of libsegb200.so (C ABI declared in include/segb200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is
present, the first call raises.  PyTorch is used only to own device memory and
streams; every pointer handed to the library is a raw `tensor.data_ptr()`.
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("SEGB_SO") or os.path.join(_HERE, "libsegb200.so")   # SEGB_SO: development builds
_LIB = None

c_i32, c_i64, c_f64, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p

DP_FFBS, DP_VITERBI_GMM, DP_VITERBI_KMEANS = 0, 1, 2
DP_SCORES_FINITE = 0x100        # OR-ed into the mode: scores hold no NaN / +inf (include/segb200.h)
DP_OK, DP_INFEASIBLE, DP_EMPTY_SLICE, DP_NAN = 0, 1, 2, 3
E_UNSUPPORTED = -2


class Corpus(ctypes.Structure):
    _fields_ = [("n_utt", c_i32), ("S", c_i32), ("N_max", c_i32), ("n_slices_min", c_i32),
                ("n_slices_max", c_i32), ("n_pos", c_i64), ("pos_off", c_vp), ("seg_id", c_vp),
                ("seg_dur", c_vp), ("bounds", c_vp), ("tok_id", c_vp)]


class FixedVar(ctypes.Structure):
    _fields_ = [("D", c_i32), ("K_max", c_i32), ("x_is_f64", c_i32), ("n_emb", c_i64), ("X", c_vp),
                ("mu_N_numT", c_vp), ("prec_NT", c_vp), ("prec_predT", c_vp), ("mu_NT", c_vp),
                ("log_prod_prec_pred", c_vp), ("counts", c_vp), ("assignments", c_vp), ("K", c_vp),
                ("n_total", c_vp), ("precision", c_vp), ("mu_0", c_vp), ("precision_0", c_vp),
                ("alpha", c_f64), ("lms", c_f64), ("sum_log_precision_0", c_f64),
                ("model", c_i32), ("v_0", c_i32), ("k_0", c_f64)]


class KMeansM(ctypes.Structure):
    _fields_ = [("D", c_i32), ("K_max", c_i32), ("x_is_f64", c_i32), ("n_emb", c_i64), ("X", c_vp),
                ("mean_num", c_vp), ("means", c_vp), ("meansT", c_vp), ("random_means", c_vp),
                ("counts", c_vp), ("assignments", c_vp), ("K", c_vp)]


class BigramLM(ctypes.Structure):
    _fields_ = [("K", c_i32), ("intrp_lambda", c_f64), ("a", c_f64), ("b", c_f64), ("unigram_counts", c_vp),
                ("bigram_counts", c_vp)]


_PROTOS = {
    "segb_last_error": (ctypes.c_char_p, []),
    "segb_version": (ctypes.c_int, []),
    "segb_launch_count": (c_i64, []),
    "segb_dp_banded": (ctypes.c_int, [ctypes.POINTER(Corpus), c_i32, c_i32, c_vp, c_i32, c_f64, c_f64, c_vp,
                                      c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "segb_fixedvar_add_items": (ctypes.c_int, [ctypes.POINTER(FixedVar), c_vp, c_vp, c_i32, c_vp]),
    "segb_fixedvar_del_items": (ctypes.c_int, [ctypes.POINTER(FixedVar), c_vp, c_i32, c_vp, c_i64, c_vp]),
    "segb_fixedvar_build": (ctypes.c_int, [ctypes.POINTER(FixedVar), c_vp, c_vp, c_i32, c_vp]),
    "segb_fixedvar_log_pred_row": (ctypes.c_int, [ctypes.POINTER(FixedVar), c_i32, c_vp, c_vp]),
    "segb_fixedvar_log_marg": (ctypes.c_int, [ctypes.POINTER(FixedVar), c_vp, c_vp, c_i64, c_f64, c_f64, c_vp, c_vp]),
    "segb_fixedvar_assign_items": (ctypes.c_int, [ctypes.POINTER(FixedVar), c_vp, c_i32, c_i32, c_f64, c_vp,
                                                  c_vp, c_vp, c_vp]),
    "segb_gibbs_sweep_fixedvar": (ctypes.c_int, [ctypes.POINTER(FixedVar), ctypes.POINTER(Corpus), c_vp, c_i32,
                                                 c_i32, c_f64, c_f64, c_f64, c_i32, c_vp, c_vp, c_vp, c_vp,
                                                 c_vp, c_vp]),
    "segb_gibbs_work_bytes": (c_i64, [c_i32, c_i32, c_i32]),
    "segb_gibbs_set_max_ctas": (ctypes.c_int, [c_i32]),
    "segb_gibbs_sweep_fixedvar_coop": (ctypes.c_int, [ctypes.POINTER(FixedVar), ctypes.POINTER(Corpus), c_vp, c_i32,
                                                      c_i32, c_f64, c_f64, c_f64, c_i32, c_vp, c_vp, c_vp, c_vp,
                                                      c_vp, c_vp]),
    "segb_fbgmm_gibbs_items_coop": (ctypes.c_int, [ctypes.POINTER(FixedVar), c_vp, c_i32, c_f64, c_vp, c_vp, c_vp,
                                                   c_vp]),
    "segb_kmeans_neg_sqrd_norm_row": (ctypes.c_int, [ctypes.POINTER(KMeansM), c_i32, c_vp, c_vp]),
    "segb_kmeans_best": (ctypes.c_int, [ctypes.POINTER(KMeansM), c_vp, c_i64, c_vp, c_vp, c_vp]),
    "segb_kmeans_add_items": (ctypes.c_int, [ctypes.POINTER(KMeansM), c_vp, c_vp, c_i32, c_vp]),
    "segb_kmeans_build": (ctypes.c_int, [ctypes.POINTER(KMeansM), c_vp, c_vp, c_i32, c_vp]),
    "segb_kmeans_del_items": (ctypes.c_int, [ctypes.POINTER(KMeansM), c_vp, c_i32, c_vp]),
    "segb_kmeans_move_items": (ctypes.c_int, [ctypes.POINTER(KMeansM), c_vp, c_vp, c_i32, c_vp]),
    "segb_kmeans_clean": (ctypes.c_int, [ctypes.POINTER(KMeansM), c_vp, c_i64, c_vp]),
    "segb_kmeans_band_scores": (ctypes.c_int, [ctypes.POINTER(KMeansM), ctypes.POINTER(Corpus), c_i64, c_i64,
                                               c_vp, c_f64, c_vp, c_vp]),
    "segb_kmeans_segment_sweep": (ctypes.c_int, [ctypes.POINTER(KMeansM), ctypes.POINTER(Corpus), c_vp, c_i32,
                                                 c_f64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "segb_kmeans_collect": (ctypes.c_int, [ctypes.POINTER(KMeansM), ctypes.POINTER(Corpus), c_i32, c_i32, c_vp,
                                           c_vp, c_vp, c_vp]),
    "segb_kmeans_set_means": (ctypes.c_int, [ctypes.POINTER(KMeansM), c_vp, c_vp, c_vp]),
    "segb_mma_x_tiles_bytes": (c_i64, [c_i64, c_i32]),
    "segb_mma_w_tiles_bytes": (c_i64, [c_i32, c_i32]),
    "segb_mma_cand_bytes": (c_i64, [c_i64]),
    "segb_mma_pack_x": (ctypes.c_int, [c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "segb_mma_pack_means": (ctypes.c_int, [c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "segb_mma_filter": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "segb_mma_refine_work_bytes": (c_i64, [c_i64, c_i32]),
    "segb_mma_refine": (ctypes.c_int, [ctypes.POINTER(KMeansM), c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp,
                                       c_vp]),
    "segb_mma_refine2_work_bytes": (c_i64, [c_i64, c_i32, c_i32]),
    "segb_mma_refine2": (ctypes.c_int, [ctypes.POINTER(KMeansM), c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_i32,
                                        c_vp, c_vp, c_vp, c_vp]),
    "segb_mma8_x_tiles_bytes": (c_i64, [c_i64, c_i32]),
    "segb_mma8_w_tiles_bytes": (c_i64, [c_i32, c_i32]),
    "segb_mma8_pack_x": (ctypes.c_int, [c_vp, c_i64, c_i32, ctypes.c_float, c_vp, c_vp, c_vp, c_vp]),
    "segb_mma8_pack_means": (ctypes.c_int, [c_vp, c_i32, c_i32, ctypes.c_float, c_vp, c_vp, c_vp, c_vp]),
    "segb_mma8_filter": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "segb_mma8_refine": (ctypes.c_int, [ctypes.POINTER(KMeansM), c_vp, c_vp, c_vp, ctypes.c_float, c_vp, c_vp, c_i64, c_vp,
                                        c_i64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "segb_fvmma_x_tiles_bytes": (c_i64, [c_i64, c_i32]),
    "segb_fvmma_w_tiles_bytes": (c_i64, [c_i32, c_i32]),
    "segb_fvmma_pack_x": (ctypes.c_int, [c_vp, c_i64, c_i32, c_vp, c_vp]),
    "segb_fvmma_log_marg": (ctypes.c_int, [ctypes.POINTER(FixedVar), c_vp, c_vp, c_i64, c_vp, c_vp]),
    "segb_fvf_x_tiles_bytes": (c_i64, [c_i64, c_i32, c_i32]),
    "segb_fvf_w_tiles_bytes": (c_i64, [c_i32, c_i32, c_i32]),
    "segb_fvf_model_bytes": (c_i64, [c_i32, c_i32, c_i32]),
    "segb_fvf_work_bytes": (c_i64, [c_i64]),
    "segb_fvf_pack_x": (ctypes.c_int, [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "segb_fvf_pack_model": (ctypes.c_int, [ctypes.POINTER(FixedVar), c_i32, c_vp, c_vp, c_vp, c_vp]),
    "segb_fvf_filter": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, ctypes.c_float, c_vp, c_vp]),
    "segb_fvf_refine": (ctypes.c_int, [c_vp, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, ctypes.c_float,
                                       c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "segb_fvf8_x_tiles_bytes": (c_i64, [c_i64, c_i32]),
    "segb_fvf8_w_tiles_bytes": (c_i64, [c_i32, c_i32]),
    "segb_fvf8_w_err_bytes": (c_i64, [c_i32]),
    "segb_fvf8_pack_x": (ctypes.c_int, [c_vp, c_i64, c_i32, ctypes.c_float, ctypes.c_float, c_vp, c_vp, c_vp, c_vp]),
    "segb_fvf8_pack_model": (ctypes.c_int, [c_i32, c_i32, c_vp, c_vp, ctypes.c_float, ctypes.c_float, c_vp, c_vp, c_vp, c_vp]),
    "segb_fvf8_filter": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, c_vp, c_vp]),
    "segb_fvf8_refine": (ctypes.c_int, [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, c_vp, c_vp, c_vp,
                                        c_vp, c_vp, c_vp]),
    "segb_fused_kmeans_best": (ctypes.c_int, [ctypes.POINTER(KMeansM), c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "segb_fused_fv_log_marg": (ctypes.c_int, [c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, ctypes.c_float, c_vp, c_vp,
                                              c_vp, c_vp, c_vp, c_vp]),
    "segb_fixedvar_band_scores": (ctypes.c_int, [ctypes.POINTER(Corpus), c_i64, c_i64, c_vp, c_f64, c_f64, c_vp, c_vp]),
    "segb_fvf_choose_tokens": (ctypes.c_int, [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, ctypes.POINTER(Corpus), c_i64,
                                              c_i64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "segb_tokens_from_bounds": (ctypes.c_int, [ctypes.POINTER(Corpus), c_i32, c_i32, c_vp]),
    "segb_frozen_new_work_bytes": (c_i64, [c_i64]),
    "segb_frozen_new_list": (ctypes.c_int, [ctypes.POINTER(Corpus), c_i64, c_i64, c_vp, c_i32, c_vp, c_i32, c_vp, c_vp,
                                            c_vp, c_vp]),
    "segb_frozen_clamp": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp,
                                         c_vp]),
    "segb_fixedvar_frozen_collect": (ctypes.c_int, [ctypes.POINTER(FixedVar), ctypes.POINTER(Corpus), c_i64, c_i64, c_vp,
                                                    c_vp, c_vp, c_vp]),
    "segb_fixedvar_frozen_update": (ctypes.c_int, [ctypes.POINTER(FixedVar), ctypes.POINTER(Corpus), c_i64, c_i64, c_vp,
                                                   c_vp, c_vp, c_vp, c_vp]),
    "segb_kmeans_frozen_clean_work_bytes": (c_i64, [c_i32, c_i32]),
    "segb_kmeans_frozen_clean": (ctypes.c_int, [ctypes.POINTER(KMeansM), ctypes.POINTER(Corpus), c_i64, c_i64, c_vp,
                                                c_vp, c_vp]),
    "segb_fixedvar_del_items_lm": (ctypes.c_int, [ctypes.POINTER(FixedVar), ctypes.POINTER(BigramLM), c_vp, c_i32, c_vp,
                                                  c_i64, c_vp]),
    "segb_bigram_lm_update": (ctypes.c_int, [ctypes.POINTER(BigramLM), c_vp, c_i32, c_i32, c_vp]),
    "segb_bigram_lm_log_prob_row": (ctypes.c_int, [ctypes.POINTER(BigramLM), c_i32, c_vp, c_vp]),
    "segb_gibbs_sweep_bigram": (ctypes.c_int, [ctypes.POINTER(FixedVar), ctypes.POINTER(BigramLM), ctypes.POINTER(Corpus),
                                               c_vp, c_i32, c_i32, c_f64, c_f64, c_f64, c_i32, c_vp, c_vp, c_vp, c_vp,
                                               c_vp, c_vp]),
    "segb_host_init_boundaries": (c_i64, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_f64, c_i64, c_i64, c_vp]),
    "segb_debug_gibbs_prof": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "segb_debug_gibbs_bar_base": (ctypes.c_int, [ctypes.c_uint32]),
    "segb_gibbs_sweep_bigram_coop": (ctypes.c_int, [ctypes.POINTER(FixedVar), ctypes.POINTER(BigramLM),
                                                    ctypes.POINTER(Corpus), c_vp, c_i32, c_f64, c_f64, c_f64, c_i32, c_vp,
                                                    c_vp, c_vp, c_vp, c_vp, c_vp]),
    "segb_fixedvar_log_marg_k_work_bytes": (c_i64, [c_i32, c_i32]),
    "segb_fixedvar_log_marg_k": (ctypes.c_int, [ctypes.POINTER(FixedVar), c_vp, c_vp, c_vp, c_vp, c_vp]),
    "segb_kmeans_sum_neg_sqrd_norm_k": (ctypes.c_int, [ctypes.POINTER(KMeansM), c_vp, c_vp, c_vp, c_vp]),
    "segb_diag_log_marg_k": (ctypes.c_int, [ctypes.POINTER(FixedVar), c_vp, c_vp]),
}

EXPORTS = sorted(_PROTOS)


def load():
    """Load libsegb200.so and attach prototypes.  Raises if it was not built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                "libsegb200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `bash segmentalist_b200/csrc/build.sh`; there is no CPU fallback." % SO_PATH)
        lib = ctypes.CDLL(SO_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)       # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def lib():
    """The library, for compute calls: additionally requires a CUDA device."""
    if not torch.cuda.is_available():
        raise RuntimeError("segmentalist_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return load()


def check(rc):
    if rc != 0:
        msg = load().segb_last_error().decode("utf-8", "replace")
        if rc < 0:
            raise AssertionError("segb200: %s (code %d)" % (msg, rc))
        raise RuntimeError("segb200: %s (cuda error %d)" % (msg, rc))


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def members_by_component(assign, K_max):
    """Items grouped by component for the diagnostics kernels: (order, seg_off) with `order` the item ids
    stably sorted by assignment and seg_off[k] the start of component k's members (torch sort = plumbing)."""
    order = torch.argsort(assign, stable=True)
    seg_off = torch.searchsorted(assign[order].contiguous(),
                                 torch.arange(K_max + 1, dtype=assign.dtype, device=assign.device))
    return order.contiguous(), seg_off.to(torch.int64).contiguous()


def dev(a, dtype=None):
    """NumPy -> contiguous CUDA tensor."""
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()
