"""
Host-side logic of the sharded frozen-state sweep (no CUDA needed: the same functions run on
CPU tensors under the `gloo` backend in tests/test_dist_gloo.py and on CUDA tensors under NCCL).

The update of a frozen sweep is a pure reduction: per-component sum of the embeddings of the
chosen segments and per-component counts.  Ranks own contiguous utterance ranges; one all-reduce
of (sum_x, cnt) makes every rank rebuild identical means, and `clean_components`
(kmeans_components.py:263-266) is replayed identically everywhere from the global counts.
"""
import numpy as np
import torch
import torch.distributed as dist


_LOCAL_ONLY = 0


class local_only(object):
    """Context manager: inside it the sweeps behave as a single rank even when a process group is
    initialised (used to compare an N-rank sweep with the 1-rank sweep of the same corpus)."""

    def __enter__(self):
        global _LOCAL_ONLY
        _LOCAL_ONLY += 1

    def __exit__(self, *exc):
        global _LOCAL_ONLY
        _LOCAL_ONLY -= 1
        return False


def dist_on():
    return (not _LOCAL_ONLY) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_ranges(n_utt, world):
    """Contiguous utterance range [lo, hi) of every rank; sizes differ by at most one."""
    base, rem = divmod(int(n_utt), int(world))
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def reduce_stats(sum_x, cnt, extra=None):
    """All-reduce (SUM) the sufficient statistics in place; `extra` is an optional float64
    tensor of scalars (e.g. the sweep objective) reduced with them."""
    if dist_on():
        dist.all_reduce(sum_x, op=dist.ReduceOp.SUM)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        if extra is not None:
            dist.all_reduce(extra, op=dist.ReduceOp.SUM)
    return sum_x, cnt, extra


def reduce_packed(red, cnt_f, cnt):
    """The sweep's ONE collective: `red` is a flat float64 buffer [sum_x | counts]; the int64
    counts are copied into its tail `cnt_f` (exact below 2^53), the whole buffer is all-reduced
    (SUM) and the counts copied back.  No-op without an initialised process group."""
    if dist_on():
        cnt_f.copy_(cnt)
        dist.all_reduce(red, op=dist.ReduceOp.SUM)
        cnt.copy_(cnt_f)
    return red, cnt


def compaction_plan(cnt, K):
    """Replay clean_components' swap-with-last deletions (kmeans_components.py:149-166,263-266)
    on the counts of the K active slots.  Returns (K_new, dst, src): after the deletions slot
    dst[i] holds what was in slot src[i]; slots [K_new, K) become inactive."""
    cnt = np.asarray(cnt)[:K]
    slot_src = np.arange(K)
    K_new = K
    for k in np.where(cnt == 0)[0][::-1]:
        K_new -= 1
        if k != K_new:
            slot_src[k] = slot_src[K_new]
    dst = np.where(slot_src[:K_new] != np.arange(K_new))[0]
    return int(K_new), dst, slot_src[dst]


def clamp_plan(ks, K_before):
    """add_item's `k > K -> K; k == K -> K += 1` rule (kmeans_components.py:103-106) applied to
    the component choices `ks` of tokens taken in utterance order.  Returns (new ks, K)."""
    ks = np.array(ks, dtype=np.int64)
    K = int(K_before)
    for i in range(len(ks)):
        k = ks[i]
        if k > K:
            k = K
        if k == K:
            K += 1
        ks[i] = k
    return ks, K


def means_from_stats(sum_x, cnt, old_means):
    """means[k] = (dtype)(sum_x[k] / cnt[k]) where cnt[k] > 0 (kmeans_components.py:110);
    emptied components keep their stale mean until cleaned (:131-132).  NumPy version of
    segb_kmeans_set_means, used by the CPU tests."""
    means = np.array(old_means, copy=True)
    live = np.asarray(cnt) > 0
    means[live] = (np.asarray(sum_x)[live] / np.asarray(cnt)[live, None]).astype(means.dtype)
    return means
