#!/usr/bin/env python
"""
bench.py -- headline benchmark of the segmentalist hot path on B200.

Workload (BASELINE.json configs[2], the configuration the metric "utterances/sec
per sweep (1/2/4/8 B200)" is quoted on; it fits one GPU): frozen-state k-means
Viterbi segmentation sweep, synthetic D=130 unit-norm float32 embeddings,
K=5000 components, 200k utterances (N ~ U{15..25} landmarks, max_span 6),
sharded by utterance over the ranks (strong scaling: the total is fixed).

One "step" = one sweep: pack means -> tcgen05 filter GEMM -> exact refine ->
banded scores -> Viterbi DP -> token collection -> NCCL all-reduce of
(sum_x, counts) -> means update.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nproc-per-node 8 ... bench.py --gpus 8 ...
    python bench.py --impl reference        # CPU arm: oracle port on the host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

D, K_MAX, S_MAX = 130, 5000, 6
N_LO, N_HI = 15, 25
TOTAL_UTTS = 200000
NOISE = 0.05
# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the three reported kernels,
# from ONE `ncu --set full --clock-control none` capture of this command at the default configuration on
# one GPU (profiles/r1_ncu_summary_v4.md).  Reported only when the run uses that configuration.
NCU_TRAFFIC_BYTES = {"filter": 6.054860e9 + 0.668228e9, "dp": 0.193679e9 + 0.008181e9,
                     "fv_logmarg_per_row": (607.070976e6 + 4.2e6) / 1048576}
METRIC = "utterances/sec per sweep"
WORKLOAD = "kmeans_viterbi_frozen_sweep D=130 K=5000 U=200k max_span=6 (BASELINE configs[2])"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utts", type=int, default=TOTAL_UTTS, help="total utterances over all ranks")
    ap.add_argument("--K", type=int, default=K_MAX)
    ap.add_argument("--cpu-sample", type=int, default=32, help="utterances timed by the CPU baseline leg")
    ap.add_argument("--scorer", default="mma", choices=["mma", "exact"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp16", "fp8"],
                    help="first-level filter GEMM of the k-means scorer: fp16 (kind::f16), e4m3 (kind::f8f6f4) with the fp16 pass as "
                         "second level, or auto (e4m3 until a sweep leaves > 35 %% of the rows undecided)")
    ap.add_argument("--fused", action="store_true", help="the fused score kernel (fp32 rows in, conversion + filter GEMM + refine in one launch) instead of pre-packed fp16 image + filter kernel + refine kernel")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gibbs", action="store_true", help="skip the sequential secondary workloads (BASELINE configs[1], [3], [4])")
    ap.add_argument("--no-fbgmm", action="store_true", help="skip the sharded frozen FBGMM sweep")
    ap.add_argument("--no-ingest", action="store_true", help="skip the set-up timing through the public constructor")
    ap.add_argument("--ingest-utts", type=int, default=TOTAL_UTTS)
    ap.add_argument("--no-diffuse", action="store_true", help="skip the diffuse-model k-means sweep (K_act < K_max)")
    ap.add_argument("--fbgmm-utts", type=int, default=0, help="utterances of the frozen FBGMM sweep (default: --utts)")
    ap.add_argument("--diffuse-utts", type=int, default=40000)
    ap.add_argument("--only-diffuse", action="store_true", help="run only the diffuse-model k-means secondary (one GPU)")
    ap.add_argument("--gibbs-utts", type=int, default=2000)
    ap.add_argument("--only-gibbs", action="store_true", help="run only the secondary Gibbs workload")
    ap.add_argument("--diag-utts", type=int, default=800)
    ap.add_argument("--diag-k-true", type=int, default=5000)
    ap.add_argument("--diag-init", default="one-by-one", choices=["rand", "one-by-one"])
    ap.add_argument("--only-diag", action="store_true", help="run only the diagonal-covariance secondary workload")
    ap.add_argument("--bigram-utts", type=int, default=8000)
    ap.add_argument("--only-bigram", action="store_true", help="run only the bigram cluster-sampling secondary workload")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic corpus structure (vectorised)
# ------------------------------------------------------------------------------------------------

def corpus_structure(n_utt, seed):
    """lengths, banded seg_id [n_pos, S] (local embedding ids), seg_dur [n_pos, S], initial bounds."""
    rng = np.random.RandomState(seed)
    S = S_MAX
    lengths = rng.randint(N_LO, N_HI + 1, size=n_utt).astype(np.int64)
    pos_off = np.concatenate([[0], np.cumsum(lengths)])
    n_pos = int(pos_off[-1])
    n_seg_of = {N: sum(min(S, N - s) for s in range(N)) for N in range(N_LO, N_HI + 1)}
    emb_off = np.concatenate([[0], np.cumsum([n_seg_of[int(N)] for N in lengths])]).astype(np.int64)
    seg_id = np.full((n_pos, S), -1, dtype=np.int32)
    seg_dur = np.full((n_pos, S), np.nan)
    for N in range(N_LO, N_HI + 1):
        idx = np.where(lengths == N)[0]
        if len(idx) == 0:
            continue
        first = np.concatenate([[0], np.cumsum([min(S, N - s) for s in range(N)])])
        tmpl = np.full((N, S), -1, dtype=np.int64)
        for t in range(1, N + 1):
            for l in range(1, min(t, S) + 1):
                tmpl[t - 1, l - 1] = first[t - l] + (l - 1)          # start-major row order
        live = tmpl >= 0
        gaps = rng.randint(3, 15, size=(len(idx), N))
        B = np.concatenate([np.zeros((len(idx), 1), dtype=np.int64), np.cumsum(gaps, axis=1)], axis=1)
        dur = np.full((len(idx), N, S), np.nan)
        for l in range(1, S + 1):
            dur[:, l - 1:, l - 1] = B[:, l:] - B[:, :N + 1 - l]
        rows = pos_off[idx][:, None] + np.arange(N)[None, :]
        ids = np.where(live[None], tmpl[None] + emb_off[idx][:, None, None], -1)
        seg_id[rows] = ids.astype(np.int32)
        seg_dur[rows] = dur
    # initial boundaries: coin flips, a forced boundary every S landmarks and at the end
    within = np.arange(n_pos) - np.repeat(pos_off[:-1], lengths)
    b = (rng.rand(n_pos) < 0.5) | ((within + 1) % S == 0)
    b[pos_off[1:] - 1] = True
    return lengths, seg_id, seg_dur, b.astype(np.uint8), int(emb_off[-1])


def make_embeddings_gpu(n_emb, centres, seed, device):
    """x = normalise(c_z + NOISE * N(0, I)), z uniform over the rows of `centres` (device tensor)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    K_true = centres.shape[0]
    X = torch.empty(n_emb, D, dtype=torch.float32, device=device)
    Z = torch.empty(n_emb, dtype=torch.int32, device=device)
    step = 1 << 21
    for lo in range(0, n_emb, step):
        hi = min(n_emb, lo + step)
        z = torch.randint(0, K_true, (hi - lo,), generator=g, device=device)
        x = centres[z] + NOISE * torch.randn(hi - lo, D, generator=g, device=device)
        X[lo:hi] = x / x.norm(dim=1, keepdim=True)
        Z[lo:hi] = z.to(torch.int32)
    return X, Z


CPU_SAMPLE_UTTS = 64          # head of rank 0's shard: generated on the CPU so that both arms see the same utterances


def shard_structure(n_utt, rank):
    """A rank's shard.  Rank 0's first CPU_SAMPLE_UTTS utterances are the fixed "head" corpus
    (corpus_structure(CPU_SAMPLE_UTTS, 999)) whose embeddings come from sample_rows_cpu: the CPU arms
    (cpu_baseline, --impl reference) rebuild exactly these utterances without a GPU."""
    if rank != 0 or n_utt <= CPU_SAMPLE_UTTS:
        return corpus_structure(n_utt, seed=1000 + rank) + (0,)
    hl, hi_, hd, hb, hn = corpus_structure(CPU_SAMPLE_UTTS, seed=999)
    rl, ri, rd, rb, rn = corpus_structure(n_utt - CPU_SAMPLE_UTTS, seed=1000)
    seg_id = np.concatenate([hi_, np.where(ri >= 0, ri + hn, -1).astype(np.int32)])
    return (np.concatenate([hl, rl]), seg_id, np.concatenate([hd, rd]), np.concatenate([hb, rb]), hn + rn, hn)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows, self.proc = [], None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_mhz_min": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own pure functions (baseline/_ref, a py3-shimmed copy built by
# __graft_entry__.build()) on host cores; the oracle port where that directory is absent
# ------------------------------------------------------------------------------------------------

def centres_cpu(K):
    """Cluster centres shared by all ranks and by both arms (NumPy, fixed seed)."""
    rng = np.random.RandomState(12345)
    c = rng.standard_normal((K, D)).astype(np.float32)
    return (c / np.linalg.norm(c, axis=1, keepdims=True)).astype(np.float32)


def sample_rows_cpu(n_rows, centres, seed):
    """(X, z) of the first n_rows embeddings of a rank's shard, CPU-reproducible."""
    rng = np.random.RandomState(seed)
    z = rng.randint(0, centres.shape[0], n_rows)
    X = centres[z] + NOISE * rng.standard_normal((n_rows, D)).astype(np.float32)
    return (X / np.linalg.norm(X, axis=1, keepdims=True)).astype(np.float32), z.astype(np.int32)


def cpu_modules():
    """("reference", namespace over the reference's own modules) when baseline/_ref is present, else
    ("port", the oracle module): both expose Utterances, KMeansComponents, KMeans, SegmentalKMeansWordseg,
    forward_backward_kmeans_viterbi."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref, "segmentalist")):
        try:
            if ref not in sys.path:
                sys.path.insert(0, ref)
            import types
            from segmentalist import kmeans, kmeans_acoustic_wordseg, kmeans_components, utterances
            ns = types.SimpleNamespace(
                Utterances=utterances.Utterances, KMeansComponents=kmeans_components.KMeansComponents,
                KMeans=kmeans.KMeans, SegmentalKMeansWordseg=kmeans_acoustic_wordseg.SegmentalKMeansWordseg,
                forward_backward_kmeans_viterbi=kmeans_acoustic_wordseg.forward_backward_kmeans_viterbi)
            return "reference", ns
        except Exception as exc:          # e.g. the Cython extension does not load on this box
            sys.stderr.write("bench: baseline/_ref unusable (%r); timing the oracle port\n" % (exc,))
    from oracle import seg_oracle as so
    return "port", so


def _cpu_segmenter(ns, X_sub, lengths, seg_id_band, seg_dur_band, means):
    """Wrap flat arrays into SegmentalKMeansWordseg / KMeansComponents objects of `ns` (the reference's
    classes or the oracle's): attributes set directly, no constructor (the model is given)."""
    from segmentalist_b200.utterances import band_to_packed
    S = seg_id_band.shape[1]
    utts = ns.Utterances.__new__(ns.Utterances)
    utts.lengths = [int(n) for n in lengths]
    utts.D = len(lengths)
    utts.N_max = int(max(lengths))
    width = utts.N_max * (utts.N_max + 1) // 2
    utts.vec_ids = np.full((utts.D, width), -1, dtype=np.int64)
    utts.durations = np.full((utts.D, width), np.nan)
    utts.boundaries = np.zeros((utts.D, utts.N_max), dtype=bool)
    pos = 0
    for u, N in enumerate(utts.lengths):
        n_packed = N * (N + 1) // 2
        utts.vec_ids[u, :n_packed] = band_to_packed(seg_id_band[pos:pos + N].astype(np.int64), N, S, -1)
        utts.durations[u, :n_packed] = band_to_packed(seg_dur_band[pos:pos + N], N, S, np.nan)
        utts.boundaries[u, N - 1] = True
        pos += N
    comps = ns.KMeansComponents.__new__(ns.KMeansComponents)
    comps.X, comps.means = X_sub, means
    comps.N, comps.D = X_sub.shape
    comps.K_max = comps.K = means.shape[0]
    km = ns.KMeans.__new__(ns.KMeans)
    km.components = comps
    seg = ns.SegmentalKMeansWordseg.__new__(ns.SegmentalKMeansWordseg)
    seg.utterances, seg.acoustic_model = utts, km
    seg.n_slices_min, seg.n_slices_max, seg.wip = 0, S, 0
    return seg


def _cpu_phase1(ns, seg):
    """The pure part of a frozen sweep through the CPU implementation's own functions, per utterance:
    get_vec_embed_neg_len_sqrd_norms (kmeans_acoustic_wordseg.py:334-351) -> forward_backward_kmeans_viterbi
    (:449-555) -> get_max_assignments (kmeans_components.py:256-261)."""
    utts, comps = seg.utterances, seg.acoustic_model.components
    totals, bounds, ks = [], [], []
    for u in range(utts.D):
        N = utts.lengths[u]
        n_packed = (N ** 2 + N) // 2
        scores = seg.get_vec_embed_neg_len_sqrd_norms(utts.vec_ids[u, :n_packed], utts.durations[u, :n_packed])
        obj, b = ns.forward_backward_kmeans_viterbi(scores, N, seg.n_slices_min, seg.n_slices_max, u)
        utts.boundaries[u, :N] = b
        totals.append(float(obj))
        bounds.append(np.asarray(b, dtype=bool).copy())
        ks.append([int(k) for k in comps.get_max_assignments(utts.get_segmented_embeds_i(u))])
    return totals, bounds, ks


_POOL = {}


def _pool_init(X, lengths, seg_id, seg_dur, means):
    """Worker initialiser: the corpus sample and the model reach every process ONCE (fork + initargs),
    not once per job inside the timed region."""
    kind, ns = cpu_modules()
    _POOL.update(X=X, lengths=lengths, seg_id=seg_id, seg_dur=seg_dur, means=means, ns=ns)


def _pool_job(span):
    u0, u1 = span
    P = _POOL
    pos_off = np.concatenate([[0], np.cumsum(P["lengths"])])
    lo, hi = pos_off[u0], pos_off[u1]
    ids = P["seg_id"][lo:hi]
    e_lo, e_hi = ids[ids >= 0].min(), ids[ids >= 0].max() + 1
    sub_ids = np.where(ids >= 0, ids - e_lo, -1)
    seg = _cpu_segmenter(P["ns"], P["X"][e_lo:e_hi], P["lengths"][u0:u1], sub_ids, P["seg_dur"][lo:hi], P["means"])
    t0 = time.perf_counter()
    totals, _, _ = _cpu_phase1(P["ns"], seg)
    return time.perf_counter() - t0, totals


def cpu_sample(n_utt, K):
    """The CPU-reproducible head of rank 0's shard (its first n_utt <= CPU_SAMPLE_UTTS utterances):
    structure, embeddings and the generating centres (the model both arms score against)."""
    lengths, seg_id, seg_dur, _, n_emb = corpus_structure(CPU_SAMPLE_UTTS, seed=999)
    centres = centres_cpu(K)
    X, z = sample_rows_cpu(n_emb, centres, seed=3000)
    n_utt = min(n_utt, CPU_SAMPLE_UTTS)
    n_pos = int(np.sum(lengths[:n_utt]))
    ids = seg_id[:n_pos]
    e_hi = int(ids.max()) + 1
    return lengths[:n_utt], ids, seg_dur[:n_pos], X[:e_hi], z[:e_hi], centres


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path (baseline/_ref; kind "reference";
    the oracle port, kind "port", only where that directory is missing) on all host cores, one process per
    core, on the head of the SAME seeded corpus the GPU arm uses (rank 0's first utterances) against the
    K-component model of the generating centres."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, _ = cpu_modules()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_core = 2
    n_utt = min(cores * per_core, CPU_SAMPLE_UTTS)
    lengths, seg_id, seg_dur, X, _, centres = cpu_sample(n_utt, args.K)
    ctx = mp.get_context("fork")

    def run_pool(n_proc, n_iter):
        per = -(-n_utt // n_proc)
        spans = [(lo, min(n_utt, lo + per)) for lo in range(0, n_utt, per)]
        out = []
        with ctx.Pool(n_proc, initializer=_pool_init, initargs=(X, lengths, seg_id, seg_dur, centres)) as pool:
            for _ in range(n_iter):
                t0 = time.perf_counter()
                pool.map(_pool_job, spans)
                out.append(time.perf_counter() - t0)
        return out
    # NumPy's large temporaries make the CPU path memory-bound; on some hosts fewer processes than
    # cores are faster.  Calibrate once (untimed) and use the best process count.
    calib = {n: run_pool(n, 2)[1] for n in sorted({1, max(1, cores // 2), min(cores, n_utt)})}   # 2nd pass: imports warm
    used = min(calib, key=calib.get)
    times = run_pool(used, args.warmup + args.steps)[args.warmup:]
    per_step = float(np.mean(times))
    value = n_utt / per_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "utt/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "K": args.K, "D": D, "max_span": S_MAX,
                   "sample": "the first %d utterances of the GPU arm's corpus (rank 0, same seeds) per step, scored "
                             "against the full K=%d model" % (n_utt, args.K)},
        "cpu_baseline": {"value": value, "unit": "utt/s", "cores": used, "kind": kind,
                         "sample": "%d utterances per step, %d processes (host has %d cores; calibration s/step: %s); "
                                   "%s" % (n_utt, used, cores, {k: round(v, 2) for k, v in calib.items()},
                                           "kamperh/segmentalist's own classes and functions from baseline/_ref "
                                           "(py2->py3 text shim, numerics untouched)" if kind == "reference"
                                           else "oracle port (baseline/_ref missing)")},
        "e2e": {"value": value, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# secondary workload: sequential collapsed Gibbs (BASELINE configs[1]); replicas only, so N = 1
# ------------------------------------------------------------------------------------------------

def measure_fp64_peak():
    """Dense float64 throughput of this GPU (TFLOP/s): cuBLAS DGEMM 4096^3, best of 3 -- the denominator of the
    Gibbs sweep's float64 roofline (MEASURED_PEAKS.json has no float64 entry)."""
    import torch
    try:
        n = 4096
        a = torch.randn(n, n, dtype=torch.float64, device="cuda")
        b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        torch.matmul(a, b)
        best = None
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def run_gibbs_extra(args):
    """UnigramAcousticWordseg.gibbs_sample on synthetic D=130, K=1000, 2k utterances, max_span 6,
    through the reference-facing API; CPU oracle timed on a bounded sample of the same corpus
    with identical seeds (so the first utterances are also a parity check)."""
    import random

    import torch
    from oracle import seg_oracle as so
    from segmentalist_b200 import fbgmm, gaussian_components_fixedvar as gcf, synth
    from segmentalist_b200 import unigram_acoustic_wordseg as uaw
    K, n_utt = 1000, args.gibbs_utts
    mats, vids, durs, lms = synth.make_corpus_dicts(n_utt, D=D, K_true=K, n_min=N_LO, n_max=N_HI,
                                                    n_slices_max=S_MAX, noise=NOISE, seed=31)
    var = 0.002 * np.ones(D)
    prior = gcf.FixedVarPrior(var, np.zeros(D), var / 0.05)
    random.seed(3)
    np.random.seed(3)
    t0 = time.perf_counter()
    seg = uaw.UnigramAcousticWordseg(fbgmm.FBGMM, 10., K, prior, mats, vids, durs, lms, p_boundary_init=0.5,
                                     beta_sent_boundary=-1, n_slices_max=S_MAX)
    setup_s = time.perf_counter() - t0
    n_seg = int(sum(m.shape[0] for m in mats.values()))
    order = list(range(n_utt))
    seg._sweep(order, 1, False)                       # warm-up sweep
    torch.cuda.synchronize()
    reps = 2
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(reps):
        seg._sweep(order, 1, False)
    ev1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    dev_s = ev0.elapsed_time(ev1) * 1e-3 / reps
    K_act = seg.acoustic_model.components.K
    n_tok0 = seg.acoustic_model.get_n_assigned()
    # float64 roofline: 4 flop per (segment or token, ACTIVE component, dimension): mu - x, its square, times the
    # predictive precision, accumulate (gaussian_components_fixedvar.py:247-252); candidates are scored once per
    # sweep, tokens once more for their assignment draw
    fp64_flops = 4.0 * D * K_act * (n_seg + n_tok0)
    peak64 = measure_fp64_peak()
    out = {"workload": "unigram_fbgmm_fixedvar_gibbs_sweep D=130 K=1000 U=%d max_span=6 (BASELINE configs[1])" % n_utt,
           "utt_per_s": n_utt / max(wall, dev_s), "ms_per_sweep": max(wall, dev_s) * 1e3,
           "device_ms_per_sweep": dev_s * 1e3, "us_per_utterance": dev_s * 1e6 / n_utt, "candidate_segments": n_seg,
           "segment_component_evals_per_s": n_seg * K / max(wall, dev_s),
           "K_active": K_act, "setup_s": setup_s, "dtype": "f64",
           "roofline": {"kernel": "fv_gibbs_kernel (cooperative, sequential collapsed Gibbs: latency-bound by construction)",
                        "bound": "fp64", "achieved": fp64_flops / dev_s / 1e12, "peak": peak64, "unit": "TFLOP/s",
                        "frac": fp64_flops / dev_s / 1e12 / peak64 if peak64 else None,
                        "peak_source": "measured in this run: torch.matmul float64 4096^3 (cuBLAS DGEMM), best of 3",
                        "algorithmic_flops_per_sweep": fp64_flops,
                        "note": "one chain cannot fill the GPU: every token's draw sees the statistics left by the previous one"}}
    # replicas: R independent chains side by side, each a cooperative launch with n_sm / R CTAs
    try:
        reps_out = []
        for R in (2, 4):
            segs, rngs = [], []
            for r in range(R):
                random.seed(100 + r)
                np.random.seed(100 + r)
                segs.append(uaw.UnigramAcousticWordseg(fbgmm.FBGMM, 10., K, prior, mats, vids, durs, lms, p_boundary_init=0.5,
                                                       beta_sent_boundary=-1, n_slices_max=S_MAX))
                rngs.append(random.Random(200 + r))
            orders = [order] * R
            uaw.run_replica_sweeps(segs, orders, rngs)             # warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            uaw.run_replica_sweeps(segs, orders, rngs)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            reps_out.append({"replicas": R, "ms_per_replica_sweep": dt * 1e3, "aggregate_utt_per_s": R * n_utt / dt,
                             "us_per_utterance_per_chain": dt * 1e6 / n_utt})
            del segs
        out["replicas"] = reps_out
    except Exception as exc:
        out["replicas"] = {"error": repr(exc)}
    # whole-model resampling between sweeps (FBGMM.gibbs_sample, consider_unassigned=False): one
    # cooperative launch over all assigned tokens
    n_tok = seg.acoustic_model.get_n_assigned()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    seg.acoustic_model.gibbs_sample(1, consider_unassigned=False)
    torch.cuda.synchronize()
    am_s = time.perf_counter() - t0
    out["am_gibbs_sample"] = {"tokens": int(n_tok), "ms": am_s * 1e3, "tokens_per_s": n_tok / am_s,
                              "note": "includes the host-side record (log_marg) of the reference API"}
    if not args.no_cpu:
        # CPU oracle: same construction + seeds, a bounded number of gibbs_sample_i calls
        n_cpu = 24
        random.seed(3)
        np.random.seed(3)
        oprior = so.FixedVarPrior(var, np.zeros(D), var / 0.05)
        oseg = so.UnigramAcousticWordseg(so.FBGMM, 10., K, oprior, mats, vids, durs, lms, p_boundary_init=0.5,
                                         beta_sent_boundary=-1, n_slices_max=S_MAX)
        random.seed(3)
        np.random.seed(3)
        gseg = uaw.UnigramAcousticWordseg(fbgmm.FBGMM, 10., K, prior, mats, vids, durs, lms, p_boundary_init=0.5,
                                          beta_sent_boundary=-1, n_slices_max=S_MAX)
        st = random.getstate()
        t0 = time.perf_counter()
        for u in range(n_cpu):
            oseg.gibbs_sample_i(u)
        dt = time.perf_counter() - t0
        random.setstate(st)
        gseg._sweep(list(range(n_cpu)), 1, False)
        same = bool(np.array_equal(gseg.utterances.boundaries[:n_cpu], oseg.utterances.boundaries[:n_cpu]) and
                    np.array_equal(gseg.acoustic_model.components.assignments,
                                   oseg.acoustic_model.components.assignments))
        out["cpu_baseline"] = {"value": n_cpu / dt, "unit": "utt/s", "cores": 1, "kind": "port",
                               "sample": "%d gibbs_sample_i calls of the same seeded corpus/model" % n_cpu,
                               "seconds": dt, "identical_samples_on_sample": same}
    return out


def run_diag_extra(args):
    """BASELINE configs[4]: diagonal-covariance FBGMM unigram segmentation, D=130, K_max=5000, long
    utterances (100-120 landmarks), one cooperative Gibbs sweep through the reference-facing API; CPU
    oracle timed on the first utterances of the same seeded corpus (also a parity check)."""
    import random

    import torch
    from oracle import seg_oracle as so
    from segmentalist_b200 import fbgmm, synth, unigram_acoustic_wordseg as uaw
    from segmentalist_b200.niw import NIW
    K, n_utt = 5000, args.diag_utts
    # K_true = K_max generating clusters and the reference's "one-by-one" initialisation (every token is added by a
    # collapsed Gibbs draw against the tokens before it, unigram_acoustic_wordseg.py:225-236): cluster mates join
    # the component their first token opened, so the chain starts -- and stays -- with about one component per
    # generating cluster that has tokens.  (A "rand" start puts 3 unrelated tokens into each of the 5000 slots;
    # with no free slot to open, the chain then collapses everything into one broad component.)
    mats, vids, durs, lms = synth.make_corpus_dicts(n_utt, D=D, K_true=args.diag_k_true, n_min=100, n_max=120,
                                                    n_slices_max=S_MAX, noise=NOISE, seed=53)
    # S_0 / v_0 = the corpus's within-cluster variance scale (with S_0 = 0.002 a new component's predictive is 130x
    # narrower than the data and a long chain collapses every token into one diffuse component)
    prior_args = dict(m_0=np.zeros(D), k_0=0.05, v_0=D + 3, S_0=0.002 * (D + 3) * np.ones(D))

    def build(mod, am_mod, prior, init=None):
        random.seed(5)
        np.random.seed(5)
        return mod.UnigramAcousticWordseg(am_mod.FBGMM, 10., K, prior, mats, vids, durs, lms, p_boundary_init=0.5,
                                          beta_sent_boundary=-1, n_slices_max=S_MAX, covariance_type="diag",
                                          init_am_assignments=init or args.diag_init)
    t_setup = time.perf_counter()
    seg = build(uaw, fbgmm, NIW(**prior_args))
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup
    K_init = seg.acoustic_model.components.K
    n_seg = int(sum(m.shape[0] for m in mats.values()))
    order = list(range(n_utt))
    seg._sweep(order, 1, False)                       # warm-up sweep
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    seg._sweep(order, 1, False)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    K_act = seg.acoustic_model.components.K
    out = {"workload": "unigram_fbgmm_diag_gibbs_sweep D=130 K_max=5000 K_true=%d U=%d N~U{100..120} max_span=6 %s init "
                       "(BASELINE configs[4])" % (args.diag_k_true, n_utt, args.diag_init),
           "utt_per_s": n_utt / wall, "ms_per_sweep": wall * 1e3, "candidate_segments": n_seg, "K_active": K_act,
           "K_after_init": K_init, "setup_s": t_setup,
           "student_t_log_evals_per_s": n_seg * float(K_act) * D / wall, "dtype": "f64",
           "tokens": int(seg.acoustic_model.get_n_assigned()),
           "roofline": {"bound": "sfu (SURVEY 8d: D log evaluations per segment x component)",
                        "achieved": n_seg * float(K_act) * D / wall, "peak": 16.0 * 148 * 1.965e9, "unit": "log evals/s",
                        "frac": n_seg * float(K_act) * D / wall / (16.0 * 148 * 1.965e9),
                        "peak_source": "16 MUFU lg2/clk/SM x 148 SMs x 1965 MHz (float32 special-function rate; nominal)",
                        "note": "the reference's Student's t terms are float64 (gaussian_components_diag.py:347-360): log() is a "
                                "~40-instruction float64 sequence here, not one MUFU op, so the float32 SFU bound is out of "
                                "reach by construction; identical samples need float64"},
           "note": "evals counted over the ACTIVE components of the sampled state"}
    if not args.no_cpu:
        # the pure-Python oracle cannot afford the one-by-one start (one K x D Student's t evaluation per token in
        # NumPy): parity and the CPU rate are taken on the "rand" start of the same corpus (all 5000 slots occupied)
        n_cpu = 2
        oseg = build(so, so, so.NIW(**prior_args), "rand")
        gseg = build(uaw, fbgmm, NIW(**prior_args), "rand")
        st = random.getstate()
        t0 = time.perf_counter()
        for u in range(n_cpu):
            oseg.gibbs_sample_i(u)
        dt = time.perf_counter() - t0
        random.setstate(st)
        gseg._sweep(list(range(n_cpu)), 1, False)
        same = bool(np.array_equal(gseg.utterances.boundaries[:n_cpu], oseg.utterances.boundaries[:n_cpu]) and
                    np.array_equal(gseg.acoustic_model.components.assignments,
                                   oseg.acoustic_model.components.assignments))
        out["cpu_baseline"] = {"value": n_cpu / dt, "unit": "utt/s", "cores": 1, "kind": "port",
                               "sample": "%d gibbs_sample_i calls of the same seeded corpus, 'rand' start (5000 occupied slots)" % n_cpu,
                               "seconds": dt, "identical_samples_on_sample": same}
    return out


def run_bigram_extra(args):
    """BASELINE configs[3]: bigram FBGMM cluster sampling (bigram_acoustic_wordseg, smoothed ML bigram LM),
    D=130, K=5000, max_span 6, through the reference-facing API: one full sweep (segmentation + bigram
    assignment sampling) and one assignments-only sweep (pure cluster sampling); CPU oracle timed on the
    first utterances of the same seeded corpus (also a parity check)."""
    import random

    import torch
    from oracle import seg_oracle as so
    from segmentalist_b200 import bigram_acoustic_wordseg as baw, gaussian_components_fixedvar as gcf, synth
    K, n_utt = 5000, args.bigram_utts
    mats, vids, durs, lms = synth.make_corpus_dicts(n_utt, D=D, K_true=K, n_min=N_LO, n_max=N_HI,
                                                    n_slices_max=S_MAX, noise=NOISE, seed=41)
    var = 0.002 * np.ones(D)
    lm_params = {"type": "smooth", "intrp_lambda": 0.1, "a": 10.0, "b": 10.0}

    def build(mod, prior):
        random.seed(4)
        np.random.seed(4)
        return mod.BigramAcousticWordseg(K, prior, lm_params, mats, vids, durs, lms, p_boundary_init=0.5,
                                         beta_sent_boundary=-1, n_slices_max=S_MAX, fb_type="unigram")
    seg = build(baw, gcf.FixedVarPrior(var, np.zeros(D), var / 0.05))
    order = list(range(n_utt))
    seg._sweep(order, 1, False, False)                # warm-up sweep
    torch.cuda.synchronize()
    n_tok = seg.acoustic_model.get_n_assigned()
    t0 = time.perf_counter()
    seg._sweep(order, 1, False, False)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t0 = time.perf_counter()
    seg._sweep(order, 1, False, True)
    torch.cuda.synchronize()
    wall_a = time.perf_counter() - t0
    n_seg = int(sum(m.shape[0] for m in mats.values()))
    out = {"workload": "bigram_fbgmm_cluster_sampling D=130 K=K_true=5000 U=%d max_span=6 (BASELINE configs[3])" % n_utt,
           "utt_per_s": n_utt / wall, "ms_per_sweep": wall * 1e3, "candidate_segments": n_seg,
           "assignments_only": {"ms_per_sweep": wall_a * 1e3, "tokens": int(n_tok), "tokens_per_s": n_tok / wall_a,
                                "note": "one K_max-slot draw per token under the bigram prior row of the previous label"},
           "K_active": seg.acoustic_model.components.K, "dtype": "f64",
           "note": "full sweeps: one cooperative launch (components sharded over the SMs, CTA 0 keeps the LM); assignments-only sweeps: four launches per utterance on one stream; no host synchronisation inside a sweep"}
    if not args.no_cpu:
        n_cpu = 4
        oseg = build(so, so.FixedVarPrior(var, np.zeros(D), var / 0.05))
        gseg = build(baw, gcf.FixedVarPrior(var, np.zeros(D), var / 0.05))
        st = random.getstate()
        t0 = time.perf_counter()
        for u in range(n_cpu):
            oseg.gibbs_sample_i(u)
        dt = time.perf_counter() - t0
        random.setstate(st)
        gseg._sweep(list(range(n_cpu)), 1, False, False)
        same = bool(np.array_equal(gseg.utterances.boundaries[:n_cpu], oseg.utterances.boundaries[:n_cpu]) and
                    np.array_equal(gseg.acoustic_model.components.assignments,
                                   oseg.acoustic_model.components.assignments) and
                    np.array_equal(gseg.lm.unigram_counts, oseg.lm.unigram_counts))
        out["cpu_baseline"] = {"value": n_cpu / dt, "unit": "utt/s", "cores": 1, "kind": "port",
                               "sample": "%d gibbs_sample_i calls of the same seeded corpus/model" % n_cpu,
                               "seconds": dt, "identical_samples_on_sample": same}
    return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

# ------------------------------------------------------------------------------------------------
# helpers of the GPU arm
# ------------------------------------------------------------------------------------------------

def _checksum(t):
    """Order-sensitive 64-bit checksum of a tensor's bits (device, no host copy of the data)."""
    import torch
    v = t.contiguous().reshape(-1).view(torch.uint8).to(torch.int64)
    w = (torch.arange(v.numel(), device=v.device, dtype=torch.int64) % 65521) + 1
    return torch.stack([v.sum(), (v * w).sum()])


def _sub_corpus(lengths, seg_id, seg_dur, bounds0, lo_u, hi_u):
    """Utterances [lo_u, hi_u) of a flat corpus with embedding ids rebased to the shard."""
    pos_off = np.concatenate([[0], np.cumsum(lengths)])
    p0, p1 = int(pos_off[lo_u]), int(pos_off[hi_u])
    ids = seg_id[p0:p1]
    e_lo, e_hi = int(ids[ids >= 0].min()), int(ids.max()) + 1
    return lengths[lo_u:hi_u], np.where(ids >= 0, ids - e_lo, -1).astype(np.int32), seg_dur[p0:p1], bounds0[p0:p1], p0, p1, e_lo, e_hi


def multi_gpu_parity(args, world, rank, dev, sweep, comps):
    """(a) after the sweeps' all-reduce every rank holds bit-identical means / numerators / counts;
    (b) an N-rank sweep over a small fixed corpus equals the 1-rank sweep of the same corpus: k-means
    (means bit for bit, boundaries, assignments; the initial model is diffuse, so inactive slots win tokens
    and the device clamp / compaction run across ranks) and the frozen FBGMM sweep (decisions identical,
    statistics to 1e-13)."""
    import torch
    import torch.distributed as dist
    from segmentalist_b200 import sharding
    from segmentalist_b200.batch import FrozenFBGMMSweep, FrozenKMeansSweep
    from segmentalist_b200.gaussian_components_fixedvar import FixedVarPrior, GaussianComponentsFixedVar
    from segmentalist_b200.kmeans_components import KMeansComponents
    from segmentalist_b200.utterances import DeviceCorpus
    out = {}
    if comps is not None:
        cs = torch.cat([_checksum(comps._means), _checksum(comps._mean_num), _checksum(comps._counts)])
        if world > 1:
            allc = [torch.empty_like(cs) for _ in range(world)]
            dist.all_gather(allc, cs)
            out["means_identical_across_ranks"] = bool(all(torch.equal(a, allc[0]) for a in allc))
        else:
            out["means_identical_across_ranks"] = True

    # ---- small fixed corpus, identical on every rank
    U_s, K_s, K_true = 1024, 256, 64
    lengths, seg_id, seg_dur, bounds0, n_emb = corpus_structure(U_s, seed=4242)
    centres = torch.from_numpy(centres_cpu(K_true)).to(dev)
    X, _ = make_embeddings_gpu(n_emb, centres, seed=4243, device=dev)
    rnd = X[torch.arange(K_s, device=dev) * 37 % n_emb].clone()
    var = 0.002 * np.ones(D)
    n_pos = int(lengths.sum())
    g = torch.Generator(device=dev).manual_seed(99)
    uni = [(torch.rand(n_pos, dtype=torch.float64, device=dev, generator=g),
            torch.rand(n_pos, dtype=torch.float64, device=dev, generator=g)) for _ in range(2)]

    def run(lo_u, hi_u, local):
        ln, ids, dur, b0, p0, p1, e_lo, e_hi = _sub_corpus(lengths, seg_id, seg_dur, bounds0, lo_u, hi_u)
        Xs = X[e_lo:e_hi].contiguous()
        res = {}
        ctx = sharding.local_only() if local else None
        if ctx:
            ctx.__enter__()
        try:
            # k-means: diffuse initial model (token i -> component (global id * 7919) % K_s)
            corpus = DeviceCorpus(ln, ids, dur, b0, 0, S_MAX, S_MAX)
            c = KMeansComponents.from_device(Xs, K_s, rnd)
            tok = corpus.tok_id[corpus.tok_id >= 0].long()
            c._assign[tok] = (((tok + e_lo) * 7919) % K_s).to(torch.int32)
            sw = FrozenKMeansSweep(c, corpus, wip=0.0, scorer="mma")
            sw.init_means_from_assignments()
            tot = [sw.sweep() for _ in range(3)]
            res["km"] = (c._means.clone(), corpus.bounds.clone(), c._assign.clone(), int(c._K.item()), tot)
            # frozen FBGMM sweep (FFBS + sampled components), trained-like start
            corpus2 = DeviceCorpus(ln, ids, dur, b0, 0, S_MAX, S_MAX)
            f = GaussianComponentsFixedVar.from_device(Xs, FixedVarPrior(var, np.zeros(D), var / 0.05), K_s, alpha=10., lms=1.0)
            tok = corpus2.tok_id[corpus2.tok_id >= 0].long()
            f._assign[tok] = (((tok + e_lo) * 7919) % K_s).to(torch.int32)
            fs = FrozenFBGMMSweep(f, corpus2, fb_type="standard")
            fs.init_from_assignments()
            tot2 = [fs.sweep(u[0][p0:p1].contiguous(), u[1][p0:p1].contiguous()) for u in uni]
            res["fb"] = (f._mu_NT.clone(), corpus2.bounds.clone(), f._assign.clone(), int(f._K.item()), tot2,
                         f._counts.clone())
        finally:
            if ctx:
                ctx.__exit__(None, None, None)
        return res, (p0, p1, e_lo, e_hi)

    full, _ = run(0, U_s, True)
    lo_u, hi_u = sharding.shard_ranges(U_s, world)[rank]
    part, (p0, p1, e_lo, e_hi) = run(lo_u, hi_u, False)
    km_f, km_p = full["km"], part["km"]
    ok_km = (torch.equal(km_f[0], km_p[0]) and torch.equal(km_f[1][p0:p1], km_p[1]) and
             torch.equal(km_f[2][e_lo:e_hi], km_p[2]) and km_f[3] == km_p[3] and
             all(abs(a - b) <= 1e-9 * abs(a) for a, b in zip(km_f[4], km_p[4])))
    fb_f, fb_p = full["fb"], part["fb"]
    ok_fb = (torch.equal(fb_f[1][p0:p1], fb_p[1]) and torch.equal(fb_f[2][e_lo:e_hi], fb_p[2]) and fb_f[3] == fb_p[3] and
             torch.equal(fb_f[5], fb_p[5]) and
             bool(torch.allclose(fb_f[0], fb_p[0], rtol=1e-13, atol=1e-300)) and
             all(abs(a - b) <= 1e-9 * abs(a) for a, b in zip(fb_f[4], fb_p[4])))
    flags = torch.tensor([int(ok_km), int(ok_fb)], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    out["nrank_equals_1rank_kmeans_small_corpus"] = bool(flags[0].item())
    out["nrank_equals_1rank_fbgmm_small_corpus"] = bool(flags[1].item())
    out["small_corpus"] = {"utterances": U_s, "K_max": K_s, "K_after_kmeans": km_f[3], "K_after_fbgmm": fb_f[3],
                           "sweeps": 3, "note": "diffuse initial model: inactive-slot wins, device clamp and compaction across ranks"}
    return out


def fv_logmarg_roofline(args, X, Z, M, peak_tf, peak_src, timed):
    """FBGMM.log_marg_i of n rows x K_max slots on tensor cores: ONE fp16 tcgen05 pass (the filter GEMM)
    + exact float64 refine, against a trained-like model (component = generating cluster)."""
    import torch
    from segmentalist_b200 import fbgmm as fbgmm_mod
    from segmentalist_b200.batch import FvScorer
    from segmentalist_b200.gaussian_components_fixedvar import FixedVarPrior, GaussianComponentsFixedVar
    out = {}
    for aniso, prec in ((False, "fp16"), (False, "fp8"), (True, "fp16")):
        n_fv = min(M, (4 if not aniso else 1) * 1024 * 1024)
        rng = np.random.RandomState(0)
        var = 0.002 * (0.5 + rng.rand(D)) if aniso else 0.002 * np.ones(D)
        var_0 = 0.04 * (0.5 + rng.rand(D)) if aniso else var / 0.05
        am = fbgmm_mod.FBGMM.__new__(fbgmm_mod.FBGMM)
        am.alpha, am.lms, am.covariance_type = 10., 1.0, "fixed"
        am.components = GaussianComponentsFixedVar.from_device(X[:n_fv], FixedVarPrior(var, np.zeros(D), var_0),
                                                               args.K, alpha=10., lms=1.0)
        n_tok = min(n_fv, 20 * args.K)
        zh = Z[:n_tok].cpu().numpy()
        _, first = np.unique(zh, return_index=True)        # labels in order of first appearance
        rank_of = np.empty(args.K, dtype=np.int64)
        rank_of[zh[np.sort(first)]] = np.arange(len(first))
        am.components._add_many(np.arange(n_tok), rank_of[zh])
        fv = FvScorer(am.components, precision=prec)
        fv.score()
        pk_tf, pk_src = (2.0 * peak_tf, "2 x " + peak_src + " (e4m3 rate; no measured fp8 entry)") if fv.fp8 else (peak_tf, peak_src)
        if fv.fused:
            t_pack, t_filter, t_refine = timed(fv.pack_model, 3), timed(fv.fused_score, 3), 0.0
        else:
            t_pack, t_filter, t_refine = timed(fv.pack_model, 3), timed(fv.filter, 3), timed(fv.refine, 3)
        fl = (4.0 if aniso else 2.0) * D * n_fv * args.K
        ids = np.arange(n_tok, n_tok + 2048) if n_fv >= n_tok + 2048 else np.arange(min(2048, n_fv))
        exact = am.log_marg_items(ids)
        got = fv.log_marg[torch.from_numpy(ids).to(fv.log_marg.device)].cpu().numpy()
        rel = float((np.abs(got - exact) / np.abs(exact)).max())
        kp = 16 * ((D + (3 if aniso else 6) + 15) // 16) * (2 if aniso else 1)
        if fv.fp8:
            kp = 32 * ((D + 21 + 31) // 32)
        r = {"kernel": ("score_fused_kernel<fv> (fp32 rows in, ONE fp16 tcgen05 pass, exact float64 re-scoring of the kept "
                        "components + logsumexp behind the GEMM: the whole log_marg_i step)" if fv.fused else
                        "kmeans_filter_kernel<5,1,0,F8> as the log_marg_i filter (ONE e4m3 tcgen05 pass, kind::f8f6f4) + fv_refine_kernel "
                        "(exact float64 re-scoring of the kept components, logsumexp); undecided rows -> exhaustive scan" if fv.fp8 else
                        "kmeans_filter_kernel<%s> as the log_marg_i filter (ONE fp16 tcgen05 pass, fp32 TMEM, top-3 chunk epilogue) "
                        "+ fv_refine_kernel (exact float64 re-scoring of the kept components, logsumexp)" % ("9,2" if aniso else "9,1")),
             "bound": "tensor", "achieved": fl / (t_filter * 1e-3) / 1e12, "peak": pk_tf, "unit": "TFLOP/s",
             "frac": fl / (t_filter * 1e-3) / 1e12 / pk_tf, "peak_source": pk_src,
             "kernel_ms": t_filter, "refine_ms": t_refine, "pack_model_ms": t_pack,
             "frac_incl_refine_and_pack": fl / ((t_filter + t_refine + t_pack) * 1e-3) / 1e12 / pk_tf,
             "rows": n_fv, "K": args.K, "K_active": am.components.K, "anisotropic_variances": aniso,
             "algorithmic_flops_per_launch": fl,
             "algorithmic_flops_note": "%d*D per segment x component evaluation (SURVEY 8d)" % (4 if aniso else 2),
             "executed_tflops": 2.0 * kp * n_fv * 128 * ((args.K + 1 + 127) // 128) / (t_filter * 1e-3) / 1e12,
             "fallback_rows": int(fv.n_fallback.item()), "max_rel_err_vs_exact_float64": rel,
             "threshold_nats": fv.T}
        if fv.fp8:
            out["e4m3_first_level"] = r
        elif not aniso:
            tr = ncu_traffic("score_fused_kernel_fv" if fv.fused else "kmeans_filter_kernel@r2_raw_fv.csv")
            r["traffic"] = tr["bytes_per_launch"] if (tr and args.K == K_MAX and n_fv == 4 * 1024 * 1024) else None
            r["traffic_source"] = tr.get("source") if tr else None
            r["algorithmic_bytes_per_launch"] = float(n_fv * D * 4 + 12 * n_fv) if fv.fused else float(fv.x_tiles.numel() + fv.cand.numel())
            out = r
        else:
            out["anisotropic"] = r
        del fv, am
    return out


def world_has_head(n_head, world):
    """Whether rank 0's shard starts with the CPU-reproducible head -- decided identically on every rank
    (n_head itself is 0 on the other ranks)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return bool(n_head)
    t = torch.tensor([int(bool(n_head))], device="cuda", dtype=torch.int32)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(t.item())


def fbgmm_frozen_secondary(args, world, rank, dev, X, Z, corpus, lengths, seg_id, seg_dur, n_head, peak_tf, barrier,
                           max_over_ranks):
    """The sharded frozen-model sweep of the unigram FBGMM segmenter (UnigramAcousticWordseg.segment_frozen's
    engine) on the SAME corpus as the headline: tensor-core log_marg_i -> banded scores -> batched FFBS ->
    sampled components -> one all-reduce of [sum_x | counts] -> closed-form rebuild.  Device time, max over
    ranks; CPU-oracle parity on the head of rank 0's shard."""
    import torch
    import torch.distributed as dist
    from segmentalist_b200 import _lib
    from segmentalist_b200.batch import FrozenFBGMMSweep
    from segmentalist_b200.gaussian_components_fixedvar import FixedVarPrior, GaussianComponentsFixedVar
    var = 0.002 * np.ones(D)
    comps = GaussianComponentsFixedVar.from_device(X, FixedVarPrior(var, np.zeros(D), var / 0.05), args.K, alpha=10., lms=1.0)
    sweep = FrozenFBGMMSweep(comps, corpus, fb_type="standard", time_power_term=1.0, wip=0.0)
    _lib.check(_lib.lib().segb_tokens_from_bounds(corpus.struct(), 0, corpus.n_utt, _lib.stream_ptr()))
    tok = corpus.tok_id[corpus.tok_id >= 0].long()
    comps._assign.fill_(-1)
    comps._assign[tok] = Z[tok]
    sweep.init_from_assignments()
    n_pos = corpus.n_pos
    g = torch.Generator(device=dev).manual_seed(500 + rank)

    def uniforms():
        return (torch.rand(n_pos, dtype=torch.float64, device=dev, generator=g),
                torch.rand(n_pos, dtype=torch.float64, device=dev, generator=g))
    for _ in range(2):
        sweep.sweep(*uniforms())
    barrier()
    steps = max(2, min(args.steps, 3))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fb = 0
    for _ in range(steps):
        sweep.sweep(*uniforms())
        fb += sweep.last_fallback
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    phases = sweep.profile_phases(*uniforms())
    out = {"workload": "unigram_fbgmm_frozen_sweep D=130 K=%d U=%d max_span=6 FFBS + sampled components, sharded x%d"
                       % (args.K, args.utts, world),
           "utt_per_s": args.utts / (ms * 1e-3), "ms_per_sweep": ms, "n_gpus": world, "K_active": sweep.K_host,
           "fallback_rows_per_sweep": fb / steps, "phases_ms": phases,
           "segment_component_evals_per_s": float(X.shape[0]) * args.K * world / (ms * 1e-3), "dtype": "f64 scores (%s tensor filter + float64 refine)" % ("e4m3" if sweep.fv.fp8 else "fp16"),
           "filter_precision": "auto -> " + ("e4m3 first level" if sweep.fv.fp8 else "fp16")}
    # ---- CPU oracle on the head of rank 0's shard: same model state, same uniforms.  The sweep itself is a
    # collective (all-reduce of the statistics): EVERY rank runs it; rank 0 keeps the model state from before it.
    check = (not args.no_cpu) and world_has_head(n_head, world)
    if check:
        pre = None
        if rank == 0:
            pre = dict(K=comps.K, mu_N_numerators=comps.mu_N_numerators, precision_Ns=comps.precision_Ns,
                       precision_preds=comps.precision_preds, log_prod=comps.log_prod_precision_preds, counts=comps.counts)
        u_fb, u_as = uniforms()
        sweep.sweep(u_fb, u_as)
        torch.cuda.synchronize()
    if check and rank == 0:
        from oracle import seg_oracle as so
        n_s = min(16, corpus.n_utt)
        hi = int(corpus.pos_off_h[n_s])
        ids = seg_id[:hi]
        e_hi = int(ids.max()) + 1
        oc = so.FixedVarComponents.__new__(so.FixedVarComponents)
        oc.X = X[:e_hi].cpu().numpy()
        oc.N, oc.D, oc.K_max, oc.K = e_hi, D, args.K, pre["K"]
        oc.precision, oc.mu_0, oc.precision_0 = comps.precision, comps.mu_0, comps.precision_0
        oc.mu_N_numerators, oc.precision_Ns = pre["mu_N_numerators"], pre["precision_Ns"]
        oc.precision_preds, oc.log_prod_precision_preds = pre["precision_preds"], pre["log_prod"]
        oc.counts, oc.lm = pre["counts"], None
        oc.neg_half_D_log_2pi = -0.5 * D * np.log(2. * np.pi)
        oc.assignments = -1 * np.ones(e_hi, dtype=np.int64)
        am = so.FBGMM.__new__(so.FBGMM)
        am.alpha, am.lms, am.covariance_type, am.components, am.prior = 10., 1.0, "fixed", oc, None
        oseg = so.UnigramAcousticWordseg.__new__(so.UnigramAcousticWordseg)
        oseg.utterances = _cpu_segmenter(so, oc.X, lengths[:n_s], ids, seg_dur[:hi], np.zeros((1, D), np.float32)).utterances
        oseg.acoustic_model, oseg.fb_type = am, "standard"
        oseg.n_slices_min, oseg.n_slices_max, oseg.wip, oseg.time_power_term, oseg.beta_sent_boundary = 0, S_MAX, 0.0, 1.0, -1
        t0 = time.perf_counter()
        lps, choices = so.frozen_fbgmm_phase1(oseg, u_fb[:hi].cpu().numpy(), u_as[:hi].cpu().numpy(), range(n_s))
        dt = time.perf_counter() - t0
        gpu_b = corpus.bounds[:hi].cpu().numpy().astype(bool)
        cpu_b = np.concatenate([oseg.utterances.boundaries[u, :lengths[u]] for u in range(n_s)])
        gpu_choice = sweep.choice.cpu().numpy()
        gpu_lp = sweep.log_prob[:n_s].cpu().numpy()
        same = bool(np.array_equal(gpu_b, cpu_b) and all(int(gpu_choice[e]) == j for e, j in choices) and
                    np.allclose(gpu_lp, np.asarray(lps), rtol=1e-9))
        out["cpu_baseline"] = {"value": n_s / dt, "unit": "utt/s", "cores": 1, "kind": "port",
                               "sample": "the first %d utterances of rank 0's shard: get_vec_embed_log_probs + forward_backward + "
                                         "component draws of the oracle against the full K=%d model" % (n_s, args.K),
                               "seconds": dt, "identical_decisions_on_sample": same}
    return out


def kmeans_diffuse_secondary(args, world, rank, dev, barrier, max_over_ranks):
    """The hard case for the filter and for the host-free update: K_true = 200 generating clusters, K_max =
    5000 components initialised "spread" (token t -> component t mod K): a diffuse model, components dying
    (K_act < K_max), inactive slots (random data rows) winning tokens, undecided rows.  Three untimed sweeps
    (the first ones are dominated by exhaustive scans of the flat initial model), then timed sweeps; device
    time, max over ranks; CPU parity (raw argmax incl. inactive-slot wins, boundaries) on rank 0's first
    utterances."""
    import torch
    import torch.distributed as dist
    from segmentalist_b200.batch import FrozenKMeansSweep
    from segmentalist_b200.kmeans_components import KMeansComponents
    from segmentalist_b200.utterances import DeviceCorpus
    total = args.diffuse_utts
    n_utt = total // world + (1 if rank < total % world else 0)
    lengths, seg_id, seg_dur, bounds0, n_emb = corpus_structure(n_utt, seed=7000 + rank)
    centres = torch.from_numpy(centres_cpu(200)).to(dev)
    X, _ = make_embeddings_gpu(n_emb, centres, seed=7100 + rank, device=dev)
    corpus = DeviceCorpus(lengths, seg_id, seg_dur, bounds0, 0, S_MAX, S_MAX)
    rnd = X[torch.randperm(n_emb, device=dev, generator=torch.Generator(device=dev).manual_seed(8))[:args.K]].clone()
    if world > 1:
        dist.broadcast(rnd, src=0)
    comps = KMeansComponents.from_device(X, args.K, rnd)
    tok = corpus.tok_id[corpus.tok_id >= 0].long()
    comps._assign[tok] = ((torch.arange(tok.numel(), device=dev) + 1000003 * rank) % args.K).to(torch.int32)
    sweep = FrozenKMeansSweep(comps, corpus, wip=0.0, scorer="mma")
    sweep.init_means_from_assignments()
    traj = []
    for _ in range(3):
        sweep.sweep()
        traj.append((sweep.K_host, sweep.last_fallback))
    barrier()
    steps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sweep.sweep()
        traj.append((sweep.K_host, sweep.last_fallback))
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    out = {"workload": "kmeans_viterbi_frozen_sweep, diffuse model: D=130 K_max=%d K_true=200 U=%d spread init, sharded x%d"
                       % (args.K, total, world),
           "utt_per_s": total / (ms * 1e-3), "ms_per_sweep": ms, "n_gpus": world,
           "K_active_and_fallback_rows_per_sweep_rank0": traj, "phases_ms": sweep.profile_phases(),
           "filter_precision": "auto -> " + ("e4m3 first level" if sweep.mma.fp8 else "fp16 (the e4m3 pass left > 35 % of the rows undecided in the first sweep)"),
           "note": "clamp of inactive-slot winners and clean_components run as device kernels (csrc/frozen.cu), "
                   "lists exchanged with a fixed-size all-gather; no host logic per sweep"}
    if rank == 0 and not args.no_cpu:
        kind, ns = cpu_modules()
        n_s = min(8, corpus.n_utt)
        hi = int(corpus.pos_off_h[n_s])
        ids = seg_id[:hi]
        e_hi = int(ids.max()) + 1
        means_now = comps._means.cpu().numpy()
        sweep.score()
        sweep.segment()
        torch.cuda.synchronize()
        gpu_bounds = corpus.bounds[:hi].cpu().numpy().astype(bool)
        gpu_k = sweep.best_k[:e_hi].cpu().numpy()
        seg = _cpu_segmenter(ns, X[:e_hi].cpu().numpy(), lengths[:n_s], ids, seg_dur[:hi], means_now)
        _, bounds, ks = _cpu_phase1(ns, seg)
        same = bool(np.array_equal(np.concatenate(bounds), gpu_bounds))
        n_inactive = 0
        for u in range(n_s):
            emb = seg.utterances.get_segmented_embeds_i(u)
            same = same and [int(gpu_k[e]) for e in emb] == ks[u]
            n_inactive += sum(1 for k in ks[u] if k >= sweep.K_host)
        out["parity_with_cpu_on_sample"] = {"identical": same, "kind": kind, "utterances": n_s,
                                            "tokens_won_by_inactive_slots": n_inactive, "K_active": sweep.K_host}
    return out


def ingestion_secondary(args, dev):
    """Set-up cost through the PUBLIC constructor at benchmark scale (process_embeddings
    unigram_acoustic_wordseg.py:571-646, Utterances.__init__ utterances.py:74-157, the banded device layout):
    SegmentalKMeansWordseg(K, embedding_mats, vec_ids_dict, durations_dict, landmarks_dict, ...) on
    reference-format dicts of n utterances, timed on the host clock, phase by phase."""
    import psutil
    import torch
    from segmentalist_b200 import kmeans_acoustic_wordseg as kaw, utterances as ut
    n_utt = args.ingest_utts
    if psutil.virtual_memory().available < 60e9:
        n_utt = min(n_utt, 50000)
    t0 = time.perf_counter()
    lengths, seg_id, seg_dur, _, n_emb = corpus_structure(n_utt, seed=9000)
    Xd, _ = make_embeddings_gpu(n_emb, torch.from_numpy(centres_cpu(args.K)).to(dev), seed=9001, device=dev)
    X = Xd.cpu().numpy()
    del Xd
    pos_off = np.concatenate([[0], np.cumsum(lengths)])
    emb_off = np.concatenate([[0], np.cumsum((seg_id >= 0).sum(axis=1))])[pos_off]      # embeddings before each utterance
    mats, vids, durs, lms = {}, {}, {}, {}
    S = S_MAX
    for N in range(N_LO, N_HI + 1):                       # reference-format packed vectors, one length group at a time
        idx = np.where(lengths == N)[0]
        if len(idx) == 0:
            continue
        t = np.repeat(np.arange(1, N + 1), np.arange(1, N + 1))
        j = np.arange(len(t)) - t * (t - 1) // 2
        l = t - j
        ok = l <= S
        rows = pos_off[idx][:, None] + (t - 1)[None, :]
        cols = np.where(ok, l - 1, 0)[None, :]
        ids = np.where(ok[None, :], seg_id[rows, cols] - emb_off[idx][:, None], -1).astype(np.int64)
        du = np.where(ok[None, :], np.nan_to_num(seg_dur[rows, cols], nan=-1.), -1).astype(np.int64)
        for a, u in enumerate(idx):
            label = "utt%07d" % u
            vids[label], durs[label] = ids[a], du[a]
            mats[label] = X[emb_off[u]:emb_off[u + 1]]
            lms[label] = list(range(1, N + 1))
    t_gen = time.perf_counter() - t0
    import random
    random.seed(1)
    np.random.seed(1)
    t0 = time.perf_counter()
    emb, vec_ids, labels = ut.process_embeddings(mats, vids)
    t_pe = time.perf_counter() - t0
    del emb, vec_ids
    t0 = time.perf_counter()
    seg = kaw.SegmentalKMeansWordseg(args.K, mats, vids, durs, lms, n_slices_max=S, p_boundary_init=0.5,
                                     init_am_assignments="spread")
    torch.cuda.synchronize()
    t_ctor = time.perf_counter() - t0
    t0 = time.perf_counter()
    rec = seg.segment_frozen(1)
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    return {"utterances": int(n_utt), "embeddings": int(n_emb), "setup_s": t_ctor,
            "process_embeddings_s": t_pe, "first_frozen_sweep_s": t_first, "input_generation_s": t_gen,
            "embeddings_per_s": n_emb / t_ctor, "components_after_first_sweep": rec["components"][-1],
            "note": "setup_s = the whole public constructor (process_embeddings + Utterances incl. the random boundary "
                    "initialisation + banded device layout + H2D of X + KMeans construction); vectorised host code, "
                    "no per-utterance Python loop, no padded-triangular intermediate"}


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPUs NVML reports as local to its GPU (and let first-touch place the pinned
    staging buffers on that NUMA node).  At 4-8 ranks the end-to-end path is host-bound: in round 1 every
    rank ran on NUMA node 0's cores with its pinned buffers wherever the process had landed.  Best effort:
    a cpuset that excludes the GPU's node leaves the affinity unchanged."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        ideal = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = ideal & allowed
        info.update(gpu_local_cpus=len(ideal), allowed_cpus=len(allowed), usable=len(use))
        if use:
            os.sched_setaffinity(0, use)
            info["bound"] = True
            info["cpus"] = "%d-%d" % (min(use), max(use))
    except Exception as exc:
        info["error"] = repr(exc)
    return info


def ncu_traffic(kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of `kernel` from the committed
    ncu capture (profiles/r2_ncu_traffic.json: written by tools/summarize_ncu.py from the .ncu-rep of this
    command at the default configuration), or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
        return d.get(kernel)
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from segmentalist_b200 import _lib, sharding
    from segmentalist_b200.batch import FrozenFBGMMSweep, FrozenKMeansSweep, FvScorer
    from segmentalist_b200.kmeans_components import KMeansComponents
    from segmentalist_b200.utterances import DeviceCorpus

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        # a rank that fails alone must not leave the others waiting for the driver's time limit
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    lib = _lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([float(v)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # ---- this rank's shard (strong scaling: args.utts in total).  Centres come from NumPy (shared by all
    # ranks and by the CPU arms); rank 0's first CPU_SAMPLE_UTTS utterances are generated on the CPU.
    n_utt = args.utts // world + (1 if rank < args.utts % world else 0)
    lengths, seg_id, seg_dur, bounds0, n_emb, n_head = shard_structure(n_utt, rank)
    centres_h = centres_cpu(args.K)
    X, Z = make_embeddings_gpu(n_emb, torch.from_numpy(centres_h).to(dev), seed=2000 + rank, device=dev)
    if n_head:
        Xh, zh = sample_rows_cpu(n_head, centres_h, seed=3000)
        X[:n_head] = torch.from_numpy(Xh).to(dev)
        Z[:n_head] = torch.from_numpy(zh).to(dev)
    corpus = DeviceCorpus(lengths, seg_id, seg_dur, bounds0, 0, S_MAX, S_MAX)
    perm = torch.randperm(n_emb, device=dev, generator=torch.Generator(device=dev).manual_seed(7))[:args.K]
    rnd = X[perm].clone()
    if world > 1:
        dist.broadcast(rnd, src=0)
    comps = KMeansComponents.from_device(X, args.K, rnd)
    # initial model: every token starts in the component of its generating cluster, so all K_max
    # components are populated and stay alive (K_act = K_max); the diffuse case (K_act < K_max,
    # inactive slots winning tokens, undecided rows) is measured separately below (secondary_kmeans_diffuse)
    tok = corpus.tok_id[corpus.tok_id >= 0].long()
    comps._assign[tok] = Z[tok]
    sweep = FrozenKMeansSweep(comps, corpus, wip=0.0, scorer=args.scorer, fused=bool(args.fused), precision=args.precision)
    sweep.init_means_from_assignments()
    sweep_fused = bool(sweep.mma is not None and sweep.mma.fused)
    filter_precision = None
    x_gb = X.numel() * 4 / 1e9
    n_pos, M = corpus.n_pos, n_emb
    evals_per_sweep_local = float(M) * args.K

    # ---- warm-up + timed region (device time, max over ranks)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        sweep.sweep()
    barrier()
    launches0 = lib.segb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sweep.mma is not None:
        sweep.mma.timing = []              # CUDA events around the scoring kernels of every timed sweep
    ev0.record()
    fallback = 0
    for _ in range(args.steps):
        sweep.sweep()
        fallback += sweep.last_fallback
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    in_sweep_ms = None
    filter_precision = ("e4m3 first level, fp16 second level" if (sweep.mma is not None and sweep.mma.fp8) else "fp16")   # after "auto" settled
    if sweep.mma is not None:
        tm, sweep.mma.timing = sweep.mma.timing, None
        in_sweep_ms = (sum(e[0].elapsed_time(e[1]) for e in tm) / len(tm), sum(e[1].elapsed_time(e[2]) for e in tm) / len(tm))
    launches = lib.segb_launch_count() - launches0
    tot = torch.tensor([evals_per_sweep_local, float(M), float(fallback)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    elapsed_ms = max_over_ranks(elapsed_ms)
    ms_per_step = elapsed_ms / args.steps
    value = args.utts / (ms_per_step * 1e-3)
    evals_per_s = float(tot[0].item()) / (ms_per_step * 1e-3)

    # ---- multi-GPU parity gates (every N; trivially true at N = 1)
    parity = {}
    try:
        parity = multi_gpu_parity(args, world, rank, dev, sweep, comps)
    except Exception as exc:
        parity = {"error": repr(exc)}

    # ---- dominant kernel alone: tcgen05 filter GEMM (tensor roofline) and the DP kernel (HBM roofline)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    peak_bw = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json, bf16 dense burst)" if "bf16_tflops" in peaks else "fallback"
    roofline, roofline_dp = None, None
    phases = sweep.profile_phases()
    default_cfg = (world == 1 and args.utts == TOTAL_UTTS and args.K == K_MAX)
    if rank == 0:
        sp = _lib.stream_ptr()
        if args.scorer == "mma":
            fused = sweep.mma.fused
            run_k = (lambda: sweep.mma.fused_score(sweep.best_val, sweep.best_k)) if fused else sweep.mma.filter
            run_k()
            k_ms_b2b = timed(run_k, 5)                    # five launches back to back: 100 ms of pure tensor work
            k_ms = in_sweep_ms[0]                         # the launches of the timed sweeps themselves
            flops = 2.0 * D * M * args.K                  # algorithmic: 2*D per segment x component eval
            ach = flops / (k_ms * 1e-3) / 1e12
            peak_sus = float(peaks.get("bf16_tflops_sustained", 0.0)) or None
            fp8 = bool(sweep.mma.fp8)
            # r2_ncu_traffic.json: the default (e4m3) filter under its plain name, the fp16 one keyed by its capture file
            kname = "score_fused_kernel" if fused else ("kmeans_filter_kernel" if fp8 else "kmeans_filter_kernel@r2_raw_kmeans16.csv")
            tr = ncu_traffic(kname) if default_cfg else None
            pk_tf, pk_src = peak_tf, peak_src
            if fp8:       # e4m3 dense rate = 2 x the bf16 rate; MEASURED_PEAKS.json has no fp8 entry
                pk_tf, pk_src, peak_sus = 2.0 * peak_tf, "2 x " + peak_src + " (e4m3 runs at twice the bf16 rate; no measured fp8 entry)", (2.0 * peak_sus if peak_sus else None)
            roofline = {"kernel": ("score_fused_kernel<kmeans> (fp32 rows -> fp16 operand tiles in shared memory, tcgen05 fp16 -> fp32 "
                                   "TMEM, top-3 epilogue, exact float32 refine of the survivors: the whole scoring step)" if fused
                                   else "kmeans_filter_kernel<F8> (tcgen05 kind::f8f6f4 e4m3 -> fp32 TMEM, fused top-3 epilogue; first level of "
                                        "the e4m3 -> fp16 -> exact cascade)" if fp8
                                   else "kmeans_filter_kernel (tcgen05 fp16 -> fp32 TMEM, fused top-3 epilogue)"),
                        "bound": "tensor", "achieved": ach, "peak": pk_tf, "unit": "TFLOP/s",
                        "frac": ach / pk_tf,
                        "traffic": tr["bytes_per_launch"] if tr else None,
                        "traffic_source": tr.get("source") if tr else None,
                        "algorithmic_bytes_per_launch": float(X.numel() * 4 + 8 * M) if fused else
                                                        float(sweep.mma.x_tiles.numel() + sweep.mma.cand.numel()),
                        "peak_source": pk_src,
                        "kernel_ms": k_ms, "algorithmic_flops_per_launch": flops,
                        "timing": "kernel_ms = mean over the launches INSIDE the %d timed sweeps (CUDA events on the launching "
                                  "stream around the kernel); kernel_ms_back_to_back = 5 launches of the kernel alone with "
                                  "nothing between them (the power-capped clock of a pure tensor loop)" % args.steps,
                        "kernel_ms_back_to_back": k_ms_b2b, "frac_back_to_back": flops / (k_ms_b2b * 1e-3) / 1e12 / pk_tf,
                        "frac_of_sustained_peak": (ach / peak_sus) if peak_sus else None,
                        "frac_of_measured_bf16_peak": ach / peak_tf,      # e4m3 first level: algorithmic rate against the bf16 figure
                        "filter_precision": "e4m3" if fp8 else "fp16",
                        "refine_ms_in_sweep": in_sweep_ms[1]}
        cs = corpus.struct()

        def time_dp(mode, reps=20, u=None):
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
            torch.cuda.synchronize()
            for a, b_ in evs:
                a.record()
                _lib.check(lib.segb_dp_banded(cs, 0, corpus.n_utt, _lib.ptr(sweep.scores), mode, 0.0,
                                              1.0, _lib.ptr(u), None, _lib.ptr(corpus.bounds), _lib.ptr(sweep.log_prob), None,
                                              None, _lib.ptr(sweep.status), sp))
                b_.record()
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b_) for a, b_ in evs)
            return ts[len(ts) // 2], ts[0]
        # the launch the sweep makes: embeddings verified finite -> SEGB_DP_SCORES_FINITE (no NaN compares)
        dp_mode = _lib.DP_VITERBI_KMEANS | (_lib.DP_SCORES_FINITE if sweep.scores_finite else 0)
        dp_ms_load, _ = time_dp(dp_mode)
        time.sleep(1.0)
        dp_ms, dp_ms_best = time_dp(dp_mode)
        dp_ms_checked, _ = time_dp(_lib.DP_VITERBI_KMEANS)
        u_dp = torch.rand(n_pos, dtype=torch.float64, device=dev)
        ffbs_ms, _ = time_dp(_lib.DP_FFBS, reps=5, u=u_dp)
        sweep.segment()                                        # leave Viterbi boundaries behind
        dp_bytes = 8.0 * n_pos * S_MAX + n_pos + 8.0 * corpus.n_utt + 8.0 * (corpus.n_utt + 1) + 4.0 * corpus.n_utt
        tr = ncu_traffic("dp_staged_kernel") if default_cfg else None
        roofline_dp = {"kernel": "dp_staged_kernel (Viterbi, float64 banded scores, cp.async.bulk staging, thread per utterance)", "bound": "hbm",
                       "achieved": dp_bytes / (dp_ms * 1e-3) / 1e9, "peak": peak_bw, "unit": "GB/s",
                       "frac": dp_bytes / (dp_ms * 1e-3) / 1e9 / peak_bw,
                       "frac_under_load": dp_bytes / (dp_ms_load * 1e-3) / 1e9 / peak_bw,
                       "scores_finite_flag": bool(sweep.scores_finite), "kernel_ms_with_nan_checks": dp_ms_checked,
                       "traffic": tr["bytes_per_launch"] if tr else None,
                       "traffic_source": tr.get("source") if tr else None,
                       "kernel_ms": dp_ms, "kernel_ms_best": dp_ms_best, "kernel_ms_under_load": dp_ms_load,
                       "ffbs_kernel_ms": ffbs_ms, "ffbs_frac": (dp_bytes + 8.0 * n_pos) / (ffbs_ms * 1e-3) / 1e9 / peak_bw,
                       "timing": "median of 20 individually timed launches after a 1 s pause; under_load = same, "
                                 "immediately after the back-to-back filter launches (power-capped SM clock); "
                                 "ffbs = the batched forward-filter backward-sample variant (exp/log bound)",
                       "algorithmic_bytes_per_launch": dp_bytes}

    # ---- the other scoring kernel: FBGMM log_marg_i, ONE fp16 tcgen05 pass + exact float64 refine
    roofline_fv = None
    if rank == 0 and args.scorer == "mma":
        try:
            roofline_fv = fv_logmarg_roofline(args, X, Z, M, peak_tf, peak_src, timed)
        except Exception as exc:
            roofline_fv = {"error": repr(exc)}

    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed sweeps + kernel-alone timings (100 ms samples; sm_mhz_min = the power-capped clock during the GEMM launches)"

    # ---- end to end: host buffers in, host results out, every step
    e2e = None
    if not args.no_e2e:
        X_host = torch.empty(X.shape, dtype=torch.float32, pin_memory=True)
        X_host.copy_(X)
        means_host = torch.empty(comps._means.shape, dtype=torch.float32, pin_memory=True)
        means_host.copy_(comps._means)
        bounds_host = torch.empty(n_pos, dtype=torch.uint8, pin_memory=True)
        assign_host = torch.empty(M, dtype=torch.int32, pin_memory=True)
        total_host = []

        def e2e_step():
            comps._means.copy_(means_host, non_blocking=True)               # H2D model
            comps._meansT.copy_(comps._means.t())
            if args.scorer == "mma":
                # H2D embeddings in chunks, overlapped with fp16 tile packing + filter + refine
                total_host.append(sweep.sweep(X_host=X_host))
            else:
                comps._X.copy_(X_host, non_blocking=True)
                total_host.append(sweep.sweep())
            bounds_host.copy_(corpus.bounds, non_blocking=True)             # D2H segmentation
            assign_host.copy_(comps._assign, non_blocking=True)             # D2H assignments
            means_host.copy_(comps._means, non_blocking=True)               # D2H model
            torch.cuda.synchronize()
        e2e_step()
        barrier()
        n_e2e = max(2, min(args.steps, 3))
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(n_e2e):
            e2e_step()
        ev1.record()
        barrier()
        wall = (time.perf_counter() - t0) / n_e2e
        dev_ms = ev0.elapsed_time(ev1) / n_e2e
        te = max_over_ranks(max(wall * 1e3, dev_ms))
        h2d = X_host.numel() * 4 + means_host.numel() * 4
        d2h = bounds_host.numel() + assign_host.numel() * 4 + means_host.numel() * 4 + 8
        # the platform's host->device limit with every rank copying at once: the same pinned buffer, the same 1M-row
        # chunks, nothing else on the GPU (what the end-to-end step can at best hide its compute behind)
        barrier()
        ev0.record()
        for lo in range(0, X_host.shape[0], 1 << 20):
            comps._X[lo:lo + (1 << 20)].copy_(X_host[lo:lo + (1 << 20)], non_blocking=True)
        ev1.record()
        barrier()
        copy_ms = max_over_ranks(ev0.elapsed_time(ev1))
        e2e = {"value": args.utts / (te * 1e-3), "unit": "utt/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": te, "numa": numa,
               "h2d_copy_only_ms": copy_ms,
               "h2d_copy_only_gbs_all_ranks": float(D) * 4 * float(tot[1].item()) / (copy_ms * 1e-3) / 1e9,
               "note": "per rank: pinned-host X (1M-row chunks on a copy stream, overlapped with fp16 tile packing + filter + refine) + means -> HBM, sweep, boundaries/assignments/means back; each rank bound to its GPU's CPUs (numa)"}
        del X_host

    # ---- CPU baseline + parity gate on the CPU-reproducible head of rank 0's shard
    cpu_baseline = None
    if rank == 0 and not args.no_cpu and n_head:
        kind, ns = cpu_modules()
        n_s = min(args.cpu_sample, CPU_SAMPLE_UTTS, corpus.n_utt)
        hi = int(corpus.pos_off_h[n_s])
        ids = seg_id[:hi]
        e_hi = int(ids.max()) + 1
        means_now = comps._means.cpu().numpy()
        sweep.score()                                   # GPU answers for the same utterances, same means
        sweep.segment()
        torch.cuda.synchronize()
        gpu_bounds = corpus.bounds[:hi].cpu().numpy().astype(bool)
        gpu_tot = sweep.log_prob[:n_s].cpu().numpy()
        gpu_k = sweep.best_k[:e_hi].cpu().numpy()
        seg = _cpu_segmenter(ns, X[:e_hi].cpu().numpy(), lengths[:n_s], ids, seg_dur[:hi], means_now)
        t0 = time.perf_counter()
        totals, bounds, ks = _cpu_phase1(ns, seg)
        dt = time.perf_counter() - t0
        parity_cpu = bool(np.array_equal(np.concatenate(bounds), gpu_bounds) and np.array_equal(np.asarray(totals), gpu_tot))
        for u in range(n_s):
            emb = seg.utterances.get_segmented_embeds_i(u)
            parity_cpu = parity_cpu and [int(gpu_k[e]) for e in emb] == ks[u]
        cpu_baseline = {"value": n_s / dt, "unit": "utt/s", "cores": 1, "kind": kind,
                        "sample": "the first %d utterances (%d candidate segments) of rank 0's shard vs the full K=%d model; %s"
                                  % (n_s, int((ids >= 0).sum()), args.K,
                                     "the reference's own functions (baseline/_ref)" if kind == "reference"
                                     else "oracle port of the reference's pure functions"),
                        "seconds": dt, "parity_with_gpu_on_sample": parity_cpu}

    # ---- sharded FBGMM sweep (frozen-model UnigramAcousticWordseg mode): every N
    fbgmm_frozen = None
    if not args.no_fbgmm:
        try:
            fbgmm_frozen = fbgmm_frozen_secondary(args, world, rank, dev, X, Z, corpus, lengths, seg_id, seg_dur,
                                                  n_head, peak_tf, barrier, max_over_ranks)
        except Exception as exc:
            fbgmm_frozen = {"error": repr(exc)}
    diffuse = None
    if not args.no_diffuse:
        try:
            diffuse = kmeans_diffuse_secondary(args, world, rank, dev, barrier, max_over_ranks)
        except Exception as exc:
            diffuse = {"error": repr(exc)}

    gibbs, diag_x, bigram_x, ingest = None, None, None, None
    if rank == 0 and world == 1 and not args.no_ingest:
        try:
            del sweep, comps, X, Z, corpus
            torch.cuda.empty_cache()
            ingest = ingestion_secondary(args, dev)
        except Exception as exc:
            ingest = {"error": repr(exc)}
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_gibbs:
        torch.cuda.empty_cache()
        try:
            bigram_x = run_bigram_extra(args)
        except Exception as exc:
            bigram_x = {"error": repr(exc)}
        try:
            gibbs = run_gibbs_extra(args)
        except Exception as exc:                      # the headline line must still be printed
            gibbs = {"error": repr(exc)}
        try:
            diag_x = run_diag_extra(args)
        except Exception as exc:
            diag_x = {"error": repr(exc)}
    if rank == 0:
        if roofline is not None:
            # the driver keeps `roofline`: the other kernels' rooflines ride inside it
            roofline["others"] = [r for r in (roofline_dp, roofline_fv) if r]
            roofline["fallback_rows_per_sweep"] = float(tot[2].item()) / args.steps
        line = {
            "metric": METRIC, "value": value, "unit": "utt/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "utterances": args.utts, "K": args.K, "D": D, "max_span": S_MAX,
                       "candidate_segments": int(tot[1].item()), "scorer": args.scorer,
                       "init": "tokens start in the component of their generating cluster (K_act = K_max)",
                       "parallelism": "utterance shards x%d + NCCL all-reduce(sum_x, counts)" % world,
                       "fused_scorer": bool(args.scorer == "mma" and sweep_fused),
                       "filter_precision": filter_precision,
                       "l2": "inputs (%.1f GB of embeddings per rank) exceed L2; no flush needed" % x_gb},
            "segment_component_evals_per_s": evals_per_s,
            "fallback_rows_per_sweep": float(tot[2].item()) / args.steps,
            "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roofline,
            "roofline_dp": roofline_dp, "roofline_fixedvar_logmarg": roofline_fv, "cpu_baseline": cpu_baseline,
            "phases_ms": phases, "parity": parity,
            "secondary_fbgmm_frozen": fbgmm_frozen, "secondary_kmeans_diffuse": diffuse, "ingestion": ingest,
            "secondary_gibbs_fixedvar": gibbs, "secondary_gibbs_diag": diag_x, "secondary_bigram": bigram_x,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.only_diffuse:
        import torch
        torch.cuda.set_device(0)

        def _sync():
            torch.cuda.synchronize()
        print(json.dumps({"secondary_kmeans_diffuse": kmeans_diffuse_secondary(
            args, 1, 0, torch.device("cuda", 0), _sync, lambda v: float(v))}))
    elif args.only_gibbs:
        print(json.dumps({"secondary_gibbs_fixedvar": run_gibbs_extra(args)}))
    elif args.only_diag:
        print(json.dumps({"secondary_gibbs_diag": run_diag_extra(args)}))
    elif args.only_bigram:
        print(json.dumps({"secondary_bigram": run_bigram_extra(args)}))
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
