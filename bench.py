#!/usr/bin/env python
"""bench.py -- plmDCA gradient-evaluations/second on the BASELINE workload, plus the other
half of BASELINE's metric (mfDCA end to end) and the large configurations as extra blocks.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Headline workload (BASELINE.json configs[1], "C2"): plmdca compute_fn protein on the seeded
synthetic MSA N=5000, L=150, q=21, seqid 0.8, lambda_h = lambda_J = 0.2(L-1),
max_iterations = 500.

A STEP is one complete plmDCA fit of that MSA: L-BFGS from the reference's initial
point until the optimiser stops (as in the reference, the float32 More-Thuente line
search ends the run long before 500 iterations, SURVEY finding 4).  The metric counts
every objective/gradient evaluation inside the L-BFGS loop (1 + sum of line-search
trials); `value` = evaluations of the K timed fits / their CUDA-event time with the MSA,
weights and work buffers resident in HBM; `e2e` = the same metric through the
reference's own entry point plmdcaBackend(file, ...) with host buffers: FASTA parse,
de-duplication, reweighting, H2D, fit and D2H of the result inside the timed region.

Extra blocks of the N=1 line (each measured live in this run; --no-extras skips them):
  parity    the fit against the frozen reference fit (tests/golden/plm_c2_fit.npz) in both
            arithmetic modes: evaluation count, max-rel FN deviation, top-L overlap, next to the
            reference's own nt=1 vs nt=all-cores band
  mfdca     BASELINE configs[2] (N=20000, L=300): MeanFieldDCA(file).compute_sorted_FN_APC() end
            to end incl. FASTA ingest and reweighting, site-pairs/s, per-stage CUDA-event ms, the
            inverse against the MEASURED FP64 tensor-pipe peak, a CPU baseline of the same stages
  reweight  BASELINE configs[4]'s all-pairs Hamming reweighting (N=200000, L=250) on one GPU
  c4        BASELINE configs[3] (N=50000, L=500): a 10-iteration fit on one GPU

At N > 1 GPUs the SAME C2 fit is owner-sharded by sites (pydca_b200.distributed.ShardedPlm):
every rank owns a site range, the matching slice of the parameter vector and of the L-BFGS
vectors; per evaluation only the cross blocks move, by P2P stores over NVLink fused into
the unpack / export kernels, i.e. strong scaling.  The gathered result is compared with a
single-GPU fit of the same input in the same run (`multi_gpu_check`).

--impl reference times the reference's own OpenMP implementation (oracle/_ref, built
from /root/reference unmodified) on the host cores: each step is a bounded sample
(the first REF_ITERS L-BFGS iterations of the same fit).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SEQ, SEQ_LEN, Q = 5000, 150, 21
SEQID = 0.8
MAX_ITER = 500
REF_ITERS = 3            # L-BFGS iterations per reference-arm step (bounded sample)
CPU_BASELINE_ITERS = 8   # bounded sample for the cpu_baseline object of our own line


def workload(tmpdir):
    from pydca_b200.synthetic import synthetic_msa, write_fasta
    msa = synthetic_msa(N_SEQ, SEQ_LEN, Q)
    path = os.path.join(tmpdir, "c2_N%d_L%d.fa" % (N_SEQ, SEQ_LEN))
    if not os.path.exists(path):
        write_fasta(msa, path, Q)
    lam = 0.2 * (SEQ_LEN - 1)
    return msa, path, lam


def parallelism(n_gpus, exchange=None):
    if n_gpus > 1 and exchange == "peer":
        return ("owner-sharded sites x%d: P2P stores of the cross blocks (fused unpack/export kernels) + "
                "flag barriers + rank-ordered float64 scalar sums over peer memory; no NCCL in the loop" % n_gpus)
    if n_gpus > 1:
        return "site-shard x%d + NCCL all-reduce(g) per evaluation" % n_gpus
    return "single GPU"


def config_dict():
    """The workload only (identical at every GPU count; how it is parallelised is reported
    next to it as `parallelism`)."""
    return {"workload": "C2: plmdca compute_fn protein, synthetic MSA N=%d L=%d q=%d, seqid %.1f, "
                        "lambda_h=lambda_J=0.2(L-1), max_iterations=%d; one step = one full fit"
                        % (N_SEQ, SEQ_LEN, Q, SEQID, MAX_ITER),
            "N": N_SEQ, "L": SEQ_LEN, "q": Q, "max_iterations": MAX_ITER,
            "l2": "working set per evaluation (W 40 MB + S 64 MB + G >= 40 MB + 16 L-BFGS vectors x 20 MB) "
                  "exceeds the 126 MB L2; no explicit flush"}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return out


def cpu_reference_fit(path, lam, iters, threads):
    """The reference's own lbfgs + PlmDCA::gradient (oracle/_ref) or, if that binary is
    absent, single gradient calls of the C port.  Returns (evals, seconds, kind, cores)."""
    import oracle
    if oracle.have_ref():
        R = oracle.ref()
        rp = R.plm(path, "protein", SEQ_LEN, Q, SEQID, lam, lam, threads)   # parse + dedup + reweight (untimed)
        t0 = time.perf_counter()
        _, status, evals, fx = rp.fit(iters)
        dt = time.perf_counter() - t0
        rp.close()
        return evals, dt, "reference", threads
    P = oracle.port()
    msa = P.read_fasta(path, "protein", SEQ_LEN)
    w = P.plm_weights(P.reweight_counts(msa, SEQID))
    x = P.plm_init(msa, w, Q)
    P.set_num_threads(threads)
    t0 = time.perf_counter()
    n = max(1, iters)
    for _ in range(n):
        P.plm_gradient(msa, w, Q, lam, lam, x)
    return n, time.perf_counter() - t0, "port", threads


def run_reference(args, rank, world):
    if rank != 0:
        return
    tmp = tempfile.mkdtemp(prefix="dca_bench_")
    msa, path, lam = workload(tmp)
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_reference_fit(path, lam, REF_ITERS, cores)
    evals = 0
    secs = 0.0
    kind = "reference"
    for _ in range(args.steps):
        e, dt, kind, _ = cpu_reference_fit(path, lam, REF_ITERS, cores)
        evals += e
        secs += dt
    value = evals / secs
    sample = ("each step = first %d L-BFGS iterations of the C2 fit (%s lbfgs + PlmDCA::gradient, %d OpenMP threads); "
              "parse/dedup/reweight untimed" % (REF_ITERS, "reference" if kind == "reference" else "oracle port", cores))
    line = {"impl": "reference", "metric": "plmdca_grad_evals_per_sec", "value": value, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(),
            "parallelism": "host cores (OpenMP)",
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def find_hbm_peak(peaks):
    """(GB/s, source) from MEASURED_PEAKS.json -- the driver's key names are not fixed, so take the
    first numeric entry whose key looks like an HBM / copy bandwidth (burst figure preferred: these
    kernels are timed alone); else the fallback of /opt/skills/guides/B200_PROFILING.md."""
    found = []

    def walk(node, prefix):
        if isinstance(node, dict):
            for k, v in node.items():
                walk(v, (prefix + "." + str(k)) if prefix else str(k))
        elif isinstance(node, (int, float)) and not isinstance(node, bool):
            key = prefix.lower()
            if any(t in key for t in ("nvlink", "pcie", "l2", "smem", "shared", "h2d", "d2h", "tf", "flop")):
                return
            for rank, tag in enumerate(("hbm", "dram", "copy", "mem_bw", "membw")):
                if tag in key:
                    found.append((prefix, float(node), rank))
                    break
    walk(peaks, "")
    found = [(k, v if v < 1e5 else v / 1e9, r) for k, v, r in found if v > 0]       # bytes/s -> GB/s
    found = [(k, v, r) for k, v, r in found if 1000.0 < v < 20000.0]
    if not found:
        return 6650.0, "fallback 6650 GB/s (no MEASURED_PEAKS.json HBM entry)"
    found.sort(key=lambda kv: (kv[2], 0 if "burst" in kv[0].lower() else 1 if "sustain" not in kv[0].lower() else 2))
    return found[0][1], "MEASURED_PEAKS.json %s" % found[0][0]


def roofline_block(obj, x0, peaks):
    """CUDA-event time of every kernel of one evaluation (events on the stream the
    kernels run on), dominant kernel vs the HBM roofline, plus the two on-chip roofs
    that actually bound it."""
    reps = 5
    obj.eval_timed(x0)
    acc = None
    for _ in range(reps):
        t = obj.eval_timed(x0)
        acc = t if acc is None else {k: acc[k] + t[k] for k in t}
    ms = {k: v / reps for k, v in acc.items()}
    ms.pop("export", None)                 # owner-sharded runs only (no kernel here: an empty event gap)
    g = obj.describe()
    N, L, q, ld, P, nsplit = g["N"], g["L"], g["q"], g["ld"], g["P"], g["nsplit"]
    Lq = L * q
    # algorithmic (compulsory) bytes of each kernel, one launch
    bytes_k = {
        "unpack": 4 * P + 4 * Lq * Lq,
        "logits": 4 * Lq * Lq + N * L + 4 * N * Lq,            # read W once, MSA once; write S
        "softmax_scan": 8 * N * Lq + N * L + 4 * N,            # read S, write r
        "scatter": 4 * N * Lq + int(1.75 * N * L) + 4 * Lq * Lq * nsplit,
        "assemble": 4 * Lq * Lq * nsplit + 8 * P,
    }
    adds = N * L * Lq                                          # gather adds of logits == scatter
    top = max(ms, key=ms.get)
    hbm, hbm_source = find_hbm_peak(peaks)
    eval_bytes = 8 * P + N * L + 4 * N                          # SURVEY 8(d) per-evaluation figure
    total_ms = sum(ms.values())
    sm_clock_ghz = 1965.0 / 1e3                                # clocks.max.sm of the B200 (bench records the live clock too)
    lane_peak = 148 * 128 * sm_clock_ghz                       # Gop/s FP32 lanes
    smem_peak = 148 * 32 * sm_clock_ghz                        # G words/s shared-memory bandwidth
    # DRAM traffic per launch of the same kernel from the committed `ncu --set full` capture
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        ncu_name = {"unpack": "plm_unpack_kernel<21>", "logits": "plm_logits_kernel<21, 128>",
                    "softmax_scan": "plm_softmax_kernel<21, 1, 1>", "scatter": "plm_scatter_kernel<21>",
                    "assemble": "plm_assemble_kernel<21>"}[top]
        traffic = tj["kernels"][ncu_name]["dram_bytes"]
        traffic_src = tj["source"]
    except Exception:
        pass
    # What binds the dominant kernel (ncu, profiles/): the shared-memory pipe -- an add of the gather / scatter
    # kernels needs one 4-byte operand selected by a residue, one LDS word.  On the similarity-sorted layout four rows
    # that share a residue at a site take ONE word (logits: the 128-column kernel; scatter: the four-row group sums),
    # so the words the pipe moves are adds x (1 - 3/4 u), u = fraction of such (site, 4-row group) pairs in this MSA.
    # `achieved` / `frac` are the EXECUTED words against the pipe's peak; the algorithmic adds per second are reported
    # next to them (they can exceed the one-word-per-add roof, which is the point of the fast path).
    try:
        stats = obj.layout_stats()
    except Exception:
        stats = {"sorted": False, "group_sums": False, "uniform4": 0.0, "logits_columns": 128}
    u4 = stats["uniform4"]
    words_per_add = {"logits": 1.0 - 0.75 * u4 if stats["logits_columns"] == 128 else 1.0,
                     "scatter": 1.0 - 0.75 * u4 if stats["group_sums"] else 1.0}
    gadds = {"logits": adds / (ms["logits"] * 1e-3) / 1e9, "scatter": adds / (ms["scatter"] * 1e-3) / 1e9}
    gwords = {k: gadds[k] * words_per_add[k] for k in gadds}
    top_gather = top if top in gadds else max(gadds, key=lambda k: ms[k])
    out = {"bound": "smem", "kernel": top_gather,
           "achieved": gadds[top_gather], "peak": smem_peak,
           "unit": "Gadd/s = G operand words/s at one shared-memory word per add (SURVEY 8d counts the adds)",
           "frac": gadds[top_gather] / smem_peak,
           "note": "algorithmic gather adds per launch (N*L*L*q) over the kernel's CUDA-event time, against the "
                   "shared-memory pipe at ONE operand word per add -- the roof of the dense kernels.  The equal-residue "
                   "fast path feeds four adds from one word where four neighbouring rows agree, so this fraction is no "
                   "longer bounded by 1; `executed` is what the pipe really moves, `fp32_lane` the hard ceiling",
           "executed": {"words_per_add": words_per_add[top_gather], "gwords_per_s": gwords[top_gather],
                        "frac_of_smem_pipe": gwords[top_gather] / smem_peak,
                        "note": "adds x (1 - 3/4 uniform4): the operand words the kernel loads (ncu: shared pipe 74 % busy "
                                "for the scatter, 72 % for the logits kernel -- profiles/r2l_eval_ncu_full_key_metrics.csv)"},
           "layout": stats,
           "traffic": traffic, "traffic_source": traffic_src,
           "peak_source": "148 SMs x 32 LDS words/clk x %.3f GHz (clocks.max.sm)" % sm_clock_ghz,
           "fp32_lane": {"note": "same kernel against the FP32-lane roof (148 x 128 lanes)",
                         "achieved_gops": gadds[top_gather], "peak_gops": lane_peak, "frac": gadds[top_gather] / lane_peak},
           "hbm": {"note": "secondary: SURVEY 8(d) algorithmic bytes of one evaluation (8P + N*L + 4N) over the "
                           "CUDA-event time of its kernels vs the measured HBM peak; small by construction "
                           "(intensity ~117 adds/byte vs a ridge of ~6)",
                   "algorithmic_bytes": eval_bytes, "ms": total_ms,
                   "achieved": eval_bytes / (total_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                   "frac": eval_bytes / (total_ms * 1e-3) / 1e9 / hbm, "peak_source": hbm_source},
           "kernel_ms": ms, "kernel_share": {k: v / total_ms for k, v in ms.items()},
           "kernel_algorithmic_bytes": bytes_k,
           "on_chip": {"gather_adds_per_kernel": adds,
                       "logits_gadds_per_s": gadds["logits"], "scatter_gadds_per_s": gadds["scatter"],
                       "logits_words_per_add": words_per_add["logits"], "scatter_words_per_add": words_per_add["scatter"],
                       "smem_roof_gwords_per_s": smem_peak, "fp32_lane_roof_gops": lane_peak,
                       "logits_frac_of_one_word_per_add_roof": gadds["logits"] / smem_peak,
                       "scatter_frac_of_one_word_per_add_roof": gadds["scatter"] / smem_peak,
                       "logits_executed_frac_of_smem_pipe": gwords["logits"] / smem_peak,
                       "scatter_executed_frac_of_smem_pipe": gwords["scatter"] / smem_peak,
                       "eval_frac_of_fp32_lane_roof": 2 * adds / (total_ms * 1e-3) / 1e9 / lane_peak}}
    # the HBM-bound part of the path: the two L-BFGS vector passes of one iteration (m = 5)
    try:
        from pydca_b200 import _lib as _l
        vm = _l.lbfgs_vector_bench(P, 5, 20)
        gram_bytes = 4 * P * (4 + 8 + 2)          # x, xp, g, gp + 4 stored s/y pairs read; s, y written
        dir_bytes = 4 * P * (1 + 10 + 1)          # g + 5 s + 5 y read; d written
        step_bytes = 4 * P * 3
        out["hbm_bound_kernels"] = {
            "note": "L-BFGS vector algebra: algorithmic bytes / CUDA-event time vs the measured HBM peak",
            "vec_update_gram": {"ms": vm["update_gram"], "bytes": gram_bytes,
                                "achieved_gbs": gram_bytes / (vm["update_gram"] * 1e-3) / 1e9,
                                "frac": gram_bytes / (vm["update_gram"] * 1e-3) / 1e9 / hbm},
            "vec_direction": {"ms": vm["direction"], "bytes": dir_bytes,
                              "achieved_gbs": dir_bytes / (vm["direction"] * 1e-3) / 1e9,
                              "frac": dir_bytes / (vm["direction"] * 1e-3) / 1e9 / hbm},
            "vec_step": {"ms": vm["step"], "bytes": step_bytes,
                         "achieved_gbs": step_bytes / (vm["step"] * 1e-3) / 1e9,
                         "frac": step_bytes / (vm["step"] * 1e-3) / 1e9 / hbm}}
    except Exception as exc:
        out["hbm_bound_kernels"] = {"error": repr(exc)}
    return out


def fn_scores(x, L, q):
    from pydca_b200 import _lib
    return _lib.plm_fn_apc(x, L, q)[0]


def parity_block(msa, w, lam):
    """The C2 fit in both arithmetic modes against the frozen reference fit
    (tests/golden/plm_c2_fit.npz: the unmodified reference's lbfgs() + PlmDCA::gradient at
    num_threads = 1 and = all cores, written by tests/golden/make_c2_fit.py)."""
    from pydca_b200 import _lib
    gold = np.load(os.path.join(ROOT, "tests", "golden", "plm_c2_fit.npz"))
    ref = gold["nt1_fn"].astype(np.float64)
    refN = gold["ntN_fn"].astype(np.float64)
    top = np.argsort(-ref, kind="stable")[:SEQ_LEN]
    out = {"golden": "tests/golden/plm_c2_fit.npz (reference at num_threads=1)",
           "reference": {"evaluations": int(gold["nt1_evaluations"]), "status": int(gold["nt1_status"]),
                         "fx": float(gold["nt1_fx"]),
                         "self_band_fn_max_rel": float(np.max(np.abs(refN - ref) / ref)),
                         "self_band_note": "reference num_threads=1 vs num_threads=%d" % int(gold["ntN_threads"])},
           "tolerance": "north_star: 1e-5 relative on FN scores"}
    obj = _lib.PlmObjective(msa, w, Q, lam, lam)
    try:
        for mode in ("reference", "accurate"):
            obj.set_arithmetic(mode)
            x0 = obj.init()
            obj.fit(x0, MAX_ITER)                                   # warm
            x, res = obj.fit(x0, MAX_ITER)
            fn = fn_scores(x, SEQ_LEN, Q)
            out["arithmetic_" + mode] = {
                "evaluations": res.evaluations, "iterations": res.iterations, "status": res.status, "fx": res.fx,
                "fn_max_rel_dev": float(np.max(np.abs(fn - ref) / ref)),
                "topL_overlap": int(len(set(top.tolist()) & set(np.argsort(-fn, kind="stable")[:SEQ_LEN].tolist()))),
                "topL": SEQ_LEN, "evals_per_s": res.evaluations / (res.device_ms * 1e-3), "fit_ms": res.device_ms}
    finally:
        obj.close()
    out["explanation"] = ("the reference sums its inner products and the L2 term sequentially in float32 over all P; "
                          "that bias (1.8 % of |g|^2 here) steers its L-BFGS trajectory. arithmetic=reference "
                          "reproduces those sums bit for bit (csrc/seqsum.cu); arithmetic=accurate (default, the "
                          "headline) uses float64 reductions and lands where the reference itself lands when "
                          "built with float64 sums (DESIGN.md section 6)")
    return out


def mfdca_block(tmp, peaks):
    """BASELINE configs[2]: mfdca compute_fn protein, N=20000 L=300 (n = 6000), pseudocount 0.5,
    through the class a user calls, FASTA file in, sorted FN-APC list out."""
    from pydca_b200 import _lib, msa_numerics
    from pydca_b200.meanfield_dca import MeanFieldDCA
    from pydca_b200.synthetic import synthetic_msa, write_fasta
    N, L, pc = 20000, 300, 0.5
    n = L * (Q - 1)
    npairs = L * (L - 1) // 2
    msa = synthetic_msa(N, L, Q)
    path = write_fasta(msa, os.path.join(tmp, "c3_N%d_L%d.fa" % (N, L)), Q)
    MeanFieldDCA(path, "protein", pseudocount=pc, seqid=SEQID).compute_sorted_FN_APC()      # warm
    reps, walls, stages, t_ctor = 3, [], None, []
    for _ in range(reps):
        t0 = time.perf_counter()
        inst = MeanFieldDCA(path, "protein", pseudocount=pc, seqid=SEQID)
        t1 = time.perf_counter()
        scores = inst.compute_sorted_FN_APC()
        walls.append(time.perf_counter() - t0)
        t_ctor.append(t1 - t0)
        stages = inst.last_stage_ms
    assert len(scores) == npairs
    wall = float(np.median(walls))
    # device-resident part alone (MSA + weights already on the host as arrays)
    counts = _lib.sequence_counts(msa, SEQID, compare_f64=True)
    w = 1.0 / counts.astype(np.float64)
    res = msa_numerics.mean_field_scores(msa, Q, w, pc)
    res = msa_numerics.mean_field_scores(msa, Q, w, pc)
    st = {k: float(v) for k, v in zip(("pair_freqs_and_h2d", "covariance", "inverse", "frobenius_norm", "direct_info"),
                                      res["stage_ms"])}
    dev_s = 1e-3 * sum(st.values())
    peak = _lib.fp64_mma_peak(5)
    inv_tflops = float(n) ** 3 / (st["inverse"] * 1e-3) / 1e12
    hbm, hbm_src = find_hbm_peak(peaks)
    hist_bytes = N * L + 8 * N + 8 * n * n
    out = {"workload": "C3: mfdca compute_fn protein, synthetic MSA N=%d L=%d q=%d, pseudocount %.1f, seqid %.1f "
                       "(n = %d covariance)" % (N, L, Q, pc, SEQID, n),
           "e2e_seconds": wall, "pairs_per_s": npairs / wall, "site_pairs": npairs,
           "e2e_call": "MeanFieldDCA(msa_file, 'protein', pseudocount, seqid).compute_sorted_FN_APC(): FASTA parse + "
                       "dedup + reweighting + H2D + frequencies + covariance + inverse + FN + D2H (couplings kept, "
                       "as compute_params needs them) + APC + sort",
           "constructor_seconds": float(np.median(t_ctor)), "stage_ms_in_call": stages,
           "device_resident": {"seconds": dev_s, "pairs_per_s": npairs / dev_s, "stage_ms": st,
                               "call": "dca_mf_scores on host arrays (weights given), scores only"},
           "inverse": {"flops": float(n) ** 3, "note": "symmetric blocked Gauss-Jordan: n^3 flops (Cholesky+potri count)",
                       "achieved_tflops": inv_tflops, "peak_tflops": peak,
                       "peak_source": "measured in this run: register-only mma.sync.m8n8k4.f64 stream on all SMs "
                                      "(dca_fp64_mma_peak)", "frac": inv_tflops / peak},
           "pair_histogram": {"algorithmic_bytes": hist_bytes, "ms": st["pair_freqs_and_h2d"],
                              "achieved_gbs": hist_bytes / (st["pair_freqs_and_h2d"] * 1e-3) / 1e9, "peak_gbs": hbm,
                              "frac": hist_bytes / (st["pair_freqs_and_h2d"] * 1e-3) / 1e9 / hbm,
                              "note": "SURVEY 8(d): N*L + 8N + 8n^2 bytes; stage time includes the H2D of the MSA"}}
    # CPU baseline of the same stages: numpy restatement of the reference's numba functions
    # (oracle/mf_oracle.py) on a row subsample for the O(N) stages, the full 6000^2 inverse
    try:
        from oracle import mf_oracle
        import oracle
        nsub = 2000
        t0 = time.perf_counter()
        c_sub = oracle.port().reweight_counts_rows(msa, SEQID, np.arange(nsub, dtype=np.int64), f64=True)
        t_rw = (time.perf_counter() - t0) * (N / nsub)
        wsub = 1.0 / c_sub.astype(np.float64)
        t0 = time.perf_counter()
        fi = mf_oracle.reg_single_site_freqs(mf_oracle.single_site_freqs(msa[:nsub], Q, wsub), Q, pc)
        fij = mf_oracle.reg_pair_site_freqs(mf_oracle.pair_site_freqs(msa[:nsub], Q, wsub), Q, pc)
        t_freq = (time.perf_counter() - t0) * (N / nsub)
        t0 = time.perf_counter()
        C = mf_oracle.construct_corr_mat(fi, fij, L, Q)
        t_corr = time.perf_counter() - t0
        t0 = time.perf_counter()
        np.linalg.inv(C)
        t_inv = time.perf_counter() - t0
        total = t_rw + t_freq + t_corr + t_inv
        out["cpu_baseline"] = {"value": total, "unit": "s", "cores": os.cpu_count() or 1, "kind": "port",
                               "pairs_per_s": npairs / total,
                               "stage_s": {"reweight": t_rw, "frequencies": t_freq, "covariance": t_corr, "inverse": t_inv},
                               "sample": "oracle (numpy / C restatement of meanfield_dca/msa_numerics.py): reweighting and "
                                         "frequencies on the first %d of %d sequences, extrapolated linearly in N (the "
                                         "work is exactly linear in rows); covariance assembly and numpy.linalg.inv at "
                                         "full size" % (nsub, N)}
    except Exception as exc:
        out["cpu_baseline"] = {"value": None, "kind": "unavailable", "sample": repr(exc)}
    return out


def reweight_block():
    """All-pairs Hamming reweighting at BASELINE configs[4]'s size on one GPU (symmetric tiles)."""
    from pydca_b200 import _lib
    from pydca_b200.synthetic import synthetic_msa
    out = {}
    for tag, N, L in (("c2", N_SEQ, SEQ_LEN), ("c5", 200000, 250)):
        msa = synthetic_msa(N, L, Q)
        _lib.sequence_counts_part(msa, SEQID, 0, 1)
        ms = min(_lib.sequence_counts_part(msa, SEQID, 0, 1, want_ms=True)[1] for _ in range(3))
        compares = float(N) * N * L
        lane_ops = compares / 4.0
        sm_clock_ghz = 1.965
        floor_ms = lane_ops / (148 * 128 * sm_clock_ghz * 1e9) * 1e3
        out[tag] = {"N": N, "L": L, "device_ms": ms, "byte_compares_per_s": compares / (ms * 1e-3),
                    "packed_lane_floor_ms": floor_ms, "frac_of_packed_lane_floor": floor_ms / ms,
                    "note": "floor = N^2 L / 4 packed lane-ops at 148 x 128 lanes x 1.965 GHz (SURVEY 8d); the kernel "
                            "spends 4 integer ops per 4 residues (xor, add, and, dp4a) on HALF the pairs (symmetric tiles); deep alignments are compared in similarity order and a tile stops once every pair is over the mismatch budget, so the rate is algorithmic (N^2 L byte compares / time, ordering pass included)"}
    return out


def c4_block(hbm_peak=None):
    """BASELINE configs[3] on ONE GPU: N=50000, L=500 (P = 55 M): a 10-iteration fit."""
    from pydca_b200 import _lib
    from pydca_b200.synthetic import synthetic_msa
    N, L = 50000, 500
    lam = 0.2 * (L - 1)
    msa = synthetic_msa(N, L, Q)
    counts = _lib.sequence_counts(msa, SEQID)
    w = (1.0 / counts.astype(np.float32)).astype(np.float32)
    obj = _lib.PlmObjective(msa, w, Q, lam, lam)
    try:
        x0 = obj.init()
        obj.fit(x0, 2)
        _, res = obj.fit(x0, 10)
        ms = obj.eval_timed(x0)
        ms = obj.eval_timed(x0)
        ms.pop("export", None)
        adds = float(N) * L * L * Q
        smem_peak = 148 * 32 * 1.965
        st = obj.layout_stats()
        wl = 1.0 - 0.75 * st["uniform4"] if st["logits_columns"] == 128 else 1.0
        ws = 1.0 - 0.75 * st["uniform4"] if st["group_sums"] else 1.0
        out = {"workload": "C4 on one GPU: plmdca N=%d L=%d q=%d, first 10 L-BFGS iterations" % (N, L, Q),
               "evals_per_s": res.evaluations / (res.device_ms * 1e-3), "evaluations": res.evaluations,
               "fit_ms": res.device_ms, "kernel_ms": ms, "layout": st,
               "logits_gadds_per_s": adds / (ms["logits"] * 1e-3) / 1e9, "scatter_gadds_per_s": adds / (ms["scatter"] * 1e-3) / 1e9,
               "logits_frac_of_one_word_per_add_roof": adds / (ms["logits"] * 1e-3) / 1e9 / smem_peak,
               "scatter_frac_of_one_word_per_add_roof": adds / (ms["scatter"] * 1e-3) / 1e9 / smem_peak,
               "logits_executed_frac_of_smem_pipe": wl * adds / (ms["logits"] * 1e-3) / 1e9 / smem_peak,
               "scatter_executed_frac_of_smem_pipe": ws * adds / (ms["scatter"] * 1e-3) / 1e9 / smem_peak,
               "note": "algorithmic adds against the one-word-per-add roof (can exceed 1 with the equal-residue fast path) and "
                       "the EXECUTED shared-memory words (adds x (1 - 3/4 uniform4)) against the pipe's peak"}
    finally:
        obj.close()
    try:        # the HBM-bound L-BFGS passes at this size (220 MB vectors; C2's are 20 MB: ramp-dominated)
        P = L * Q + L * (L - 1) // 2 * Q * Q
        vm = _lib.lbfgs_vector_bench(P, 5, 5)
        hbm = hbm_peak if hbm_peak else 6552.0
        nbytes = {"update_gram": 4 * P * 14, "direction": 4 * P * 12, "step": 4 * P * 3}
        out["hbm_bound_kernels"] = {k: {"ms": vm[k], "bytes": nbytes[k], "achieved_gbs": nbytes[k] / (vm[k] * 1e-3) / 1e9,
                                        "frac": nbytes[k] / (vm[k] * 1e-3) / 1e9 / hbm} for k in nbytes}
    except Exception as exc:
        out["hbm_bound_kernels"] = {"error": repr(exc)}
    return out


def run_ours(args, rank, world, local_rank):
    from pydca_b200 import _lib
    from pydca_b200 import distributed as dd
    L = _lib.lib()
    if L.dca_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; pydca_b200 has no CPU fallback")
    _lib.check(L.dca_set_device(local_rank), "dca_set_device")
    dist = None
    torch = None
    saved_stdout = None
    cpu_group = None
    if world > 1:
        # stdout must carry exactly ONE line (the JSON): NCCL prints its version banner there when
        # NCCL_DEBUG asks for it, so everything up to the final print goes to stderr instead
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")     # host-side barriers: no kernel parked on idle GPUs
    tmp = tempfile.mkdtemp(prefix="dca_bench_")
    msa, path, lam = workload(tmp)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    # ---- device-resident objective --------------------------------------------------
    sp, obj, cb, exchange = None, None, None, None
    if world > 1:
        dev = torch.device("cuda", local_rank)
        counts = dd.sharded_counts(msa, SEQID, device=dev)
        w = (1.0 / counts.astype(np.float32)).astype(np.float32)
        if os.environ.get("DCA_B200_EXCHANGE", "peer") == "peer":
            try:
                sp = dd.ShardedPlm(msa, w, Q, lam, lam)
                obj, exchange = sp.obj, "peer"
            except Exception as exc:          # every rank raises together (ShardedPlm agrees on failure)
                print("bench.py: peer-memory exchange unavailable (%r); using NCCL all-reduce" % (exc,), file=sys.stderr)
                sp = None
        if sp is None:
            obj = _lib.PlmObjective(msa, w, Q, lam, lam, sites=dd.shard_range(SEQ_LEN, rank, world))
            cb = dd.make_allreduce_callback(obj.cuda_stream())
            exchange = "allreduce"
    else:
        counts = _lib.sequence_counts(msa, SEQID)
        w = (1.0 / counts.astype(np.float32)).astype(np.float32)
        obj = _lib.PlmObjective(msa, w, Q, lam, lam)
    x0 = obj.init()

    def barrier():
        if dist is not None:
            dist.barrier()

    def cpu_barrier():
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier(group=cpu_group)

    for _ in range(args.warmup):
        obj.fit(x0, MAX_ITER, allreduce=cb)

    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = L.dca_kernel_launches()
    evals, dev_ms = 0, 0.0
    t0 = time.perf_counter()
    status = iters = 0
    for _ in range(args.steps):
        _, res = obj.fit(x0, MAX_ITER, allreduce=cb)       # fit() synchronises its stream before returning
        evals += res.evaluations
        dev_ms += res.device_ms
        status, iters = res.status, res.iterations
    wall = time.perf_counter() - t0
    barrier()
    launches = L.dca_kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([dev_ms, wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall = float(t[0]), float(t[1])

    # ---- end to end through the reference's entry point (host buffers, file input) ----
    e2e = None
    if world == 1:
        n_e2e = max(1, min(args.steps, 5))
        _lib.set_backend_gpus(1)              # the box may show more GPUs than this run is about
        _lib.plmdca_backend(path, "protein", SEQ_LEN, Q, SEQID, lam, lam, MAX_ITER)      # warm (page cache, context)
        e_evals = 0
        per_call = []
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            t1 = time.perf_counter()
            _lib.plmdca_backend(path, "protein", SEQ_LEN, Q, SEQID, lam, lam, MAX_ITER)
            e_evals += _lib.last_fit().evaluations
            per_call.append(1e3 * (time.perf_counter() - t1))
        e_wall = time.perf_counter() - t0
        print("e2e per-call ms: " + " ".join("%.1f" % v for v in per_call), file=sys.stderr)
        P = obj.P
        e2e = {"value": e_evals / e_wall, "unit": "evals/s",
               "h2d_bytes_per_step": 2 * N_SEQ * SEQ_LEN + 4 * N_SEQ + 4 * P,
               "d2h_bytes_per_step": 4 * N_SEQ + 8 * P,
               "ms_per_step": 1e3 * e_wall / n_e2e, "steps": n_e2e,
               "call": "plmdcaBackend(msa_file, ...) -> float*: FASTA parse + dedup + reweight + H2D + fit + D2H"}
    else:
        # end to end at N GPUs through the SAME entry point: rank 0 calls plmdcaBackend(file, ...) with
        # DCA_B200_GPUS = N (one host thread per GPU inside the call, peer-memory exchange, every GPU
        # copying its slice into the returned host vector); the other ranks' processes stay idle
        last_x = [None]
        n_e2e = max(1, min(args.steps, 3))
        e_evals, e_wall = 0, 0.0
        cpu_barrier()
        if rank == 0:
            _lib.set_backend_gpus(world)
            _lib.plmdca_backend(path, "protein", SEQ_LEN, Q, SEQID, lam, lam, MAX_ITER)      # warm: contexts, arenas
            per_call = []
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                # the previous result goes back to the library before the next call, as in the single-GPU loop
                # above (holding it made every second call allocate a fresh 20 MB pinned buffer: +12 ms)
                last_x[0] = None
                t1 = time.perf_counter()
                last_x[0] = _lib.plmdca_backend(path, "protein", SEQ_LEN, Q, SEQID, lam, lam, MAX_ITER)
                e_evals += _lib.last_fit().evaluations
                per_call.append(1e3 * (time.perf_counter() - t1))
            e_wall = time.perf_counter() - t0
            print("e2e per-call ms (%d GPUs): " % world + " ".join("%.1f" % v for v in per_call), file=sys.stderr)
        cpu_barrier()
        if rank == 0:
            e2e = {"value": e_evals / e_wall, "unit": "evals/s",
                   "h2d_bytes_per_step": world * (2 * N_SEQ * SEQ_LEN + 4 * N_SEQ),
                   "d2h_bytes_per_step": 4 * N_SEQ + 4 * obj.P, "ms_per_step": 1e3 * e_wall / n_e2e, "steps": n_e2e,
                   "call": "plmdcaBackend(msa_file, ...) -> float* with DCA_B200_GPUS=%d: FASTA parse + dedup + reweight + "
                           "per-GPU H2D + owner-sharded fit (one host thread per GPU, peer-memory exchange) + per-GPU "
                           "D2H of its slice into the returned host vector" % world}

    # ---- the asymmetric (per-site) estimator on the same workload and the same N GPUs: independent site
    # ranges, no exchange, one gather (SURVEY 8e last row; NOT the reference's estimator, reported beside it)
    asym = None
    if not args.no_extras:
        cpu_barrier()
        if rank == 0:
            try:
                _lib.fit_asymmetric(msa, w, Q, lam, lam, MAX_ITER, gpus=world)              # warm
                t1 = time.perf_counter()
                xa, ra = _lib.fit_asymmetric(msa, w, Q, lam, lam, MAX_ITER, gpus=world)
                a_wall = time.perf_counter() - t1
                xj = last_x[0] if world > 1 else _lib.plmdca_backend(path, "protein", SEQ_LEN, Q, SEQID, lam, lam, MAX_ITER)
                top = SEQ_LEN
                rank_a = np.argsort(-_lib.plm_fn_apc(xa, SEQ_LEN, Q)[1], kind="stable")[:top]
                rank_j = np.argsort(-_lib.plm_fn_apc(xj, SEQ_LEN, Q)[1], kind="stable")[:top]
                asym = {"workload": "C2, estimator = asymmetric: L independent per-site problems, contiguous site ranges on "
                                    "%d GPU(s), no exchange during the fits, one gather + symmetrisation" % world,
                        "call": "dca_plm_fit_asymmetric on host arrays (weights given): H2D + fits + gather + symmetrise + D2H",
                        "wall_ms": 1e3 * a_wall, "device_ms_slowest_gpu": ra.device_ms,
                        "evaluations_slowest_gpu": ra.evaluations, "iterations_slowest_gpu": ra.iterations,
                        "status": ra.status, "evals_per_s_slowest_gpu": ra.evaluations / (ra.device_ms * 1e-3),
                        "fn_apc_topL_overlap_with_joint_fit": int(len(set(rank_a.tolist()) & set(rank_j.tolist()))),
                        "topL": top}
            except Exception as exc:
                asym = {"error": repr(exc)}
        cpu_barrier()

    # ---- numerical check of the real multi-GPU run: the gathered parameters of the sharded fit
    # against a single-GPU fit of the same input on rank 0's GPU, in this run
    multi_check = None
    if world > 1:
        if rank == 0:
            try:
                solo = _lib.PlmObjective(msa, w, Q, lam, lam)
                xs, rs = solo.fit(solo.init(), MAX_ITER)
                solo.close()
                fn_m, fn_s = fn_scores(last_x[0], SEQ_LEN, Q), fn_scores(xs, SEQ_LEN, Q)
                multi_check = {"evaluations_single_gpu": rs.evaluations, "evaluations_multi_gpu": int(e_evals // n_e2e),
                               "fn_max_rel_dev_vs_single_gpu": float(np.max(np.abs(fn_m - fn_s) / np.abs(fn_s))),
                               "x_max_abs_dev": float(np.max(np.abs(last_x[0] - xs))),
                               "x_max_abs": float(np.max(np.abs(xs))),
                               "note": "gathered x of the %d-GPU owner-sharded fit vs a 1-GPU fit of the same MSA "
                                       "(same decisions, different reduction grouping across ranks)" % world}
            except Exception as exc:
                multi_check = {"error": repr(exc)}
        cpu_barrier()

    if rank == 0:
        line = {"metric": "plmdca_grad_evals_per_sec", "value": evals / (dev_ms * 1e-3), "unit": "evals/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(),
                "parallelism": parallelism(world, exchange), "arithmetic": "accurate (float64 fixed-order reductions); see parity",
                "evals_per_step": evals / max(1, args.steps), "lbfgs_status": status, "lbfgs_iterations": iters,
                "wall_ms_per_step": 1e3 * wall / max(1, args.steps),
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches)}
        if multi_check is not None:
            line["multi_gpu_check"] = multi_check
        if asym is not None:
            line["asymmetric"] = asym
        if world == 1:
            line["roofline"] = roofline_block(obj, x0, peaks)
            if not args.no_extras:
                for name, fn in (("parity", lambda: parity_block(msa, w, lam)), ("mfdca", lambda: mfdca_block(tmp, peaks)),
                                 ("reweight", reweight_block), ("c4", lambda: c4_block(find_hbm_peak(peaks)[0]))):
                    try:
                        line[name] = fn()
                    except Exception as exc:          # an extra block never takes the headline down
                        line[name] = {"error": repr(exc)}
            cores = os.cpu_count() or 1
            try:
                ev, dt, kind, cores = cpu_reference_fit(path, lam, CPU_BASELINE_ITERS, cores)
                line["cpu_baseline"] = {"value": ev / dt, "unit": "evals/s", "cores": cores, "kind": kind,
                                        "sample": "first %d L-BFGS iterations of the same C2 fit (%d evaluations, %.1f s) with "
                                                  "the reference's own lbfgs + PlmDCA::gradient built from /root/reference "
                                                  "(-fopenmp -O3), %d threads" % (CPU_BASELINE_ITERS, ev, dt, cores)}
            except Exception as exc:
                line["cpu_baseline"] = {"value": None, "unit": "evals/s", "cores": cores, "kind": "unavailable",
                                        "sample": repr(exc)}
        sys.stdout.flush()
        if saved_stdout is not None:
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        if saved_stdout is not None:
            os.dup2(2, 1)
    if sp is not None:
        sp.close()
    else:
        obj.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the parity / mfdca / reweight / c4 blocks")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
