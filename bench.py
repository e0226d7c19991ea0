#!/usr/bin/env python
"""bench.py — K-hop propagation throughput on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload products]

A "step" is one pass of the hot path over the synthetic graph: adjacency normalisation + K hops.
  value        : K * nnz(A^) * F / t_step (edge*feat/s), raw CSR and features resident in HBM.
  e2e          : the same through SymLaplacianGraphOp.propagate (host buffers in, host tensors out).
  roofline     : the SpMM hop kernel, algorithmic gather-model bytes / CUDA-event time / measured HBM peak.
  cpu_baseline : the reference's own matmul.c (oracle/_ref) on this box's host cores, one hop.
`--impl reference` times the reference CPU implementation alone (same metric / unit / config).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N, nnz of the symmetric adjacency without self loops, F, K)
    "cora": (2708, 10556, 1433, 3),
    "pubmed": (19717, 88648, 500, 5),
    "arxiv": (169343, 1166243, 128, 3),
    "products": (2449029, 61859140, 100, 3),
    # config 5: power-law (scrambled R-MAT) graph built per rank on the device; multi-GPU runs only
    "papers100M": (111059956, 1615685872, 128, 3),
}
METRIC = "K-hop SpMM propagation throughput (normalisation + K hops)"
UNIT = "edge*feat/s"


_JSON_OUT = None


def emit(line: dict) -> None:
    """The one JSON line of this run, on the real stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def synth_graph(n, nnz, seed=0):
    from scalable_roubust_gnn_b200 import synth
    return synth.uniform_graph(n, nnz, seed)


def synth_features(n, f, seed=1):
    from scalable_roubust_gnn_b200 import synth
    return synth.features(n, f, seed)


def gather_bytes(n, nnz_hat, f):
    """CSR once, one F-vector per edge, output once."""
    return nnz_hat * 8 + (n + 1) * 4 + nnz_hat * f * 4 + n * f * 4


def comp_bytes(n, nnz_hat, f):
    return nnz_hat * 8 + (n + 1) * 4 + 2 * n * f * 4


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for nm, val in zip(names, p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
REF_SAMPLE_SCALE = 0.25   # the CPU arm runs the same generator at 1/4 of N and nnz (bounded sample)


_SAMPLE_CACHE = {}


def cpu_reference_path(n, nnz, f, k, reps=1):
    """The reference's CPU path on this host over a graph of the same generator: scipy normalisation
    (utils.py:81-93, restated in oracle.sym_norm) + K hops of its own matmul.c (oracle/_ref, OpenMP, all
    host threads; utils.py:38-47 marshalling included).  Returns a dict of timings."""
    import oracle
    kind = "reference" if oracle.have_ref() else "port"
    lib = "ref" if kind == "reference" else "oracle"
    if (n, nnz, f) not in _SAMPLE_CACHE:
        _SAMPLE_CACHE.clear()
        _SAMPLE_CACHE[(n, nnz, f)] = (synth_graph(n, nnz), synth_features(n, f))
    a, x = _SAMPLE_CACHE[(n, nnz, f)]
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        adj_norm = oracle.sym_norm(a, 0.5)
        t_norm = time.perf_counter() - t0
        cur = x
        t_hops = []
        for _ in range(k):
            t1 = time.perf_counter()
            cur = oracle.spmm_hop(adj_norm, cur, lib=lib)
            t_hops.append(time.perf_counter() - t1)
        total = time.perf_counter() - t0
        if best is None or total < best["total_s"]:
            best = {"total_s": total, "norm_s": t_norm, "hop_s": float(np.mean(t_hops)), "nnz_hat": int(adj_norm.nnz),
                    "N": n, "kind": kind}
    best["value"] = k * best["nnz_hat"] * f / best["total_s"]
    best["hop_value"] = best["nnz_hat"] * f / best["hop_s"]
    return best


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (scipy normalisation + K hops
    of its matmul.c compiled in place), all host threads, on a bounded 1/4-scale sample of the workload
    per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, nnz, f, k = WORKLOADS[args.workload]
    n_full, nnz_full = int(n * args.scale), int(nnz * args.scale)
    ns, nnzs = int(n_full * REF_SAMPLE_SCALE), int(nnz_full * REF_SAMPLE_SCALE)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_path(ns, nnzs, f, k)
    runs = [cpu_reference_path(ns, nnzs, f, k) for _ in range(max(1, args.steps))]
    dt = float(np.mean([r["total_s"] for r in runs]))
    nnz_hat_s = runs[0]["nnz_hat"]
    val = k * nnz_hat_s * f / dt
    cores = os.cpu_count()
    sample = (f"per step: full reference path (scipy normalisation + K={k} hops of matmul.c FloatCSRMulDenseOMP, "
              f"OpenMP) on a {REF_SAMPLE_SCALE:g}-scale graph of the same generator (N={ns}, nnz_hat={nnz_hat_s})")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, n_full, nnz_full + n_full, f, k),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": runs[0]["kind"], "sample": sample,
                         "norm_s": float(np.mean([r["norm_s"] for r in runs])),
                         "hop_s": float(np.mean([r["hop_s"] for r in runs])),
                         "hop_only_value": float(np.mean([r["hop_value"] for r in runs]))},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, n, nnz_hat, f, k):
    kind = "power-law (scrambled R-MAT, device-generated)" if args.workload.startswith("papers100M") else "uniform"
    return {"workload": f"{args.workload}-shaped synthetic {kind} graph", "N": n, "nnz_hat": int(nnz_hat), "F": f,
            "K": k, "r": 0.5, "scale": args.scale, "l2": "inputs (X, CSR) exceed L2; no flush needed" if n * f * 4 > 126e6
            else "flushed between steps"}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from scalable_roubust_gnn_b200 import _lib, device as dev
    from scalable_roubust_gnn_b200.operators import SymLaplacianGraphOp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from scalable_roubust_gnn_b200 import dist_bench
        return dist_bench.run(args, WORKLOADS, METRIC, UNIT, emit)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    lib = _lib.load()

    n, nnz, f, k = WORKLOADS[args.workload]
    n, nnz = int(n * args.scale), int(nnz * args.scale)
    t0 = time.perf_counter()
    a = synth_graph(n, nnz)
    x = synth_features(n, f)
    log(f"[bench] graph N={n} nnz={a.nnz} F={f} K={k} generated in {time.perf_counter() - t0:.1f}s")

    # ---- device-resident arm ---------------------------------------------------------------
    a_dev = dev.upload_csr(a)                    # raw adjacency with its float64 ones, as scipy holds it
    x_dev = dev.pack_features(torch.from_numpy(x).cuda())
    ld = x_dev.shape[1]
    hops = [x_dev] + [torch.empty_like(x_dev) for _ in range(k)]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda") if n * f * 4 <= 126e6 else None

    def step(ev=None):
        norm, flags, _ = dev.sym_norm(a_dev, 0.5)
        if ev is not None:
            ev[0].record()
        for i in range(1, k + 1):
            dev.spmm(norm, hops[i - 1], f, out=hops[i])
            if ev is not None:
                ev[i].record()
        return norm, flags

    # the clock sampler (an nvidia-smi child polling every 100 ms) starts BEFORE the warm-up: its start-up
    # (NVML initialisation) stalls the first kernels that follow it, which must not be timed ones
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        norm, flags = step()
    torch.cuda.synchronize()
    assert int(flags.item()) & ~16 == 0, f"normalisation flags {int(flags.item())}"
    nnz_hat = int(norm.indptr[-1].item())
    step()                                      # one more untimed pass after the readbacks above
    torch.cuda.synchronize()
    sampler.lines.clear()                       # keep only samples taken during the timed region
    launches0 = _lib.launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(k + 2)] for _ in range(args.steps)]
    torch.cuda.synchronize()
    for s in range(args.steps):
        if flush is not None:
            flush.zero_()
        evs[s][k + 1].record()          # step start
        step(evs[s])
    torch.cuda.synchronize()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    step_ms = [evs[s][k + 1].elapsed_time(evs[s][k]) for s in range(args.steps)]
    norm_ms = [evs[s][k + 1].elapsed_time(evs[s][0]) for s in range(args.steps)]
    hop_ms = [evs[s][i - 1].elapsed_time(evs[s][i]) for s in range(args.steps) for i in range(1, k + 1)]
    t_step = float(np.mean(step_ms)) * 1e-3
    value = k * nnz_hat * f / t_step
    peak, peak_src = measured_peak()
    hop_avg = float(np.mean(hop_ms)) * 1e-3
    bg = gather_bytes(n, nnz_hat, f)
    achieved = bg / hop_avg / 1e9
    roofline = {"bound": "hbm", "kernel": "spmm_stream_kernel<4,1,0> (one hop)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": None,
                "algorithmic_bytes_per_launch": bg, "compulsory_bytes_per_launch": comp_bytes(n, nnz_hat, f),
                "hop_ms_avg": hop_avg * 1e3, "hop_ms_min": float(np.min(hop_ms)), "frac_of_8TBps_nominal": achieved / 8000.0}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(args.workload)
        except Exception:
            pass

    # ---- end to end through the public operator API (host buffers) ----------------------------
    x_pin = torch.from_numpy(x).pin_memory()
    a_pin = sp.csr_matrix((torch.from_numpy(a.data).pin_memory().numpy(), torch.from_numpy(a.indices).pin_memory().numpy(),
                           torch.from_numpy(a.indptr).pin_memory().numpy()), shape=a.shape, copy=False)
    op = SymLaplacianGraphOp(k, r=0.5)
    e2e_steps = max(1, min(args.steps, 10))
    for _ in range(2):
        out = op.propagate(a_pin, x_pin.numpy())
    del out
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = op.propagate(a_pin, x_pin.numpy())
    t_e2e = (time.perf_counter() - t0) / e2e_steps
    checksum = float(out[-1][:: max(1, n // 1000)].double().sum())
    # the same call with the message operator folded in (SGC: last hop, SSGC: mean), SURVEY 8f-1
    from scalable_roubust_gnn_b200.operators import LastMessageOp, MeanMessageOp, OverSmoothDistanceWeightedOp
    fused = {}
    for nm, mop in (("last", LastMessageOp()), ("mean", MeanMessageOp(0, k + 1)), ("nafs", OverSmoothDistanceWeightedOp())):
        try:
            op.propagate_aggregate(a_pin, x_pin.numpy(), mop)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                agg = op.propagate_aggregate(a_pin, x_pin.numpy(), mop)
            fused[nm] = (time.perf_counter() - t0) / e2e_steps * 1e3
        except Exception as exc:           # an auxiliary figure must never take the bench line down
            fused[nm] = f"failed: {exc}"
    h2d = a.indptr.nbytes + a.indices.nbytes + a.data.nbytes + x.nbytes
    d2h = k * x.nbytes
    e2e = {"value": k * nnz_hat * f / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": t_e2e * 1e3, "api": "SymLaplacianGraphOp(K).propagate(scipy_csr, float32 ndarray) -> K+1 CPU tensors",
           "checksum": checksum,
           "fused_message_op_ms": {"last (SGC)": fused["last"], "mean (SSGC)": fused["mean"], "over_smooth_distance (NAFS)": fused["nafs"],
                                   "note": "propagate_aggregate: only the aggregate is copied back"}}

    # ---- CPU baseline on this host (bounded sample: the full reference path at 1/4 scale) ------------
    cpu = None
    if not args.no_cpu_baseline:
        ns, nnzs = int(n * REF_SAMPLE_SCALE), int(nnz * REF_SAMPLE_SCALE)
        r = cpu_reference_path(ns, nnzs, f, k)
        cpu = {"value": r["value"], "unit": UNIT, "cores": os.cpu_count(), "kind": r["kind"],
               "sample": f"full reference path (scipy normalisation + K={k} hops of matmul.c FloatCSRMulDenseOMP, OpenMP, "
                         f"all host threads) on a {REF_SAMPLE_SCALE:g}-scale graph of the same generator "
                         f"(N={r['N']}, nnz_hat={r['nnz_hat']})",
               "norm_s": r["norm_s"], "hop_s": r["hop_s"], "hop_only_value": r["hop_value"]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, n, nnz_hat, f, k),
            "norm_ms": float(np.mean(norm_ms)), "hop_ms": float(np.mean(hop_ms)),
            "step_ms_all": [round(v, 3) for v in step_ms], "norm_ms_all": [round(v, 3) for v in norm_ms],
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    emit(line)


def main():
    # stdout must carry ONE JSON line, but libraries (NCCL's version banner) write to fd 1 too: keep a
    # private copy of the real stdout for the JSON and point fd 1 at stderr for everybody else
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="products", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink N and nnz (smoke runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
