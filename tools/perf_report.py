"""Per-stage device timings over the BASELINE.json shapes (CUDA events, min/median of reps)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scalable_roubust_gnn_b200 import _lib, device as dev, spectral, synth

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="cora,pubmed,arxiv,products,products-rmat")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--out", default="gpurun_out/perf_report.json")
args = ap.parse_args()
PEAK = 6551.7

def timeit(fn, reps):
    for _ in range(3):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    return float(np.min(ts)), float(np.median(ts))

flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
report = {}
for shape in args.shapes.split(","):
    name, _, kind = shape.partition("-")
    n, nnz, f, k = synth.SHAPES[name]
    if kind == "rmatdev":      # counter-based scrambled R-MAT + procedural features, built on the device
        a_dev = synth.rmat_shard_device(n, synth.rmat_draws(n, nnz), 0, n)
        deg = (a_dev.indptr[1:] - a_dev.indptr[:-1]).cpu().numpy()
        xp = synth.hash_features_device(n, f)
        x = xp[:, :f].contiguous()
    else:
        a = (synth.rmat_graph if kind == "rmat" else synth.uniform_graph)(n, nnz)
        deg = np.diff(a.indptr)
        x = torch.from_numpy(synth.features(n, f)).cuda()
        a_dev = dev.upload_csr(a)
        xp = dev.pack_features(x)
    y = torch.empty_like(xp)
    norm, flags, _ = dev.sym_norm(a_dev, 0.5)
    torch.cuda.synchronize()
    assert int(flags.item()) & ~16 == 0, int(flags.item())
    nnz_hat = int(norm.indptr[-1].item())
    small = n * f * 4 < 126e6
    def norm_fn():
        dev.sym_norm(a_dev, 0.5)
    def hop_fn():
        if small: flush.zero_()
        dev.spmm(norm, xp, f, out=y)
    def flush_fn():
        flush.zero_()
    t_norm = timeit(norm_fn, args.reps)
    t_hop = timeit(hop_fn, args.reps)
    t_flush = timeit(flush_fn, args.reps) if small else (0.0, 0.0)
    hop_ms = t_hop[1] - t_flush[1]
    bg = nnz_hat * 8 + (n + 1) * 4 + nnz_hat * f * 4 + n * f * 4
    bc = nnz_hat * 8 + (n + 1) * 4 + 2 * n * f * 4
    rec = {"N": n, "nnz_hat": nnz_hat, "F": f, "max_deg": int(deg.max()), "norm_ms": t_norm[1], "hop_ms": hop_ms,
           "hop_ms_min": t_hop[0] - t_flush[0], "gather_GBps": bg / hop_ms / 1e6, "frac_gather": bg / hop_ms / 1e6 / PEAK,
           "comp_GBps": bc / hop_ms / 1e6, "edge_feat_per_s": nnz_hat * f / hop_ms * 1e3, "l2_flushed": small}
    report[shape] = rec
    print(shape, json.dumps(rec), flush=True)
    if name == "arxiv":
        # config 3: heat-kernel Chebyshev m=3, scales +-0.5, signal X (N x 128), fp64
        lap, _, _ = spectral.laplacian(a_dev)
        lmax = 2.0 * float(deg.max() + 1)      # upper bound stands in for ARPACK (timing only)
        coeffs = np.stack([spectral.heat_cheby_coeffs(t, lmax, 3) for t in (-0.5, 0.5)])
        xd = x.double()
        def cheb_fn():
            flush.zero_()
            spectral.cheby_filter(lap, xd, lmax, coeffs, tol=1e-4, want_f32=True)
        t_ch = timeit(cheb_fn, args.reps)
        nnz_l = int(lap.indptr[-1].item())
        # per order: CSR (12 B/entry) + gathered rows (8 B/elem) + T read/write + r read/write per scale
        per_order = nnz_l * 12 + nnz_l * f * 8 + n * f * 8 * (2 + 2 * 2)
        rec2 = {"cheby_ms_3_orders_2_scales": t_ch[1] - t_flush[1], "GBps_gather_model": 3 * per_order / (t_ch[1] - t_flush[1]) / 1e6}
        report["arxiv-cheby"] = rec2
        print("arxiv-cheby", json.dumps(rec2), flush=True)
    del a_dev, xp, y, norm, x
    torch.cuda.empty_cache()
json.dump(report, open(args.out, "w"), indent=1)
