"""One rank's local hop of a P_r x P_f grid, emulated on ONE GPU: rows [0, N / P_r) of the normalised
products-shaped matrix gather from ALL rows of a feature slice of ceil(F / P_f) columns.  Sweeps the
gather variants (LDGSTS stream kernel vs bulk / TMA row gathers, stages, rows per task, tile store),
checks every variant bit for bit against the stream kernel and prints ms per hop, gathered rows/s
and the gather-model bytes/s.

    python tools/hop_shard.py [--grid 4x2] [--workload products] [--scale 1.0] [--rmat] [--once VARIANT]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from scalable_roubust_gnn_b200 import _lib, device as dev, synth

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="products")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--grid", default="4x2", help="row blocks x feature slices, comma separated list")
ap.add_argument("--rmat", action="store_true")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--once", default=None, help="run ONE variant once (ncu target), e.g. 'bulk,2,32,0'")
args = ap.parse_args()

n, nnz, f, _ = synth.SHAPES[args.workload]
n, nnz = int(n * args.scale), int(nnz * args.scale)
a = (synth.rmat_graph if args.rmat else synth.uniform_graph)(n, nnz)
x = synth.features(n, f)
norm, flags, _ = dev.sym_norm(dev.upload_csr(a), 0.5)
torch.cuda.synchronize()
assert int(flags.item()) & ~16 == 0
indptr_h = norm.indptr.cpu().numpy()


def set_variant(v):
    kind, stages, rows, tile = v
    _lib.set_tuning("bulk_gather", 1 if kind == "bulk" else 0)
    if kind == "bulk":
        _lib.set_tuning("bulk_stages", stages)
        _lib.set_tuning("bulk_rows", rows)
        _lib.set_tuning("bulk_tile", tile)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


variants = [("stream", 0, 0, 0)] + [("bulk", s, r, t) for s in (2, 3) for r in (32, 16, 8) for t in (0, 1)]
if args.once:
    k, s, r, t = args.once.split(",")
    variants = [(k, int(s), int(r), int(t))]

results = []
for grid in args.grid.split(","):
    pr, pf = (int(v) for v in grid.split("x"))
    rows = -(-n // pr)
    f_loc = -(-f // pf)
    xs = dev.pack_features(torch.from_numpy(np.ascontiguousarray(x[:, :f_loc])).cuda())
    ld = xs.shape[1]
    nnz_loc = int(indptr_h[rows])
    out = torch.empty((rows, ld), dtype=torch.float32, device="cuda")
    ref = None
    for v in variants:
        set_variant(v)
        out.zero_()
        try:
            ms = timed(lambda: dev.spmm(norm, xs, f_loc, out=out, n_rows=rows), 1 if args.once else args.reps)
        except Exception as exc:
            print(grid, v, "FAILED", exc, flush=True)
            continue
        same = None
        if v[0] == "stream":
            ref = out.clone()
        elif ref is not None:
            same = bool(torch.equal(ref, out))
        gbytes = nnz_loc * 8 + (rows + 1) * 4 + nnz_loc * f_loc * 4 + rows * f_loc * 4
        r = {"grid": grid, "variant": "%s S=%d R=%d tile=%d" % v, "ld": ld, "rows": rows, "nnz": nnz_loc, "ms": round(ms, 4),
             "Grows_per_s": round(nnz_loc / ms / 1e6, 2), "gather_model_GBps": round(gbytes / ms / 1e6, 1),
             "bitwise_equal_to_stream": same}
        results.append(r)
        print(json.dumps(r), flush=True)
    del xs, out, ref
