"""Time SpMM kernel variants on one graph; every variant is checked bit-for-bit against variant 0."""
import argparse, itertools, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scalable_roubust_gnn_b200 import _lib, device as dev, synth

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="products")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--rmat", action="store_true")
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--configs", default="all")
args = ap.parse_args()
n, nnz, f, k = synth.SHAPES[args.workload]
n, nnz = int(n * args.scale), int(nnz * args.scale)
a = (synth.rmat_graph if args.rmat else synth.uniform_graph)(n, nnz)
x = torch.from_numpy(synth.features(n, f)).cuda()
a_dev = dev.upload_csr(a, ones_as_null=True)
norm, flags, _ = dev.sym_norm(a_dev, 0.5)
torch.cuda.synchronize()
nnz_hat = int(norm.indptr[-1].item())
bg = nnz_hat * 8 + (n + 1) * 4 + nnz_hat * f * 4 + n * f * 4
print(f"N={n} nnz_hat={nnz_hat} F={f} B_gather={bg/1e9:.2f} GB", flush=True)

def timed(xp, y):
    for _ in range(3):
        dev.spmm(norm, xp, f, out=y)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
    evs[0].record()
    for i in range(args.reps):
        dev.spmm(norm, xp, f, out=y)
        evs[i + 1].record()
    torch.cuda.synchronize()
    ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.reps)]
    return min(ts), sum(ts) / len(ts)

def setk(**kw):
    for kk, v in kw.items():
        _lib.set_tuning(kk, v)

results = []
ref = {}
def run(tag, ld, **kw):
    setk(**kw)
    xp = dev.pack_features(x, ld=ld)
    y = torch.empty_like(xp)
    try:
        tmin, tavg = timed(xp, y)
    except Exception as e:
        print(tag, "FAILED", e, flush=True); return
    out = y[:, :f].contiguous()
    if "ref" not in ref:
        ref["ref"] = out
        ok = True
    else:
        ok = torch.equal(out, ref["ref"])
    gbs = bg / (tavg * 1e-3) / 1e9
    print(f"{tag:45s} ld={ld:4d} min {tmin:7.3f} ms avg {tavg:7.3f} ms  {gbs:7.0f} GB/s  frac {gbs/6551.7:.3f}  exact={ok}", flush=True)
    results.append({"tag": tag, "ld": ld, "ms_min": tmin, "ms_avg": tavg, "gbs": gbs, "exact": ok, **kw})

ld0 = dev.pad_ld(f)
run("group U4", ld0, spmm_variant=0, group_unroll=4)
run("group U8", ld0, spmm_variant=0, group_unroll=8)
for l2, batch, rows in itertools.product((1, 0), (4, 8), (2, 4, 8, 16)):
    run(f"stream B{batch} R{rows} l2_64={l2}", ld0, spmm_variant=1, stream_batch=batch, stream_rows=rows, gather_l2_64=l2)
best = min((r for r in results if r.get("spmm_variant") == 1), key=lambda r: r["ms_avg"])
print("best stream:", best, flush=True)
kw = {kk: best[kk] for kk in ("spmm_variant", "stream_batch", "stream_rows", "gather_l2_64")}
for ld in sorted({f if f % 4 == 0 else ld0, ld0, (f + 15) // 16 * 16, (f + 31) // 32 * 32}):
    run("best stream, ld sweep", ld, **kw)
for lr in (0, 256, 1024, 4096):
    run(f"best stream, long_row={lr}", ld0, long_row=lr, **kw)
json.dump(results, open("gpurun_out/tune_spmm.json", "w"), indent=1)
