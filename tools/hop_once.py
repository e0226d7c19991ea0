"""Run the normalisation + a few hops once on a products-shaped graph (ncu target)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scalable_roubust_gnn_b200 import device as dev, synth

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="products")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--hops", type=int, default=2)
ap.add_argument("--rmat", action="store_true")
args = ap.parse_args()
n, nnz, f, k = synth.SHAPES[args.workload]
n, nnz = int(n * args.scale), int(nnz * args.scale)
a = (synth.rmat_graph if args.rmat else synth.uniform_graph)(n, nnz)
x = synth.features(n, f)
a_dev = dev.upload_csr(a)
xp = dev.pack_features(torch.from_numpy(x).cuda())
norm, flags, _ = dev.sym_norm(a_dev, 0.5)
y = torch.empty_like(xp)
cur, nxt = xp, y
for _ in range(args.hops):
    dev.spmm(norm, cur, f, out=nxt)
    cur, nxt = nxt, cur
torch.cuda.synchronize()
print("flags", int(flags.item()), "checksum", float(cur[::1000].double().sum()))
