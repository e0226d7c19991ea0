"""Device timings of the "next" rows at the BASELINE.json shapes (CUDA events, median of reps):
NAFS aggregation, per-epoch GCN sparse product (forward + backward), magnetic normalisation, and the whole
wavelet pre-processing of config 3 (arxiv shape)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp, torch
from scalable_roubust_gnn_b200 import device as dev, synth
from scalable_roubust_gnn_b200.operators import adj_to_directed_symmetric_mag_norm
from scalable_roubust_gnn_b200.operators.message_operator import nafs_combine_device
from scalable_roubust_gnn_b200.sparse_mm import DeviceAdj
from scalable_roubust_gnn_b200.spectral import SpectralModel

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/perf_next.json")
ap.add_argument("--skip-wavelet", action="store_true")
args = ap.parse_args()
PEAK = 6551.7


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))


rep = {}
# ---- products shape: NAFS over K+1 = 4 hop matrices ------------------------------------------------
n, nnz, f, k = synth.SHAPES["products"]
a = synth.uniform_graph(n, nnz)
a_dev = dev.upload_csr(a)
xp = dev.pack_features(torch.from_numpy(synth.features(n, f)).cuda())
norm, flags, _ = dev.sym_norm(a_dev, 0.5)
hops = dev.propagate(norm, xp, f, k)
out = torch.empty_like(xp)
ms = timeit(lambda: nafs_combine_device(hops, f=f, out=out))
bytes_ = (k + 1) * n * f * 4 + n * f * 4
rep["nafs_products"] = {"ms": ms, "algorithmic_GB": bytes_ / 1e9, "GBps": bytes_ / ms / 1e6, "frac_of_measured_peak": bytes_ / ms / 1e6 / PEAK}
print("nafs_products", json.dumps(rep["nafs_products"]), flush=True)

# ---- products shape: per-epoch GCN product, hidden width 64 (forward + backward = 2 hops) ----------
adj = DeviceAdj(norm)
adj.transposed()
for width in (64, 128):
    h = torch.randn(n, width, device="cuda", requires_grad=True)
    g = torch.randn(n, width, device="cuda")
    def fb():
        h.grad = None
        y = torch.mm(adj, h)
        y.backward(g)
    ms = timeit(fb)
    nnz_hat = int(norm.indptr[-1].item())
    bg = 2 * (nnz_hat * 8 + (n + 1) * 4 + nnz_hat * width * 4 + n * width * 4)
    rep[f"gcn_mm_fwd_bwd_products_w{width}"] = {"ms": ms, "gather_GBps": bg / ms / 1e6, "frac_of_measured_peak": bg / ms / 1e6 / PEAK}
    print(f"gcn w{width}", json.dumps(rep[f"gcn_mm_fwd_bwd_products_w{width}"]), flush=True)
del hops, out, adj, h, g, xp

# ---- products-sized DIRECTED graph: magnetic normalisation (host in, host out; and kernels only) -----
rng = np.random.default_rng(0)
m = nnz // 2
u, v = rng.integers(0, n, m), rng.integers(0, n, m)
keep = u != v
d = sp.csr_matrix((np.ones(keep.sum()), (u[keep], v[keep])), shape=(n, n))
d.sum_duplicates(); d.data[:] = 1.0; d.sort_indices()
t0 = time.perf_counter()
re, im = adj_to_directed_symmetric_mag_norm(d, 0.5, 0.25)
torch.cuda.synchronize()
t1 = time.perf_counter()
re, im = adj_to_directed_symmetric_mag_norm(d, 0.5, 0.25)
t2 = time.perf_counter()
rep["mag_norm_products_directed"] = {"nnz_in": int(d.nnz), "nnz_out": int(re.nnz), "host_to_host_s_first": t1 - t0, "host_to_host_s": t2 - t1}
print("mag_norm", json.dumps(rep["mag_norm_products_directed"]), flush=True)
del a_dev, norm

# ---- config 3: arxiv shape, wavelet pre-processing end to end (m = 3, scales +-0.5, tol 1e-4) --------
if not args.skip_wavelet:
    n3, nnz3, f3, _ = synth.SHAPES["arxiv"]
    a3 = synth.uniform_graph(n3, nnz3)
    x3 = synth.features(n3, f3)
    deg = np.diff(a3.indptr)
    lmax = 2.0 * float(deg.max() + 1)          # upper bound stands in for ARPACK (timing only)
    model = SpectralModel(0.5, 3, 1e-4, lmax=lmax, block=1000)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    feat = model.preprocess(a3, x3)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    rep["spectral_preprocess_arxiv"] = {"seconds": t1 - t0, "N": n3, "blocks": -(-n3 // 1000), "density": model.density(),
                                        "out_shape": list(feat.shape)}
    print("spectral", json.dumps(rep["spectral_preprocess_arxiv"]), flush=True)
json.dump(rep, open(args.out, "w"), indent=1)
