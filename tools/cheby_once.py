"""One Chebyshev filter pass on the arxiv shape (ncu target): m=3, scales +-0.5, signal N x 128 fp64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scalable_roubust_gnn_b200 import device as dev, spectral, synth
n, nnz, f, k = synth.SHAPES["arxiv"]
a = synth.uniform_graph(n, nnz)
lap, deg, flags = spectral.laplacian(dev.upload_csr(a))
lmax = 2.0 * float(np.diff(a.indptr).max() + 1)
coeffs = np.stack([spectral.heat_cheby_coeffs(t, lmax, 3) for t in (-0.5, 0.5)])
x = torch.from_numpy(synth.features(n, f)).cuda().double()
for _ in range(2):
    r, r32 = spectral.cheby_filter(lap, x, lmax, coeffs, tol=1e-4, want_f32=True)
torch.cuda.synchronize()
print("ok", float(r[0].sum()))
