"""Touch every kernel once on small inputs (target for `compute-sanitizer --tool memcheck`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, scipy.sparse as sp, torch
from helpers import sym_graph
from scalable_roubust_gnn_b200 import _lib, device as dev, masks, spectral, synth
from scalable_roubust_gnn_b200.operators import PprGraphOp, SymLaplacianGraphOp, adj_to_symmetric_norm, csr_sparse_dense_matmul

rng = np.random.default_rng(0)
for n, m, f, w in [(300, 2000, 7, False), (1000, 9000, 100, True), (777, 5000, 133, False), (50, 100, 1433, False)]:
    a = sym_graph(n, m, 1, weighted=w)
    x = rng.random((n, f), dtype=np.float32)
    SymLaplacianGraphOp(2).propagate(a, x)
    PprGraphOp(1).propagate(a, x)
    SymLaplacianGraphOp(1, r=0.3).construct_adj(a)
    csr_sparse_dense_matmul(adj_to_symmetric_norm(a, 0.5), x)
# hub rows (segments), R-MAT, tiny
a = synth.rmat_graph(20000, 600000, seed=1)
x = synth.features(20000, 100)
SymLaplacianGraphOp(2).propagate(a, x)
SymLaplacianGraphOp(1).propagate(sp.identity(3, format="csr"), np.ones((3, 5), np.float32))
SymLaplacianGraphOp(1).propagate(sp.csr_matrix((4, 4)), np.ones((4, 100), np.float32))
# directed + duplicates + explicit zeros
r = sp.csr_matrix((rng.random(3000), (rng.integers(0, 400, 3000), rng.integers(0, 400, 3000))), shape=(400, 400))
adj_to_symmetric_norm(r, 0.5)
raw = sp.csr_matrix((np.ones(8), np.array([2, 1, 1, 0, 2, 0, 1, 0]), np.array([0, 3, 6, 8])), shape=(3, 3))
adj_to_symmetric_norm(raw, 0.5)
z = sym_graph(200, 1500, 2, weighted=True); z.data[::7] = 0.0
adj_to_symmetric_norm(z.maximum(z.T).tocsr() if False else sp.csr_matrix((z.data, z.indices, z.indptr), shape=z.shape), 0.5)
# masks / edges / cheby
torch.manual_seed(2023)
g = sym_graph(500, 3000, 3)
fm, keep, gathered, csr = masks.masked_graph(g, (500, 16), 0.5, 0.5)
dev.unpack_features(masks.apply_feature_mask(torch.rand(500, 16).cuda(), fm.cuda()), 16)
ws = spectral.WaveletSparsifier(g, 0.5, 3, 1e-4, lmax=40.0, block=128)
ws.calculate_all_wavelets()
torch.cuda.synchronize()
print("sanitize_small: ok, launches", _lib.launch_count())
