"""TwoDirLaplacianGraphOp and the sparse x sparse product behind it: golden outputs of the
reference's adj_to_un_in_out_dir_symmetric_norm (dense float32 products), the oracle restatement, the device
pipeline (sparse products, csrc/spgemm.cu)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import oracle
from conftest import GOLDEN as GOLDEN_DIR
from helpers import golden_csr

CASES = ["twodir", "twodir_loops"]


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN_DIR, "reference_ext.npz"))


def _close(got, want, rtol=2e-6):
    got = got.tocsr()
    got.sort_indices()
    np.testing.assert_array_equal(got.indptr, want.indptr)
    np.testing.assert_array_equal(got.indices, want.indices)
    np.testing.assert_allclose(got.data, want.data, rtol=rtol, atol=1e-9)


@pytest.mark.parametrize("tag", CASES)
def test_oracle_two_dir_norm_vs_reference_golden(g, tag):
    a = golden_csr(g, tag + "_adj")
    r, k = g[tag + "_params"]
    for nm, m in zip(("un", "in", "out"), oracle.two_dir_norm(a, float(r))):
        _close(m, golden_csr(g, f"{tag}_{nm}"))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_device_two_dir_norm_and_propagate_vs_reference_golden(g, tag):
    from scalable_roubust_gnn_b200.operators import TwoDirLaplacianGraphOp
    a = golden_csr(g, tag + "_adj")
    r, k = g[tag + "_params"]
    op = TwoDirLaplacianGraphOp(int(k), r=float(r))
    lists = op.propagate(a, g[tag + "_x"])
    assert len(lists) == 3
    for nm, m, hops in zip(("un", "in", "out"), (op.un_adj, op.in_adj, op.out_adj), lists):
        assert isinstance(m, sp.csr_matrix) and m.dtype == np.float32
        _close(m, golden_csr(g, f"{tag}_{nm}"))
        got = np.stack([t.numpy() for t in hops])
        np.testing.assert_allclose(got, g[f"{tag}_{nm}_hops"], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("n,dens", [(300, 0.02), (1000, 0.004), (64, 0.3)])
def test_spgemm_vs_scipy(n, dens):
    from scalable_roubust_gnn_b200.sparse_mm import csr_to_scipy, scipy_sparse_mat_to_device_adj, spgemm
    a = sp.random(n, n, dens, format="csr", dtype=np.float32, random_state=1)
    b = sp.random(n, n, dens, format="csr", dtype=np.float32, random_state=2)
    a.sort_indices()
    b.sort_indices()
    c = spgemm(scipy_sparse_mat_to_device_adj(a).csr, scipy_sparse_mat_to_device_adj(b).csr)
    want = (a.astype(np.float64) @ b.astype(np.float64)).tocsr()
    want.sort_indices()
    got = csr_to_scipy(c)
    assert c.nnz == want.nnz
    np.testing.assert_array_equal(got.indptr, want.indptr)
    np.testing.assert_array_equal(got.indices, want.indices)
    np.testing.assert_allclose(got.data, want.data, rtol=1e-5, atol=1e-7)


@pytest.mark.gpu
def test_spgemm_pattern_only_counts_paths_and_drops_zeros():
    """All-ones operands count 2-step walks exactly (integer work); cancelling products disappear with drop_zeros."""
    from scalable_roubust_gnn_b200.device import DeviceCSR
    from scalable_roubust_gnn_b200.sparse_mm import csr_to_scipy, scipy_sparse_mat_to_device_adj, spgemm
    a = sp.random(400, 400, 0.02, format="csr", dtype=np.float32, random_state=4)
    a.data[:] = 1.0
    a.sort_indices()
    d = scipy_sparse_mat_to_device_adj(a).csr
    pat = DeviceCSR(d.indptr, d.indices, None, d.n, d.nnz)
    got = csr_to_scipy(spgemm(pat, pat))
    want = (a @ a).tocsr()
    want.sort_indices()
    np.testing.assert_array_equal(got.indices, want.indices)
    np.testing.assert_array_equal(got.data, want.data)
    # [[1, 1], [0, 0]] @ [[1, 0], [-1, 0]] = 0 at (0, 0)
    x = sp.csr_matrix(np.array([[1, 1], [0, 0]], dtype=np.float32))
    y = sp.csr_matrix(np.array([[1, 0], [-1, 0]], dtype=np.float32))
    keep = spgemm(scipy_sparse_mat_to_device_adj(x).csr, scipy_sparse_mat_to_device_adj(y).csr)
    drop = spgemm(scipy_sparse_mat_to_device_adj(x).csr, scipy_sparse_mat_to_device_adj(y).csr, drop_zeros=True)
    assert keep.nnz == 1 and drop.nnz == 0
    np.testing.assert_array_equal(drop.indptr.cpu().numpy(), [0, 0, 0])


@pytest.mark.gpu
def test_spgemm_reproduces_wavelet_product():
    """spspmm(Psi, Psi^-1) (base_model.py:208-214) as a sparse product, then P X, against the two-hop form."""
    from scalable_roubust_gnn_b200.sparse_mm import DeviceAdj, scipy_sparse_mat_to_device_adj, spgemm
    from scalable_roubust_gnn_b200.spectral import wavelet_localize
    phi = sp.random(300, 300, 0.03, format="csr", dtype=np.float32, random_state=7)
    inv = sp.random(300, 300, 0.03, format="csr", dtype=np.float32, random_state=8)
    x = torch.from_numpy(np.random.default_rng(0).random((300, 10), dtype=np.float32)).cuda()
    pa, ia = scipy_sparse_mat_to_device_adj(phi), scipy_sparse_mat_to_device_adj(inv)
    prod = DeviceAdj(spgemm(pa.csr, ia.csr))
    np.testing.assert_allclose(prod.mm(x).cpu().numpy(), wavelet_localize(pa, ia, x).cpu().numpy(), rtol=1e-4, atol=1e-6)
