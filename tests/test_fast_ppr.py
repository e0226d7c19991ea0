"""SymDirFastPprApproxGraphOp.  The oracle restatement is pinned to the reference in
tests/test_oracle.py; the device normalisers are compared with the reference's own outputs (reference_ext.npz)."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN as GOLDEN_DIR
from helpers import golden_csr


def test_operator_mirror_contract():
    from scalable_roubust_gnn_b200.operators.graph_operator import SymDirFastPprApproxGraphOp
    op = SymDirFastPprApproxGraphOp(3)
    assert (op.prop_steps, op.r, op.ppr_alpha) == (3, 0.5, 0.1)
    with pytest.raises(TypeError, match="must be a scipy csr sparse matrix"):
        op.propagate(np.eye(3), np.ones((3, 2), dtype=np.float32))


@pytest.mark.gpu
def test_device_fast_ppr_vs_reference_golden():
    from scalable_roubust_gnn_b200.operators.graph_operator import SymDirFastPprApproxGraphOp
    g = np.load(os.path.join(GOLDEN_DIR, "reference_ext.npz"))
    a, x = golden_csr(g, "ppr_adj"), g["ppr_x"]
    op = SymDirFastPprApproxGraphOp(2, r=0.5, ppr_alpha=0.1)
    hops = op.propagate(a, x)
    want = golden_csr(g, "fastppr_norm")
    got = op.adj.tocsr()
    got.sort_indices()
    np.testing.assert_array_equal(got.indptr, want.indptr)
    np.testing.assert_array_equal(got.indices, want.indices)
    np.testing.assert_allclose(got.data, want.data, rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(np.stack([h.numpy() for h in hops]), g["fastppr_hops"], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_device_fast_ppr_vs_oracle_larger():
    from helpers import sym_graph
    from scalable_roubust_gnn_b200.operators import utils as u
    import scipy.sparse as sp
    rng = np.random.default_rng(4)
    rows, cols = rng.integers(0, 800, 6000), rng.integers(0, 800, 6000)
    a = sp.csr_matrix((np.ones(6000), (rows, cols)), shape=(800, 800))
    a.data[:] = 1.0
    got = u.adj_to_fast_ppr_approx_symmetric_norm(a, 0.3, 0.15).tocsr()
    want = oracle.fast_ppr_norm(a, 0.3, 0.15)
    got.sort_indices()
    np.testing.assert_array_equal(got.indices, want.indices)
    np.testing.assert_allclose(got.data, want.data, rtol=2e-6, atol=1e-9)


@pytest.mark.gpu
def test_device_two_order_ppr_vs_reference_golden():
    from scalable_roubust_gnn_b200.operators.graph_operator import SymDirTwoOrderPprApproxGraphOp
    g = np.load(os.path.join(GOLDEN_DIR, "reference_ext.npz"))
    a, x = golden_csr(g, "ppr_adj"), g["ppr_x"]
    op = SymDirTwoOrderPprApproxGraphOp(2, r=0.5, ppr_alpha=0.1)
    h1, h2 = op.propagate(a, x)
    for got, key in ((op.one_adj, "twoorder_one"), (op.two_adj, "twoorder_two")):
        want = golden_csr(g, key)
        got = got.tocsr()
        got.sort_indices()
        np.testing.assert_array_equal(got.indptr, want.indptr)
        np.testing.assert_array_equal(got.indices, want.indices)
        # the reference's stationary vector comes from a float32 LAPACK eigendecomposition
        np.testing.assert_allclose(got.data, want.data, rtol=2e-5, atol=1e-8)
    np.testing.assert_allclose(np.stack([h.numpy() for h in h1]), g["twoorder_one_hops"], rtol=5e-5, atol=1e-6)
    np.testing.assert_allclose(np.stack([h.numpy() for h in h2]), g["twoorder_two_hops"], rtol=5e-5, atol=1e-6)
