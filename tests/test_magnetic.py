"""Magnetic-Laplacian operators of directed graphs: golden outputs of the reference's own
SymDirMagLaplacianGraphOp / SymDirMagComPprGraphOp (tests/golden/make_golden_ext.py), the oracle restatement,
and the device pipeline (csrc/magnetic.cu + ComGraphOp.propagate)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import oracle
from conftest import GOLDEN as GOLDEN_DIR
from helpers import golden_csr, ulp_diff64

CASES = ["mag_unw", "mag_w", "mag_tiny"]
# fp64 values: pow / sin / cos of the host libm, numpy and CUDA differ by an ulp or two per factor
ULP64 = 16


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN_DIR, "reference_ext.npz"))


def _case(g, tag):
    r, q, k = g[f"{tag}_params"]
    return golden_csr(g, f"{tag}_adj"), g[f"{tag}_x"], float(r), float(q), int(k)


def _assert_csr_close(got, want_prefix, g):
    want = golden_csr(g, want_prefix)
    got = got.tocsr()
    got.sort_indices()
    np.testing.assert_array_equal(got.indptr, want.indptr)
    np.testing.assert_array_equal(got.indices, want.indices)
    # tiny values (x * cos(pi/2) ~ 1e-17) have no meaningful ulp distance: absolute floor
    big = np.abs(want.data) > 1e-12
    assert ulp_diff64(got.data[big], want.data[big]).max(initial=0) <= ULP64
    np.testing.assert_allclose(got.data[~big], want.data[~big], atol=1e-15)


@pytest.mark.parametrize("tag", CASES)
def test_oracle_mag_norm_and_propagate_vs_reference_golden(g, tag):
    a, x, r, q, k = _case(g, tag)
    for alpha, pre in ((None, tag), (0.15, tag + "_ppr")):
        real, imag = oracle.mag_norm(a, r, q, alpha)
        _assert_csr_close(real, pre + "_real", g)
        _assert_csr_close(imag, pre + "_imag", g)
        # the propagation (with the reference's aliasing) on the reference's own matrices: bit-exact
        re, im = oracle.com_propagate(golden_csr(g, pre + "_real"), golden_csr(g, pre + "_imag"), x, k)
        np.testing.assert_array_equal(np.stack(re), g[pre + "_re_hops"])
        np.testing.assert_array_equal(np.stack(im), g[pre + "_im_hops"])


def test_com_graph_op_contract():
    from scalable_roubust_gnn_b200.operators import ComGraphOp, ComMessageOp, TwoDirGraphOp, TwoOrderPprApproxGraphOp
    op = ComGraphOp(2)
    with pytest.raises(TypeError, match="must be a scipy csr sparse matrix"):
        op.propagate(np.eye(3), np.ones((3, 2), dtype=np.float32))
    with pytest.raises(TypeError, match="must be a numpy.ndarray"):
        op.propagate(sp.eye(3).tocsr(), [[1.0]])
    with pytest.raises(ValueError, match="Dimension mismatch"):
        op.propagate(sp.eye(3).tocsr(), np.ones((4, 2), dtype=np.float32))
    assert TwoDirGraphOp(1).un_adj is None and TwoOrderPprApproxGraphOp(1).two_adj is None
    with pytest.raises(TypeError, match="The real feature matrices must be tensors!"):
        ComMessageOp().aggregate([np.ones(2)], [torch.ones(2)])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_device_mag_norm_vs_reference_golden(g, tag):
    from scalable_roubust_gnn_b200.operators import adj_to_directed_symmetric_mag_norm
    a, x, r, q, k = _case(g, tag)
    for alpha, pre in ((None, tag), (0.15, tag + "_ppr")):
        real, imag = adj_to_directed_symmetric_mag_norm(a.tocoo(), r, q, ppr_alpha=alpha)
        assert isinstance(real, sp.csr_matrix) and real.dtype == np.float64 and real.indices.dtype == np.int32
        _assert_csr_close(real, pre + "_real", g)
        _assert_csr_close(imag, pre + "_imag", g)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_com_graph_op_propagate_vs_reference_golden(g, tag):
    from scalable_roubust_gnn_b200.operators import SymDirMagComPprGraphOp, SymDirMagLaplacianGraphOp
    a, x, r, q, k = _case(g, tag)
    for op, pre in ((SymDirMagLaplacianGraphOp(k, r=r, q=q), tag),
                    (SymDirMagComPprGraphOp(k, r=r, q=q, ppr_alpha=0.15), tag + "_ppr")):
        re, im = op.propagate(a, x)
        assert len(re) == len(im) == k + 1 and all(isinstance(t, torch.Tensor) and not t.is_cuda for t in re + im)
        np.testing.assert_array_equal(re[0].numpy(), x)
        np.testing.assert_allclose(np.stack([t.numpy() for t in re]), g[pre + "_re_hops"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(np.stack([t.numpy() for t in im]), g[pre + "_im_hops"], rtol=1e-5, atol=1e-6)
        # same matrices => the hop chain itself (including the reference's aliasing) is bit-exact
        re_o, im_o = oracle.com_propagate(op.real_adj, op.imag_adj, x, k)
        np.testing.assert_array_equal(np.stack([t.numpy() for t in re]), np.stack(re_o))
        np.testing.assert_array_equal(np.stack([t.numpy() for t in im]), np.stack(im_o))


@pytest.mark.gpu
def test_com_graph_op_recurrence_mode_is_the_complex_power(g):
    """faithful=False: Z_k = (R + iI)^k x, checked against dense complex128."""
    from scalable_roubust_gnn_b200.operators import SymDirMagLaplacianGraphOp
    a, x, r, q, k = _case(g, "mag_unw")
    op = SymDirMagLaplacianGraphOp(3, r=r, q=q, faithful=False)
    re, im = op.propagate(a, x)
    z = x.astype(np.complex128)
    m = op.real_adj.toarray().astype(np.float32).astype(np.float64) + 1j * op.imag_adj.toarray().astype(np.float32)
    for step in range(1, 4):
        z = m @ z
        np.testing.assert_allclose(re[step].numpy(), z.real, rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(im[step].numpy(), z.imag, rtol=1e-4, atol=2e-6)


@pytest.mark.gpu
def test_multi_adjacency_bases_run_independent_chains():
    from helpers import sym_graph
    from scalable_roubust_gnn_b200.operators import TwoDirGraphOp
    a = sym_graph(200, 900, 8)
    n1, n2, n3 = oracle.sym_norm(a, 0.5), oracle.sym_norm(a, 0.0), oracle.sym_norm(a, 1.0)

    class Op(TwoDirGraphOp):
        def construct_adj(self, adj):
            return n1.tocsr(), n2.tocsr(), n3.tocsr()

    x = np.random.default_rng(0).random((200, 9), dtype=np.float32)
    op = Op(2)
    lists = op.propagate(a, x)
    assert len(lists) == 3 and op.un_adj is not None and op.out_adj is not None
    for hops, norm in zip(lists, (n1, n2, n3)):
        want = [x]
        for _ in range(2):
            want.append(oracle.spmm_hop(norm.tocsr(), want[-1]))
        for h, w in zip(hops, want):
            np.testing.assert_array_equal(h.numpy(), w)
