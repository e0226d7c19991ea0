"""Config 3: heat-kernel wavelets of the identity impulse.

Two device evaluations of the same recurrence — dense impulse column blocks (csrc/cheby.cu, the shape of the
reference's pygsp call) and stored entries only (csrc/chebysp.cu) — must give the same bits, and both must equal
the oracle's restatement (oracle.cheby_op; pygsp itself is absent from the image: parity unpinned).
Includes the config-3 shape itself: N = 169 343, one 1024-column impulse block, m = 3, scales -0.5 / +0.5, tol 1e-4.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import oracle
from helpers import sym_graph

pytestmark = pytest.mark.gpu


def _connected_graph(n, m, seed):
    """sym_graph + a ring, so that no node is isolated (the sparse path needs every Laplacian diagonal)."""
    w = sym_graph(n, m, seed)
    i = np.arange(n)
    ring = sp.coo_matrix((np.ones(n), (i, (i + 1) % n)), shape=(n, n)).tocsr()
    w = w.maximum(ring).maximum(ring.T).tocsr()
    w.sort_indices()
    return w


def _phis(w, order, tol, lmax, method, block=256, normalize=False, scale=0.5):
    from scalable_roubust_gnn_b200 import spectral
    ws = spectral.WaveletSparsifier(w, scale=scale, approximation_order=order, tolerance=tol, lmax=lmax, block=block,
                                    method=method)
    return ws.calculate_all_wavelets(normalize=normalize), ws


@pytest.mark.parametrize("order", [1, 2, 3, 5])
def test_sparse_equals_blocks_equals_oracle(order):
    n = 900
    w = _connected_graph(n, 2500, order)
    lap_h = oracle.combinatorial_laplacian(w)
    lmax = oracle.estimate_lmax(lap_h)
    got_s, ws = _phis(w, order, 1e-4, lmax, "sparse")
    assert ws.stats and (order == 1 or ws.stats["products"] > 0)        # the sparse path really ran
    got_b, _ = _phis(w, order, 1e-4, lmax, "blocks", block=200)
    for tau, ps, pb in zip((-0.5, 0.5), got_s, got_b):
        c = oracle.cheby_coeff_heat(tau, lmax, order)
        want = oracle.wavelet_threshold(oracle.cheby_op(lap_h, [c], np.eye(n), lmax)[0], 1e-4)
        want.sort_indices()
        for got in (ps, pb):
            assert got.dtype == np.float32
            np.testing.assert_array_equal(got.indptr, want.indptr)
            np.testing.assert_array_equal(got.indices, want.indices)
            np.testing.assert_array_equal(got.data, want.data)
    # L1 row normalisation on top (sklearn's arithmetic, oracle.l1_normalize_rows)
    normed, _ = _phis(w, order, 1e-4, lmax, "sparse", normalize=True)
    for phi, raw in zip(normed, got_s):
        ref = oracle.l1_normalize_rows(raw)
        np.testing.assert_array_equal(phi.indices, ref.indices)
        np.testing.assert_array_equal(phi.data, ref.data)


def test_no_threshold_and_negative_scale_values():
    """tol = None keeps every coefficient whose float32 value is non-zero, whatever its sign."""
    n = 400
    w = _connected_graph(n, 900, 3)
    lap_h = oracle.combinatorial_laplacian(w)
    lmax = oracle.estimate_lmax(lap_h)
    got, ws = _phis(w, 3, None, lmax, "sparse")
    assert ws.stats
    for tau, phi in zip((-0.5, 0.5), got):
        c = oracle.cheby_coeff_heat(tau, lmax, 3)
        dense = oracle.cheby_op(lap_h, [c], np.eye(n), lmax)[0]
        want = sp.csr_matrix(dense.astype(np.float32))
        want.sort_indices()
        np.testing.assert_array_equal(phi.indices, want.indices)
        np.testing.assert_array_equal(phi.data, want.data)


def test_isolated_nodes_take_the_block_path():
    n = 300
    w = sym_graph(n, 250, 4)                       # sparse enough to leave isolated nodes
    assert (np.diff(w.indptr) == 0).any()
    lap_h = oracle.combinatorial_laplacian(w)
    lmax = oracle.estimate_lmax(lap_h)
    got, ws = _phis(w, 3, 1e-4, lmax, "sparse")
    assert not ws.stats                            # fell back
    c = oracle.cheby_coeff_heat(0.5, lmax, 3)
    want = oracle.wavelet_threshold(oracle.cheby_op(lap_h, [c], np.eye(n), lmax)[0], 1e-4)
    want.sort_indices()
    np.testing.assert_array_equal(got[1].indices, want.indices)
    np.testing.assert_array_equal(got[1].data, want.data)


def test_block_driver_20k_nodes_equals_sparse():
    """The 1000-column block driver of SpectralModel (base_model.py:236-265) over 20 blocks against the sparse path."""
    n = 20000
    w = _connected_graph(n, 70000, 11)
    lmax = 1.01 * float(2 * np.diff(w.indptr).max())          # any upper bound serves as an input here
    got_s, ws = _phis(w, 3, 1e-4, lmax, "sparse")
    got_b, _ = _phis(w, 3, 1e-4, lmax, "blocks", block=1000)
    assert ws.stats["products"] > 0
    for ps, pb in zip(got_s, got_b):
        np.testing.assert_array_equal(ps.indptr, pb.indptr)
        np.testing.assert_array_equal(ps.indices, pb.indices)
        np.testing.assert_array_equal(ps.data, pb.data)


def test_arxiv_shape_block_vs_oracle_and_sparse():
    """Config 3 at its real shape: N = 169 343, nnz = 1 166 243, m = 3, scales -0.5 / +0.5, tol 1e-4.  One 1024-column
    impulse block through the dense-block kernel against oracle.cheby_op on the same block (fp64, exact), and the same
    columns of the sparse all-columns result."""
    from scalable_roubust_gnn_b200 import device as dev, spectral, synth
    n, nnz, _, _ = synth.SHAPES["arxiv"]
    w = synth.uniform_graph(n, nnz)
    i = np.arange(n)
    ring = sp.coo_matrix((np.ones(n), (i, (i + 1) % n)), shape=(n, n)).tocsr()
    w = w.maximum(ring).maximum(ring.T).tocsr()
    w.sort_indices()
    lap_h = oracle.combinatorial_laplacian(w)
    lmax = 1.01 * float(2 * np.diff(w.indptr).max())
    coeffs = np.stack([oracle.cheby_coeff_heat(t, lmax, 3) for t in (-0.5, 0.5)])
    j0, b = 50 * 1024, 1024
    x = np.zeros((n, b))
    x[j0 + np.arange(b), np.arange(b)] = 1.0
    want = oracle.cheby_op(lap_h, coeffs, x, lmax)
    lap, _, _ = spectral.laplacian(dev.upload_csr(w))
    got, got32 = spectral.cheby_filter(lap, torch.from_numpy(x).cuda(), lmax, coeffs, tol=1e-4, want_f32=True)
    for g, g32, wv in zip(got, got32, want):
        wt = wv.copy()
        wt[wt < 1e-4] = 0
        assert torch.equal(g.cpu(), torch.from_numpy(wt))
        assert torch.equal(g32.cpu(), torch.from_numpy(wt.astype(np.float32)))
    ws = spectral.WaveletSparsifier(w, 0.5, 3, 1e-4, lmax=lmax, method="sparse")
    phis = ws.calculate_all_wavelets(normalize=False)
    assert ws.stats["products"] > 10_000_000
    for phi, wv in zip(phis, want):
        wt = wv.copy()
        wt[wt < 1e-4] = 0
        blk = sp.csr_matrix(wt.astype(np.float32))
        sub = phi[:, j0:j0 + b].tocsr()
        sub.sort_indices()
        blk.sort_indices()
        np.testing.assert_array_equal(sub.indptr, blk.indptr)
        np.testing.assert_array_equal(sub.indices, blk.indices)
        np.testing.assert_array_equal(sub.data, blk.data)


def test_device_lanczos_lmax_within_the_reference_tolerance():
    """estimate_lmax on the device (Lanczos) against the exact spectrum and against the reference's ARPACK call
    (oracle.estimate_lmax: eigsh k=1, tol 5e-3): both are 1.01 x lambda_max to 5e-3 relative."""
    from scalable_roubust_gnn_b200 import device as dev, spectral
    for n, m, seed in ((300, 900, 1), (2500, 9000, 2)):
        w = _connected_graph(n, m, seed)
        lap_h = oracle.combinatorial_laplacian(w)
        exact = float(np.linalg.eigvalsh(lap_h.toarray())[-1])
        lap, _, _ = spectral.laplacian(dev.upload_csr(w))
        got = spectral.estimate_lmax_device(lap)
        assert got <= 1.01 * exact * (1 + 1e-12)               # a Ritz value never exceeds lambda_max
        assert got >= 1.01 * exact * (1 - 5e-3)
        assert abs(got - oracle.estimate_lmax(lap_h)) <= 5e-3 * 1.01 * exact
    # the sparsifier uses it when no lmax is given
    ws = spectral.WaveletSparsifier(w, 0.5, 3, 1e-4)
    assert abs(ws.lmax - 1.01 * exact) <= 5e-3 * 1.01 * exact
