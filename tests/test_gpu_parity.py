"""GPU parity suite (-m gpu): every call goes through the C ABI of libsrgnn_b200.so and is
compared with the oracle / the golden vectors of the real reference."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import oracle
from helpers import assert_same_structure, golden_csr, sym_graph, ulp_diff32, ulp_diff64
from scalable_roubust_gnn_b200 import _lib, device as dev
from scalable_roubust_gnn_b200.operators import (PprGraphOp, SymLaplacianGraphOp, adj_to_symmetric_norm,
                                                 csr_sparse_dense_matmul)

pytestmark = pytest.mark.gpu

GRAPHS = ["cora", "rand_unw", "rand_w", "loop_iso"]
RS = [0.5, 0.0, 0.3, 1.0]
# fp64 weights: the exponents 0, +-0.5, +-1 are correctly rounded on the device while the golden
# values carry the build host's np.power rounding (SVML, <= 1 ulp per factor); for any other
# exponent CUDA pow is <= 2 ulp per factor, so the product of the two factors may differ by a few
# more ulp (1 ulp of fp64 = 2.2e-16 relative; the hops use the fp32 rounding of these weights).
ULP64 = 4
ULP64_GENERIC = 12


def ulp_tol(r):
    return ULP64 if r in (0.0, 0.5, 1.0) else ULP64_GENERIC


# ---- a3: normalisation -----------------------------------------------------------------------
@pytest.mark.parametrize("name", GRAPHS)
@pytest.mark.parametrize("r", RS)
def test_construct_adj_vs_reference_golden(golden_prop, name, r):
    adj = golden_csr(golden_prop, f"{name}_adj")
    want = golden_csr(golden_prop, f"{name}_r{r}_norm")
    got = SymLaplacianGraphOp(3, r=r).construct_adj(adj)
    assert isinstance(got, sp.csr_matrix) and got.indices.dtype == np.int32 and got.data.dtype == np.float64
    assert_same_structure(got, want)
    assert ulp_diff64(got.data, want.data).max() <= ulp_tol(r)
    # the fp32 weights the hops use (utils.py:39) agree except where fp64 sat on a rounding tie
    assert (ulp_diff32(got.data.astype(np.float32), want.data.astype(np.float32)) > 1).sum() == 0


@pytest.mark.parametrize("name", GRAPHS)
def test_ppr_construct_adj_vs_reference_golden(golden_prop, name):
    adj = golden_csr(golden_prop, f"{name}_adj")
    want = golden_csr(golden_prop, f"{name}_ppr_norm")
    got = PprGraphOp(2, r=0.5, alpha=0.15).construct_adj(adj)
    assert_same_structure(got, want)
    assert ulp_diff64(got.data, want.data).max() <= ULP64


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("weighted", [False, True])
def test_construct_adj_vs_oracle_random(dtype, weighted):
    adj = sym_graph(3000, 40000, 21, weighted=weighted, dtype=dtype)
    want = oracle.sym_norm(adj, 0.5)
    got = adj_to_symmetric_norm(adj, 0.5)
    assert_same_structure(got, want)
    assert ulp_diff64(got.data, want.data).max() <= ULP64


def test_degrees_bit_exact_weighted():
    """Degree = numpy add.reduceat order (first + pairwise of the rest), incl. rows > 128 entries."""
    n = 600
    adj = sym_graph(n, 60000, 5, weighted=True)          # ~170 entries per row: pairwise recursion
    a = dev.upload_csr(adj)
    _, flags, ex = dev.sym_norm(a, 0.5, want_degree=True)
    indptr, indices, d = oracle.selfloop_structure(adj)
    assert int(flags.item()) == _lib.SRG_FLAG_WEIGHTED           # informational bit only
    np.testing.assert_array_equal(ex["degree"].cpu().numpy(), d)
    np.testing.assert_array_equal(ex["count"].cpu().numpy(), np.diff(indptr))


def test_non_canonical_input_is_canonicalised():
    rows = np.array([0, 0, 1, 1, 2, 2, 0, 1]); cols = np.array([2, 1, 0, 2, 1, 0, 1, 0])
    adj = sp.csr_matrix((np.ones(8), (rows, cols)), shape=(3, 3))   # scipy sums the duplicates
    raw = sp.csr_matrix((np.ones(8), cols, np.array([0, 3, 6, 8])), shape=(3, 3))  # unsorted + dup
    want = oracle.sym_norm(raw, 0.5)
    got = adj_to_symmetric_norm(raw, 0.5)
    assert_same_structure(got, want)
    assert ulp_diff64(got.data, want.data).max() <= ULP64


# ---- a5/a6: one hop ------------------------------------------------------------------------------
@pytest.mark.parametrize("f", [1, 2, 3, 4, 7, 8, 24, 31, 32, 33, 64, 100, 128, 129, 257, 500, 1433])
def test_spmm_bit_exact_all_widths(f):
    """Every group width / chunk count of the kernel against the C oracle, bit for bit."""
    n = 700
    adj = oracle.sym_norm(sym_graph(n, 6000, f), 0.5)
    x = np.random.default_rng(f).standard_normal((n, f)).astype(np.float32)
    want = oracle.spmm_hop(adj, x)
    got = csr_sparse_dense_matmul(adj, x)                       # reference-ABI shim, host buffers
    assert got.dtype == np.float32 and got.shape == x.shape
    np.testing.assert_array_equal(got, want)
    # device entry point on the padded layout
    a = dev.upload_csr(adj.astype(np.float32))
    xp = dev.pack_features(torch.from_numpy(x).cuda())
    y = dev.unpack_features(dev.spmm(a, xp, f), f).cpu().numpy()
    np.testing.assert_array_equal(y, want)


@pytest.mark.parametrize("f", [4, 13, 32, 50, 56, 64, 100, 128])
@pytest.mark.parametrize("stages,rows,tile", [(2, 32, 0), (3, 16, 0), (2, 32, 1), (4, 5, 1), (2, 1, 0)])
def test_spmm_bulk_gather_bit_exact(f, stages, rows, tile):
    """The bulk-gather (TMA row copy) form of the hop kernel: same FMA chain, so bit for bit the C oracle, for every
    pipeline depth / rows per task / output path, on a graph with empty rows and rows spanning several chunks."""
    n = 900
    rng = np.random.default_rng(f + stages)
    adj = sym_graph(n, 9000, f).tolil()
    adj[5, :] = 0                                   # empty rows (a caller-supplied matrix may have them)
    adj[6, :] = 0
    adj[n - 1, :] = 0
    adj[40, ::3] = 1.5                              # 300 entries: ten chunks
    adj = sp.csr_matrix(adj)
    adj.data = (adj.data * rng.standard_normal(adj.nnz)).astype(np.float32)
    adj.eliminate_zeros()
    assert (np.diff(adj.indptr) == 0).sum() >= 3
    x = rng.standard_normal((n, f)).astype(np.float32)
    want = oracle.spmm_hop(adj, x)
    a = dev.upload_csr(adj)
    xp = dev.pack_features(torch.from_numpy(x).cuda())
    for key, val in (("bulk_gather", 1), ("bulk_stages", stages), ("bulk_rows", rows), ("bulk_tile", tile)):
        _lib.set_tuning(key, val)
    try:
        y = dev.unpack_features(dev.spmm(a, xp, f), f).cpu().numpy()
    finally:
        for key, val in (("bulk_gather", -1), ("bulk_stages", 2), ("bulk_rows", 32), ("bulk_tile", 0)):
            _lib.set_tuning(key, val)
    np.testing.assert_array_equal(y, want)


def test_spmm_bulk_gather_long_rows_and_rmat():
    """Hub rows: the segment tasks and the ordered combine run through the bulk kernel as well; short rows stay
    bit-identical, hub rows agree to fp32 rounding (segment sums), exactly as in the LDGSTS form."""
    n, f = 20000, 50
    adj = oracle.sym_norm(_hub_graph(n, 5, 5000, 3), 0.5)
    lens = np.diff(adj.indptr)
    x = np.random.default_rng(0).standard_normal((n, f)).astype(np.float32)
    want = oracle.spmm_hop(adj, x)
    a = dev.upload_csr(adj.astype(np.float32))
    xp = dev.pack_features(torch.from_numpy(x).cuda())
    res = {}
    for tile in (0, 1):
        _lib.set_tuning("bulk_gather", 1)
        _lib.set_tuning("bulk_tile", tile)
        try:
            res[tile] = dev.unpack_features(dev.spmm(a, xp, f), f).cpu().numpy()
        finally:
            _lib.set_tuning("bulk_gather", -1)
            _lib.set_tuning("bulk_tile", 0)
    short = lens <= 1024
    np.testing.assert_array_equal(res[0], res[1])                # the output path changes no bit
    np.testing.assert_array_equal(res[0][short], want[short])
    np.testing.assert_allclose(res[0][~short], want[~short], rtol=1e-5, atol=1e-5)


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
def test_spmm_bit_exact_vs_compiled_reference():
    adj = oracle.sym_norm(sym_graph(5000, 60000, 3), 0.5)
    x = np.random.default_rng(0).random((5000, 100), dtype=np.float32)
    np.testing.assert_array_equal(csr_sparse_dense_matmul(adj, x), oracle.spmm_hop(adj, x, lib="ref"))


def test_spmm_edge_cases():
    # empty rows, a hub row longer than several index tiles, negative / tiny weights
    n, f = 300, 20
    rng = np.random.default_rng(9)
    dense = np.zeros((n, n))
    dense[7, :] = rng.standard_normal(n)            # hub: 300 entries
    dense[:, 7] += rng.standard_normal(n) * 1e-20
    dense[50:60, 100:140] = rng.standard_normal((10, 40))
    adj = sp.csr_matrix(dense)
    x = rng.standard_normal((n, f)).astype(np.float32)
    np.testing.assert_array_equal(csr_sparse_dense_matmul(adj, x), oracle.spmm_hop(adj, x))
    # shim accumulates into a non-zero answer exactly like matmul.c:37
    lib = _lib.load()
    ans = rng.standard_normal(n * f).astype(np.float32)
    want = ans.copy()
    oracle._lib().oracle_spmm_csr_f32(want, f, adj.data.astype(np.float32), adj.indices.astype(np.int32),
                                      adj.indptr.astype(np.int32), x.reshape(-1), f, n, f)
    d32, ii, ip = adj.data.astype(np.float32), adj.indices.astype(np.int32), adj.indptr.astype(np.int32)
    lib.FloatCSRMulDenseOMP(ans.ctypes.data, d32.ctypes.data, ii.ctypes.data, ip.ctypes.data, x.ctypes.data, n, f)
    np.testing.assert_array_equal(ans, want)
    # empty matrix / zero features
    e = sp.csr_matrix((5, 5))
    np.testing.assert_array_equal(csr_sparse_dense_matmul(e, np.ones((5, 3), np.float32)), np.zeros((5, 3), np.float32))


# ---- a1: propagate ---------------------------------------------------------------------------------
@pytest.mark.parametrize("name", GRAPHS)
@pytest.mark.parametrize("r", RS)
def test_propagate_vs_reference_golden(golden_prop, name, r):
    adj = golden_csr(golden_prop, f"{name}_adj")
    x = golden_prop[f"{name}_x"]
    op = SymLaplacianGraphOp(3, r=r)
    hops = op.propagate(adj, x)
    assert len(hops) == 4 and all(isinstance(h, torch.Tensor) and h.dtype == torch.float32 and not h.is_cuda for h in hops)
    np.testing.assert_array_equal(hops[0].numpy(), x)
    np.testing.assert_allclose(hops[1].numpy(), golden_prop[f"{name}_r{r}_hop1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(hops[3].numpy(), golden_prop[f"{name}_r{r}_hop3"], rtol=1e-5, atol=1e-6)
    # lazily materialised op.adj is the reference's normalised matrix
    assert_same_structure(op.adj, golden_csr(golden_prop, f"{name}_r{r}_norm"))


@pytest.mark.parametrize("name", GRAPHS)
def test_ppr_propagate_vs_reference_golden(golden_prop, name):
    adj = golden_csr(golden_prop, f"{name}_adj")
    hops = PprGraphOp(2, r=0.5, alpha=0.15).propagate(adj, golden_prop[f"{name}_x"])
    np.testing.assert_allclose(hops[2].numpy(), golden_prop[f"{name}_ppr_hop2"], rtol=1e-5, atol=1e-6)


def test_propagate_inputs_tensor_and_fortran_order():
    adj = sym_graph(500, 3000, 2)
    x = np.asfortranarray(np.random.default_rng(0).random((500, 12), dtype=np.float32))
    want, _ = oracle.propagate(adj, x, 2)
    for feat in (x, torch.from_numpy(np.ascontiguousarray(x))):
        got = SymLaplacianGraphOp(2).propagate(adj, feat)
        for g, w in zip(got, want):
            np.testing.assert_allclose(g.numpy(), w, rtol=1e-5, atol=1e-6)


def test_propagate_cora_shape_config1(golden_prop):
    """BASELINE config 1: Cora topology (real bundled edges), F = 1433, K = 3, r = 0.5."""
    adj = golden_csr(golden_prop, "cora_adj")
    x = np.random.default_rng(1).random((2708, 1433), dtype=np.float32)
    want, norm = oracle.propagate(adj, x, 3)
    op = SymLaplacianGraphOp(3)
    got = op.propagate(adj, x)
    for g, w in zip(got, want):
        np.testing.assert_allclose(g.numpy(), w, rtol=1e-5, atol=1e-6)
    # from identical fp32 weights the hops are bit-exact
    a32 = dev.upload_csr(norm.astype(np.float32))
    hops = dev.propagate(a32, dev.pack_features(torch.from_numpy(x).cuda()), 1433, 3)
    for h, w in zip(hops, want):
        np.testing.assert_array_equal(dev.unpack_features(h, 1433).cpu().numpy(), w)


def test_propagate_pubmed_shape_config2_with_masks(golden_masks):
    """BASELINE config 2: PubMed topology, feature mask 0.6 / edge drop 0.6 (seed 2023), K = 5."""
    n, f = 19717, 500
    up = golden_masks["pubmed_0p6_0p6_edge_index"].astype(np.int64)     # bundled, already gathered
    adj = oracle.symmetrize_edges(up, n)
    torch.manual_seed(2023)
    fmask = oracle.feature_mask((n, f), 0.6)
    x = np.random.default_rng(1).random((n, f), dtype=np.float32)
    want, _ = oracle.propagate(adj, x * fmask.numpy(), 5)
    from scalable_roubust_gnn_b200.operators.utils import propagate_host
    hops, _ = propagate_host(adj, x, 5, 0.5, feature_mask=fmask)
    for g, w in zip(hops, want[1:]):
        np.testing.assert_allclose(g.numpy(), w, rtol=1e-5, atol=1e-6)


# ---- properties at larger size (oracle too slow / not needed) --------------------------------------
def test_large_graph_properties():
    """Linearity and row-stochasticity at 1M nodes: A^(r=1) rows sum to 1, so constant features
    are a fixed point; a hop is linear in X (exactly, for power-of-two scalings)."""
    n, f = 1_000_000, 16
    adj = sym_graph(n, 8_000_000, 1)
    a = dev.upload_csr(adj, ones_as_null=True)
    norm, flags, _ = dev.sym_norm(a, 1.0)          # D^0 A~^T D^-1 ... row sums of D^-1-scaled columns
    assert int(flags.item()) & ~_lib.SRG_FLAG_WEIGHTED == 0
    x = torch.rand((n, f), device="cuda")
    xp = dev.pack_features(x)
    y1 = dev.spmm(norm, xp, f)
    y2 = dev.spmm(norm, xp * 4.0, f)
    assert torch.equal(y2, y1 * 4.0)
    # r = 0: A^ = D^-1 A~^T -> rows sum to one for a symmetric graph
    norm0, _, _ = dev.sym_norm(a, 0.0)
    ones = dev.pack_features(torch.ones((n, f), device="cuda"))
    y = dev.spmm(norm0, ones, f)[:, :f]
    assert torch.allclose(y, torch.ones_like(y), rtol=1e-5, atol=1e-6)


# ---- a3 general paths ----------------------------------------------------------------------------
@pytest.mark.parametrize("r", [0.5, 0.3])
def test_asymmetric_adjacency_vs_reference_golden(golden_prop, r):
    """Directed input with a duplicate entry: canonicalise + explicit transpose on the device."""
    adj = golden_csr(golden_prop, "asym_adj")
    want = golden_csr(golden_prop, f"asym_r{r}_norm")
    op = SymLaplacianGraphOp(2, r=r)
    got = op.construct_adj(adj)
    assert_same_structure(got, want)
    assert ulp_diff64(got.data, want.data).max() <= ulp_tol(r)
    hops = op.propagate(adj, golden_prop["asym_x"])
    np.testing.assert_allclose(hops[2].numpy(), golden_prop[f"asym_r{r}_hop2"], rtol=1e-5, atol=1e-6)


def test_asymmetric_random_vs_oracle():
    rng = np.random.default_rng(4)
    n, m = 2000, 30000
    a = sp.csr_matrix((rng.random(m) + 0.5, (rng.integers(0, n, m), rng.integers(0, n, m))), shape=(n, n))
    a.sum_duplicates()
    want = oracle.sym_norm(a, 0.5)
    got = adj_to_symmetric_norm(a, 0.5)
    assert_same_structure(got, want)
    assert ulp_diff64(got.data, want.data).max() <= ULP64
    want_p = oracle.sym_norm(a, 0.3, ppr_alpha=0.1)
    got_p = adj_to_symmetric_norm(a, 0.3, ppr_alpha=0.1)
    assert_same_structure(got_p, want_p)
    assert ulp_diff64(got_p.data, want_p.data).max() <= ULP64_GENERIC


# ---- a8: masks ---------------------------------------------------------------------------------------
def test_edge_gather_and_csr_rebuild_reproduce_bundled_fixture(golden_prop, golden_masks):
    """cora_0_0.7: device gather == bundled edge_index (sha256), device CSR == symmetrised unique."""
    import hashlib
    from scalable_roubust_gnn_b200 import masks
    cora_e = golden_prop["cora_edges"].astype(np.int64)
    n = 2708
    adj = sp.csr_matrix((np.ones(cora_e.shape[1]), (cora_e[0], cora_e[1])), shape=(n, n))
    adj = (adj + adj.T).tocsr()
    torch.manual_seed(2023)
    fmask, keep, gathered, csr = masks.masked_graph(adj, (2708, 1433), 0.0, 0.7)
    sha = hashlib.sha256(gathered.cpu().numpy().astype(np.int64).tobytes()).digest()
    assert sha == golden_masks["cora_0_0p7_edge_index_sha"].tobytes()
    want = oracle.symmetrize_edges(gathered.cpu().numpy(), n)
    want.sort_indices()
    np.testing.assert_array_equal(csr.indptr.cpu().numpy(), want.indptr)
    np.testing.assert_array_equal(csr.indices[:csr.nnz].cpu().numpy(), want.indices)
    assert csr.nnz == want.nnz


def test_edges_to_sym_csr_with_duplicates_and_loops():
    from scalable_roubust_gnn_b200 import masks
    rng = np.random.default_rng(3)
    n, e = 5000, 60000
    ei = rng.integers(0, n, (2, e)).astype(np.int64)       # duplicates, both directions, self loops
    want = oracle.symmetrize_edges(ei, n)
    want.sort_indices()
    csr = masks.edges_to_sym_csr(torch.from_numpy(ei).cuda(), n)
    assert csr.nnz == want.nnz
    np.testing.assert_array_equal(csr.indptr.cpu().numpy(), want.indptr)
    np.testing.assert_array_equal(csr.indices[:csr.nnz].cpu().numpy(), want.indices)
    empty = masks.edges_to_sym_csr(torch.zeros((2, 0), dtype=torch.int64, device="cuda"), 7)
    assert empty.nnz == 0 and empty.indptr.cpu().tolist() == [0] * 8


def test_feature_mask_application_bit_exact():
    from scalable_roubust_gnn_b200 import masks
    torch.manual_seed(2023)
    m = masks.feature_mask((1000, 37), 0.6)
    x = torch.randn(1000, 37)
    got = dev.unpack_features(masks.apply_feature_mask(x.cuda(), m.cuda()), 37).cpu()
    assert torch.equal(got, x * m)


# ---- a9: Chebyshev heat wavelets (oracle = restated pygsp recurrence; parity unpinned) ---------------
def test_laplacian_bit_exact():
    from scalable_roubust_gnn_b200 import spectral
    for weighted in (False, True):
        w = sym_graph(3000, 40000, 8, weighted=weighted)
        want = oracle.combinatorial_laplacian(w)
        lap, deg, flags = spectral.laplacian(dev.upload_csr(w))
        assert int(flags.item()) == 0
        m = int(lap.indptr[-1].item())
        np.testing.assert_array_equal(lap.indptr.cpu().numpy(), want.indptr)
        np.testing.assert_array_equal(lap.indices[:m].cpu().numpy(), want.indices)
        np.testing.assert_array_equal(lap.data[:m].cpu().numpy(), want.data)


@pytest.mark.parametrize("order", [1, 3, 8])
@pytest.mark.parametrize("b", [1, 7, 64, 130])
def test_cheby_filter_bit_exact_vs_oracle(order, b):
    from scalable_roubust_gnn_b200 import spectral
    w = sym_graph(1500, 12000, 6)
    lap_h = oracle.combinatorial_laplacian(w)
    lmax = oracle.estimate_lmax(lap_h)
    coeffs = np.stack([oracle.cheby_coeff_heat(t, lmax, order) for t in (-0.5, 0.5)])
    np.testing.assert_array_equal(coeffs, np.stack([spectral.heat_cheby_coeffs(t, lmax, order) for t in (-0.5, 0.5)]))
    x = np.random.default_rng(b).standard_normal((1500, b))
    want = oracle.cheby_op(lap_h, coeffs, x, lmax)
    lap, _, _ = spectral.laplacian(dev.upload_csr(w))
    ld = (b + 1) // 2 * 2
    xd = torch.zeros((1500, ld), dtype=torch.float64, device="cuda")
    xd[:, :b] = torch.from_numpy(x).cuda()
    got = spectral.cheby_filter(lap, xd[:, :b], lmax, coeffs)
    for g, wv in zip(got, want):
        np.testing.assert_array_equal(g.cpu().numpy(), wv)            # fp64, same op order: exact
    got_t, got32 = spectral.cheby_filter(lap, xd[:, :b], lmax, coeffs, tol=1e-4, want_f32=True)
    for g, g32, wv in zip(got_t, got32, want):
        wt = wv.copy(); wt[wt < 1e-4] = 0
        np.testing.assert_array_equal(g.cpu().numpy(), wt)
        np.testing.assert_array_equal(g32.cpu().numpy(), wt.astype(np.float32))


def test_wavelet_sparsifier_matches_oracle():
    """WaveletSparsifier end to end on a small graph: same sparsity pattern, float32 values."""
    from scalable_roubust_gnn_b200 import spectral
    w = sym_graph(700, 2500, 12)
    lap_h = oracle.combinatorial_laplacian(w)
    lmax = oracle.estimate_lmax(lap_h)
    ws = spectral.WaveletSparsifier(w, scale=0.5, approximation_order=3, tolerance=1e-4, lmax=lmax, block=256)
    phis = ws.calculate_all_wavelets(normalize=False)
    wants = []
    for tau, phi in zip((-0.5, 0.5), phis):
        c = oracle.cheby_coeff_heat(tau, lmax, 3)
        dense = oracle.cheby_op(lap_h, [c], np.eye(700), lmax)[0]
        want = oracle.wavelet_threshold(dense, 1e-4)
        want.sort_indices()
        wants.append(want)
        assert phi.dtype == np.float32 and phi.has_sorted_indices
        np.testing.assert_array_equal(phi.indptr, want.indptr)       # device sparsification: exact pattern
        np.testing.assert_array_equal(phi.indices, want.indices)
        np.testing.assert_array_equal(phi.data, want.data)
    # L1 row normalisation with sklearn's float32 arithmetic
    normed = spectral.WaveletSparsifier(w, 0.5, 3, 1e-4, lmax=lmax, block=200).calculate_all_wavelets(normalize=True)
    for phi, want in zip(normed, wants):
        ref = oracle.l1_normalize_rows(want)
        np.testing.assert_array_equal(phi.indices, ref.indices)
        np.testing.assert_array_equal(phi.data, ref.data)


# ---- power-law rows ------------------------------------------------------------------------------------
def _hub_graph(n, hubs, hub_deg, seed):
    rng = np.random.default_rng(seed)
    base = sym_graph(n, 6 * n, seed)
    rows = np.repeat(np.arange(hubs), hub_deg)
    cols = rng.integers(0, n, hubs * hub_deg)
    extra = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n))
    a = base + extra + extra.T
    a.data[:] = 1.0
    a.setdiag(0)
    a.eliminate_zeros()
    a.sort_indices()
    return a.tocsr()


@pytest.mark.parametrize("f", [100, 128, 300])
def test_long_rows_split_deterministic_and_within_tolerance(f):
    """Rows above the long-row threshold are summed as ordered 1024-entry segments: every short row
    stays bit-identical to the reference chain, hub rows agree to fp32 rounding, and the result is
    reproducible run to run; with splitting disabled the hop is bit-exact everywhere."""
    n = 20000
    adj = oracle.sym_norm(_hub_graph(n, 5, 5000, 3), 0.5)
    lens = np.diff(adj.indptr)
    assert lens.max() > 4000 and (lens > 1024).sum() >= 5
    x = np.random.default_rng(0).standard_normal((n, f)).astype(np.float32)
    want = oracle.spmm_hop(adj, x)
    a = dev.upload_csr(adj.astype(np.float32))
    xp = dev.pack_features(torch.from_numpy(x).cuda())
    y1 = dev.unpack_features(dev.spmm(a, xp, f), f).cpu().numpy()
    y2 = dev.unpack_features(dev.spmm(a, xp, f), f).cpu().numpy()
    np.testing.assert_array_equal(y1, y2)                                   # deterministic
    short = lens <= 1024
    np.testing.assert_array_equal(y1[short], want[short])                   # untouched rows: exact
    np.testing.assert_allclose(y1[~short], want[~short], rtol=1e-5, atol=1e-5)
    _lib.set_tuning("long_row", 0)
    try:
        y3 = dev.unpack_features(dev.spmm(a, xp, f), f).cpu().numpy()
    finally:
        _lib.set_tuning("long_row", 1024)
    np.testing.assert_array_equal(y3, want)


def test_rmat_graph_propagation_vs_oracle():
    from scalable_roubust_gnn_b200 import synth
    adj = synth.rmat_graph(60000, 1_500_000, seed=2)
    assert np.diff(adj.indptr).max() > 1024
    x = synth.features(60000, 100)
    want, _ = oracle.propagate(adj, x, 2)
    got = SymLaplacianGraphOp(2).propagate(adj, x)
    for g, w in zip(got, want):
        np.testing.assert_allclose(g.numpy(), w, rtol=1e-5, atol=1e-6)


def test_explicit_zeros_and_cancelled_diagonal():
    """Stored zeros and a -1 diagonal (A + I cancels it): scipy drops both; so must the device path
    (fast kernels flag the zeros, the compacting kernels redo the job)."""
    n = 400
    base = sym_graph(n, 3000, 17, weighted=True).tolil()
    base[3, 3] = -1.0          # cancelled by +I -> row 3 has no diagonal in A~
    base[5, 5] = 2.5           # ordinary weighted self loop
    a = base.tocsr()
    a.sort_indices()
    # plant explicit zeros symmetrically
    rng = np.random.default_rng(1)
    coo = a.tocoo()
    pick = rng.choice(len(coo.data), 40, replace=False)
    dense_zero = set()
    for p in pick:
        i, j = int(coo.row[p]), int(coo.col[p])
        if i != j:
            dense_zero.add((i, j)); dense_zero.add((j, i))
    data = a.data.copy()
    rows = np.repeat(np.arange(n), np.diff(a.indptr))
    for k in range(len(data)):
        if (int(rows[k]), int(a.indices[k])) in dense_zero:
            data[k] = 0.0
    az = sp.csr_matrix((data, a.indices.copy(), a.indptr.copy()), shape=(n, n))
    assert (az.data == 0).sum() >= 40
    want = oracle.sym_norm(az, 0.5)
    got = adj_to_symmetric_norm(az, 0.5)
    assert_same_structure(got, want)
    assert ulp_diff64(got.data, want.data).max() <= ULP64
    assert want[3, 3] == 0 and got[3, 3] == 0


# ---- full BASELINE size: size-independent properties + sampled exact rows -----------------------------
def test_products_scale_properties():
    """BASELINE config 4 shape (N 2.45M, nnz(A^) 64.3M, F 100) on one GPU: the oracle is too slow for the
    whole graph, so check (a) D^-1 A~ is row-stochastic (constant features are a fixed point), (b) 256
    sampled output rows bit-for-bit against the C oracle fed with the same normalised rows, (c) run-to-run
    determinism, (d) K hops == K single hops, (e) integer structure invariants."""
    from scalable_roubust_gnn_b200 import synth
    n, nnz, f, k = synth.SHAPES["products"]
    adj = synth.uniform_graph(n, nnz)
    a = dev.upload_csr(adj, ones_as_null=True)
    norm, flags, ex = dev.sym_norm(a, 0.5, want_degree=True)
    assert int(flags.item()) == 0
    indptr = norm.indptr.cpu().numpy()
    m = int(indptr[-1])
    assert m == adj.nnz + n                                          # full diagonal added, nothing dropped
    np.testing.assert_array_equal(np.diff(indptr), np.diff(adj.indptr) + 1)
    np.testing.assert_array_equal(ex["degree"].cpu().numpy(), (np.diff(adj.indptr) + 1).astype(np.float64))
    x = torch.from_numpy(synth.features(n, f)).cuda()
    xp = dev.pack_features(x)
    hops = dev.propagate(norm, xp, f, 3)
    y1 = dev.spmm(norm, xp, f)
    # (c), (d) on the logical columns and on the whole padded buffers (pad columns are zero by construction)
    y2 = dev.spmm(norm, y1, f)
    assert torch.equal(y1[:, :f], hops[1][:, :f]) and torch.equal(y2[:, :f], hops[2][:, :f])
    assert torch.equal(y1, hops[1]) and torch.equal(y2, hops[2])
    assert float(y1[:, f:].abs().sum()) == 0.0
    # (b) sampled rows, exact
    rng = np.random.default_rng(5)
    rows = np.sort(rng.choice(n, 256, replace=False))
    idx_all = norm.indices[:m].cpu().numpy()
    val_all = norm.data[:m].cpu().numpy()
    sub_ptr = np.zeros(len(rows) + 1, dtype=np.int32)
    sub_ptr[1:] = np.cumsum(indptr[rows + 1] - indptr[rows])
    sel = np.concatenate([np.arange(indptr[r], indptr[r + 1]) for r in rows])
    sub = sp.csr_matrix((val_all[sel].astype(np.float64), idx_all[sel], sub_ptr), shape=(len(rows), n))
    want = oracle.spmm_hop(sub, x.cpu().numpy())
    np.testing.assert_array_equal(y1[:, :f].cpu().numpy()[rows], want)
    # values of the sampled rows against the formula d_a^-1/2 d_b^-1/2 (fp64 -> fp32)
    deg = (np.diff(adj.indptr) + 1).astype(np.float64)
    exp_vals = ((1.0 * deg[np.repeat(rows, np.diff(sub_ptr))] ** -0.5) * deg[idx_all[sel]] ** -0.5).astype(np.float32)
    assert (np.abs(val_all[sel].view(np.int32).astype(np.int64) - exp_vals.view(np.int32).astype(np.int64)) <= 1).all()
    # (a) r = 0: rows of D^-1 A~^T sum to one
    norm0, fl0, _ = dev.sym_norm(a, 0.0)
    ones = dev.pack_features(torch.ones((n, f), device="cuda"))
    y = dev.spmm(norm0, ones, f)[:, :f]
    assert torch.allclose(y, torch.ones_like(y), rtol=1e-5, atol=1e-6)


def test_reference_style_ctypes_binding():
    """Bind FloatCSRMulDenseOMP the way SSRG/operators/utils.py:21-45 does (numpy.ctypeslib, ndpointer
    argtypes, flattened float32 arrays) against our library file: the literal ABI drop-in."""
    import numpy.ctypeslib as ctl
    from ctypes import c_int
    import os
    lib = ctl.load_library("libsrgnn_b200.so", os.path.dirname(_lib.LIB_PATH))
    arr_i = ctl.ndpointer(dtype=np.int32, ndim=1, flags="CONTIGUOUS")
    arr_f = ctl.ndpointer(dtype=np.float32, ndim=1, flags="CONTIGUOUS")
    lib.FloatCSRMulDenseOMP.argtypes = [arr_f, arr_f, arr_i, arr_i, arr_f, c_int, c_int]
    lib.FloatCSRMulDenseOMP.restype = None
    adj = oracle.sym_norm(sym_graph(4000, 50000, 9), 0.5)
    feat = np.random.default_rng(2).random((4000, 47), dtype=np.float32)
    answer = np.zeros(feat.shape).astype(np.float32).flatten()
    lib.FloatCSRMulDenseOMP(answer, adj.data.astype(np.float32), adj.indices, adj.indptr, feat.flatten(), 4000, 47)
    np.testing.assert_array_equal(answer.reshape(feat.shape), oracle.spmm_hop(adj, feat))
    lib.FloatCSRMulDense.argtypes = [arr_f, c_int, arr_f, arr_i, arr_i, arr_f, c_int, c_int]
    lib.FloatCSRMulDense.restype = c_int
    answer2 = np.zeros(feat.shape).astype(np.float32).flatten()
    assert lib.FloatCSRMulDense(answer2, adj.nnz, adj.data.astype(np.float32), adj.indices, adj.indptr, feat.flatten(), 4000, 47) == 0
    np.testing.assert_array_equal(answer2, answer)


@pytest.mark.gpu
def test_propagate_device_output_equals_host_output(golden_prop):
    """SURVEY §8b placement opt-in: the same K+1 matrices, left on the GPU."""
    from scalable_roubust_gnn_b200.operators import MeanMessageOp, PprGraphOp, SymLaplacianGraphOp
    adj = golden_csr(golden_prop, "rand_unw_adj")
    x = golden_prop["rand_unw_x"]
    for op in (SymLaplacianGraphOp(3, r=0.5), PprGraphOp(2, r=0.5, alpha=0.15)):
        host = op.propagate(adj, x)
        devl = op.propagate(adj, x, device_output=True)
        assert len(devl) == len(host) and all(t.is_cuda and t.dtype == torch.float32 and tuple(t.shape) == x.shape for t in devl)
        for h, d in zip(host, devl):
            np.testing.assert_array_equal(d.cpu().numpy(), h.numpy())
        idx = torch.tensor([3, 1, 7])
        np.testing.assert_array_equal(devl[-1][idx].cpu().numpy(), host[-1][idx].numpy())       # base_model.py:84-87
        assert isinstance(op.adj, sp.csr_matrix)
    mean = MeanMessageOp(0, 4)
    np.testing.assert_array_equal(mean.aggregate(SymLaplacianGraphOp(3).propagate(adj, x, device_output=True)).cpu().numpy(),
                                  mean.aggregate(SymLaplacianGraphOp(3).propagate(adj, x)).numpy())
    # a directed input takes the general normalisation through construct_adj
    asym = golden_csr(golden_prop, "asym_adj")
    xa = golden_prop["asym_x"]
    got = SymLaplacianGraphOp(2, r=0.5).propagate(asym, xa, device_output=True)
    np.testing.assert_allclose(got[2].cpu().numpy(), golden_prop["asym_r0.5_hop2"], rtol=1e-5, atol=1e-6)
