"""CPU suite: the oracle against the golden vectors produced by the real reference, and against
the reference's own matmul.c compiled in place (oracle/_ref)."""
import hashlib
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import oracle
from conftest import GOLDEN as GOLDEN_DIR
from helpers import assert_same_structure, golden_csr, sym_graph, ulp_diff64

GRAPHS = ["cora", "rand_unw", "rand_w", "loop_iso"]
RS = [0.5, 0.0, 0.3, 1.0]


@pytest.mark.parametrize("name", GRAPHS)
@pytest.mark.parametrize("r", RS)
def test_sym_norm_matches_reference(golden_prop, name, r):
    adj = golden_csr(golden_prop, f"{name}_adj")
    want = golden_csr(golden_prop, f"{name}_r{r}_norm")
    got = oracle.sym_norm(adj, r)
    assert_same_structure(got, want)           # integer work: exact
    assert got.indices.dtype == np.int32 and got.data.dtype == np.float64
    # fp64 values: np.power's last bit depends on the host's SIMD libm; everything else is exact
    assert ulp_diff64(got.data, want.data).max() <= 2


@pytest.mark.parametrize("name", GRAPHS)
def test_ppr_norm_matches_reference(golden_prop, name):
    adj = golden_csr(golden_prop, f"{name}_adj")
    want = golden_csr(golden_prop, f"{name}_ppr_norm")
    got = oracle.sym_norm(adj, 0.5, ppr_alpha=0.15)
    assert_same_structure(got, want)
    assert ulp_diff64(got.data, want.data).max() <= 2


@pytest.mark.parametrize("r", [0.5, 0.3])
def test_asymmetric_matches_reference(golden_prop, r):
    adj = golden_csr(golden_prop, "asym_adj")
    want = golden_csr(golden_prop, f"asym_r{r}_norm")
    got = oracle.sym_norm(adj, r)
    assert_same_structure(got, want)
    assert ulp_diff64(got.data, want.data).max() <= 2
    hops, _ = oracle.propagate(adj, golden_prop["asym_x"], 2, r=r)
    np.testing.assert_allclose(hops[2], golden_prop[f"asym_r{r}_hop2"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", GRAPHS)
@pytest.mark.parametrize("r", RS)
def test_propagate_matches_reference(golden_prop, name, r):
    """Hops from the REFERENCE's normalised matrix: the SpMM restatement must be bit-exact."""
    norm = golden_csr(golden_prop, f"{name}_r{r}_norm")
    x = golden_prop[f"{name}_x"]
    h1 = oracle.spmm_hop(norm, x)
    np.testing.assert_array_equal(h1, golden_prop[f"{name}_r{r}_hop1"])
    h3 = oracle.spmm_hop(norm, oracle.spmm_hop(norm, h1))
    np.testing.assert_array_equal(h3, golden_prop[f"{name}_r{r}_hop3"])


@pytest.mark.parametrize("name", GRAPHS)
def test_propagate_end_to_end_tolerance(golden_prop, name):
    adj = golden_csr(golden_prop, f"{name}_adj")
    hops, _ = oracle.propagate(adj, golden_prop[f"{name}_x"], 3, r=0.5)
    np.testing.assert_allclose(hops[3], golden_prop[f"{name}_r0.5_hop3"], rtol=1e-6, atol=1e-7)


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("f", [1, 3, 8, 100, 129])
def test_c_oracle_bit_exact_vs_compiled_reference(f):
    adj = oracle.sym_norm(sym_graph(500, 4000, 7), 0.5)
    x = np.random.default_rng(f).standard_normal((500, f)).astype(np.float32)
    a = oracle.spmm_hop(adj, x, lib="oracle")
    b = oracle.spmm_hop(adj, x, lib="ref")
    np.testing.assert_array_equal(a, b)


def test_degree_is_numpy_reduceat_order():
    a = sym_graph(300, 9000, 11, weighted=True)
    indptr, indices, d = oracle.selfloop_structure(a)
    at = (a + sp.identity(300, format="csr")).tocsr()
    np.testing.assert_array_equal(d, np.add.reduceat(at.data, at.indptr[:-1]))


# ---- masks: torch CPU RNG stream, seed 2023, draw order rand -> randperm (data_process.py) ------
@pytest.mark.parametrize("ds,frate,erate", [("cora_0_0p7", 0.0, 0.7), ("cora_0p7_0p7", 0.7, 0.7),
                                            ("pubmed_0p6_0p6", 0.6, 0.6), ("citeseer_0p5_0p5", 0.5, 0.5)])
def test_edge_mask_reproduces_bundled_fixture(golden_masks, ds, frate, erate):
    shape = tuple(int(v) for v in golden_masks[ds + "_shape"])
    n_edges = {"cora": 5278, "pubm": 44324, "cite": 4552}[ds[:4]]
    torch.manual_seed(2023)
    fmask = oracle.feature_mask(shape, frate)
    emask = oracle.edge_mask(n_edges, erate)
    assert fmask.dtype == torch.int32 and fmask.shape == shape
    assert len(emask) == int(golden_masks[ds + "_edge_mask_len"][0])
    np.testing.assert_array_equal(emask[:64].numpy(), golden_masks[ds + "_edge_mask_head"])
    sha = hashlib.sha256(emask.numpy().astype(np.int64).tobytes()).digest()
    assert sha == golden_masks[ds + "_edge_mask_sha"].tobytes()


def test_edge_gather_reproduces_bundled_fixture(golden_prop, golden_masks):
    """cora_0_0.7/edge_index == canonical upper edges of cora_0_0 gathered by the mask."""
    cora_e = golden_prop["cora_edges"].astype(np.int64)
    n = 2708
    adj = sp.csr_matrix((np.ones(cora_e.shape[1]), (cora_e[0], cora_e[1])), shape=(n, n))
    adj = (adj + adj.T).tocsr()
    up = oracle.upper_edges(adj)
    torch.manual_seed(2023)
    oracle.feature_mask((2708, 1433), 0.0)
    keep = oracle.edge_mask(up.shape[1], 0.7).numpy()
    gathered = up[:, keep]
    sha = hashlib.sha256(gathered.astype(np.int64).tobytes()).digest()
    assert sha == golden_masks["cora_0_0p7_edge_index_sha"].tobytes()


def test_row_partition():
    rows_per, starts = oracle.row_partition(10, 4)
    assert rows_per == 3 and starts.tolist() == [0, 3, 6, 9, 10]
    rows_per, starts = oracle.row_partition(8, 8)
    assert starts.tolist() == list(range(9))
    rows_per, starts = oracle.row_partition(3, 8)
    assert starts.tolist() == [0, 1, 2, 3, 3, 3, 3, 3, 3]


def test_cheby_matches_exact_heat_kernel():
    """The restated recurrence converges to U exp(-tau Lambda / lmax) U^T."""
    a = sym_graph(60, 200, 5)
    lap = oracle.combinatorial_laplacian(a)
    lmax = oracle.estimate_lmax(lap)
    lam, u = np.linalg.eigh(lap.toarray())
    for tau in (0.5, -0.5):
        exact = (u * np.exp(-tau * lam / lmax)) @ u.T
        c = oracle.cheby_coeff_heat(tau, lmax, 30)
        got = oracle.cheby_op(lap, [c], np.eye(60), lmax)[0]
        assert np.abs(got - exact).max() < 1e-12
        c3 = oracle.cheby_coeff_heat(tau, lmax, 3)
        got3 = oracle.cheby_op(lap, [c3], np.eye(60), lmax)[0]
        assert np.abs(got3 - exact).max() < 1e-3


def test_l1_normalize_restatement_matches_sklearn():
    """sklearn is installed here: pin the restated arithmetic of normalize(norm='l1') (wavelet/src/utils.py:112)."""
    from sklearn.preprocessing import normalize
    m = sp.random(300, 300, density=0.1, random_state=1, format="csr", dtype=np.float32)
    m.data -= np.float32(0.3)
    m[7, :] = 0
    m.eliminate_zeros()
    want = normalize(m, norm="l1", axis=1)
    got = oracle.l1_normalize_rows(m)
    np.testing.assert_array_equal(got.indices, want.indices)
    np.testing.assert_array_equal(got.data, want.data)


# ---- NAFS aggregator: the restatement against the reference's own outputs ------------
def test_nafs_combine_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN_DIR, "reference_ext.npz"))
    out = oracle.nafs_combine(list(g["nafs2_feats"]))
    np.testing.assert_allclose(out, g["nafs2_out"], rtol=1e-5, atol=1e-6)
    # the propagated list: hops rebuilt with the (pinned) oracle hop
    x = g["nafs_x"]
    hops, _ = oracle.propagate(sym_graph(300, 1500, 2), x, 3, r=0.5)
    np.testing.assert_array_equal(hops[3], g["nafs_hop3"])
    np.testing.assert_allclose(oracle.nafs_combine(hops), g["nafs_out"], rtol=1e-5, atol=1e-6)


# ---- the two PPR-approximation normalisers of directed graphs: oracle pinned, device path not built yet -----------
def _csr_close(got, want, rtol=2e-6):
    got = got.tocsr()
    got.sort_indices()
    np.testing.assert_array_equal(got.indptr, want.indptr)
    np.testing.assert_array_equal(got.indices, want.indices)
    np.testing.assert_allclose(got.data, want.data, rtol=rtol, atol=1e-9)


def test_ppr_approx_normalisers_match_reference_golden():
    g = np.load(os.path.join(GOLDEN_DIR, "reference_ext.npz"))
    a = golden_csr(g, "ppr_adj")
    x = g["ppr_x"]
    _csr_close(oracle.fast_ppr_norm(a, 0.5, 0.1), golden_csr(g, "fastppr_norm"))
    one, two = oracle.two_order_ppr_norm(a, 0.5, 0.1)
    _csr_close(one, golden_csr(g, "twoorder_one"))
    _csr_close(two, golden_csr(g, "twoorder_two"))
    # the hop chains on the reference's own matrices are the pinned FMA chain: bit-exact
    for adj_key, hops_key in (("fastppr_norm", "fastppr_hops"), ("twoorder_one", "twoorder_one_hops"),
                              ("twoorder_two", "twoorder_two_hops")):
        m = golden_csr(g, adj_key)
        cur = x
        for k in range(1, 3):
            cur = oracle.spmm_hop(m, cur)
            np.testing.assert_array_equal(cur, g[hops_key][k])
