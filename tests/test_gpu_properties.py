"""Property tests (hypothesis) on the GPU: arbitrary small CSR matrices — empty rows, hubs, explicit
zeros, negative weights, any feature width — against the oracle.  SURVEY.md §4."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle
from helpers import assert_same_structure, ulp_diff64
from scalable_roubust_gnn_b200 import device as dev
from scalable_roubust_gnn_b200.operators import SymLaplacianGraphOp, adj_to_symmetric_norm, csr_sparse_dense_matmul

pytestmark = pytest.mark.gpu
SETTINGS = dict(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))


@st.composite
def csr_and_features(draw, square=True):
    n = draw(st.integers(1, 60))
    f = draw(st.integers(1, 140))
    density = draw(st.sampled_from([0.0, 0.02, 0.1, 0.5]))
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    m = sp.random(n, n, density=density, random_state=seed % (2**31), format="csr", dtype=np.float64)
    m.data = rng.standard_normal(len(m.data))
    if draw(st.booleans()) and n > 2:                 # a hub row and an empty row
        dense = m.toarray()
        dense[0, :] = rng.standard_normal(n)
        dense[1, :] = 0
        m = sp.csr_matrix(dense)
    m.sort_indices()
    x = rng.standard_normal((n, f)).astype(np.float32)
    return m, x


@settings(**SETTINGS)
@given(csr_and_features())
def test_hop_bit_exact_on_arbitrary_csr(case):
    m, x = case
    want = oracle.spmm_hop(m, x)
    np.testing.assert_array_equal(csr_sparse_dense_matmul(m, x), want)
    a = dev.upload_csr(m.astype(np.float32))
    xp = dev.pack_features(torch.from_numpy(x).cuda())
    got = dev.unpack_features(dev.spmm(a, xp, x.shape[1]), x.shape[1]).cpu().numpy()
    np.testing.assert_array_equal(got, want)


@st.composite
def adjacency(draw):
    n = draw(st.integers(1, 80))
    seed = draw(st.integers(0, 2**31 - 1))
    kind = draw(st.sampled_from(["sym_unweighted", "sym_weighted", "directed", "with_loops", "with_zeros"]))
    rng = np.random.default_rng(seed)
    m = sp.random(n, n, density=draw(st.sampled_from([0.0, 0.05, 0.3])), random_state=seed % (2**31), format="csr")
    if kind == "sym_unweighted":
        m = ((m + m.T) > 0).astype(np.float64)
    elif kind == "sym_weighted":
        m = m.maximum(m.T)
    elif kind == "with_loops":
        m = ((m + m.T) > 0).astype(np.float64) + sp.diags((rng.random(n) > 0.5).astype(np.float64))
    elif kind == "with_zeros":
        m = m.maximum(m.T).tocsr()
        if m.nnz:
            # zero a symmetric pair so the matrix stays symmetric
            coo = m.tocoo()
            i, j = int(coo.row[0]), int(coo.col[0])
            data = m.data.copy()
            rows = np.repeat(np.arange(n), np.diff(m.indptr))
            data[((rows == i) & (m.indices == j)) | ((rows == j) & (m.indices == i))] = 0.0   # explicit zeros
            m = sp.csr_matrix((data, m.indices, m.indptr), shape=m.shape)
    m = sp.csr_matrix(m, dtype=np.float64)
    m.sort_indices()
    r = draw(st.sampled_from([0.0, 0.5, 1.0]))
    return m, r


@settings(**SETTINGS)
@given(adjacency())
def test_normalisation_matches_oracle_on_arbitrary_adjacency(case):
    adj, r = case
    want = oracle.sym_norm(adj, r)
    got = adj_to_symmetric_norm(adj, r)
    assert_same_structure(got, want)
    if want.nnz:
        assert ulp_diff64(got.data, want.data).max() <= 4


@settings(max_examples=15, deadline=None, suppress_health_check=list(HealthCheck))
@given(adjacency(), st.integers(1, 70), st.integers(0, 3))
def test_propagate_matches_oracle(case, f, k):
    adj, r = case
    x = np.random.default_rng(f).random((adj.shape[0], f), dtype=np.float32)
    want, _ = oracle.propagate(adj, x, k, r=r)
    got = SymLaplacianGraphOp(k, r=r).propagate(adj, x)
    assert len(got) == k + 1
    for g, w in zip(got, want):
        np.testing.assert_allclose(g.numpy(), w, rtol=1e-5, atol=1e-6)


# ---- the "next" rows: transpose, sparse product, magnetic normalisation, NAFS ------------------
@st.composite
def square_f32(draw):
    n = draw(st.integers(1, 70))
    seed = draw(st.integers(0, 2**31 - 1))
    m = sp.random(n, n, density=draw(st.sampled_from([0.0, 0.03, 0.2, 0.6])), random_state=seed % (2**31), format="csr",
                  dtype=np.float32)
    m.data = np.random.default_rng(seed).standard_normal(len(m.data)).astype(np.float32)
    m.sort_indices()
    return m


@settings(**SETTINGS)
@given(square_f32())
def test_transpose_exact_and_involutive(m):
    from scalable_roubust_gnn_b200.sparse_mm import csr_to_scipy, csr_transpose, scipy_sparse_mat_to_device_adj
    d = scipy_sparse_mat_to_device_adj(m).csr
    t = csr_transpose(d)
    want = m.T.tocsr()
    want.sort_indices()
    got = csr_to_scipy(t)
    np.testing.assert_array_equal(got.indptr, want.indptr)
    np.testing.assert_array_equal(got.indices, want.indices)
    np.testing.assert_array_equal(got.data, want.data)
    back = csr_to_scipy(csr_transpose(t))
    np.testing.assert_array_equal(back.indices, m.indices)
    np.testing.assert_array_equal(back.data, m.data)


@settings(**SETTINGS)
@given(square_f32(), st.integers(0, 2**31 - 1))
def test_spgemm_matches_scipy_pattern_and_values(a, seed):
    from scalable_roubust_gnn_b200.sparse_mm import csr_to_scipy, scipy_sparse_mat_to_device_adj, spgemm
    n = a.shape[0]
    b = sp.random(n, n, density=0.15, random_state=seed % (2**31), format="csr", dtype=np.float32)
    b.sort_indices()
    got = csr_to_scipy(spgemm(scipy_sparse_mat_to_device_adj(a).csr, scipy_sparse_mat_to_device_adj(b).csr))
    # scipy keeps structural entries whose products cancel, exactly like the expand / sort / compress path
    want = (a.astype(np.float64) @ b.astype(np.float64)).tocsr()
    want.sort_indices()
    np.testing.assert_array_equal(got.indptr, want.indptr)
    np.testing.assert_array_equal(got.indices, want.indices)
    # Tolerance, and why it is not 1e-5 / 1e-6: `want` is the float64 product rounded once; the device sums float32
    # products sequentially in ascending k (the order one of the reference's dense sgemm kernels may use,
    # SSRG/operators/utils.py:216-219).  With up to ~n <= 40 signed terms of magnitude <= 1 per entry the float32
    # chain carries an absolute error of up to n * 2^-24 * max|term| ~ 2.4e-6 that does NOT shrink when the terms
    # cancel, hence the absolute floor of 5e-6 (2x that bound); entries that do not cancel agree to 2e-5 relative
    # (n * eps32 = 40 * 6e-8 = 2.4e-6 would do; 2e-5 leaves room for hypothesis' adversarial magnitudes).
    np.testing.assert_allclose(got.data, want.data, rtol=2e-5, atol=5e-6)


@st.composite
def digraph(draw):
    n = draw(st.integers(1, 60))
    seed = draw(st.integers(0, 2**31 - 1))
    m = sp.random(n, n, density=draw(st.sampled_from([0.0, 0.05, 0.3])), random_state=seed % (2**31), format="csr")
    if draw(st.booleans()):
        m.data[:] = 1.0                                        # unweighted: theta in {0, +-1}
    else:
        m.data = np.round(m.data + 0.25, 3)
    m = sp.csr_matrix(m, dtype=np.float64)
    m.sort_indices()
    return m, draw(st.sampled_from([0.0, 0.3, 0.5, 1.0])), draw(st.sampled_from([0.0, 0.1, 0.25]))


@settings(**SETTINGS)
@given(digraph(), st.booleans())
def test_magnetic_norm_matches_oracle_on_arbitrary_digraphs(case, ppr):
    from scalable_roubust_gnn_b200.operators import adj_to_directed_symmetric_mag_norm
    adj, r, q = case
    alpha = 0.15 if ppr else None
    want_re, want_im = oracle.mag_norm(adj, r, q, alpha)
    got_re, got_im = adj_to_directed_symmetric_mag_norm(adj, r, q, ppr_alpha=alpha)
    for got, want in ((got_re, want_re), (got_im, want_im)):
        assert_same_structure(got, want)
        if want.nnz:
            # Tolerance (float64): every value is a product of four factors - d_u^(r-1), the symmetrised weight,
            # d_v^(-r) and cos / sin(2 pi q theta).  The oracle takes pow / cos / sin from the host libm, the kernel from
            # CUDA's (documented <= 2 ulp for pow, <= 2 ulp for sincos); neither is correctly rounded, so the two may
            # differ by up to ~(2 + 2 + 2) ulp of the factors plus 3 roundings of the products: <= ~10 ulp = 2.2e-15
            # relative.  1e-14 is 4-5x that bound; the absolute floor 1e-15 covers sin(2 pi q theta) values that
            # should be exactly 0 and come out as ~1e-16 in one of the two libraries.
            np.testing.assert_allclose(got.data, want.data, rtol=1e-14, atol=1e-15)


@settings(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck))
@given(st.integers(1, 50), st.integers(1, 150), st.integers(1, 9), st.integers(0, 2**31 - 1))
def test_nafs_properties(n, f, hops, seed):
    """weights are a softmax (positive, sum 1), the output is their convex combination row by row, identical hop
    matrices give uniform weights, and the result matches the restatement."""
    from scalable_roubust_gnn_b200.operators.message_operator import nafs_combine_device
    rng = np.random.default_rng(seed)
    feats = [rng.standard_normal((n, f)).astype(np.float32) for _ in range(hops)]
    devs = [dev.pack_features(torch.from_numpy(x).cuda()) for x in feats]
    out, w = nafs_combine_device(devs, f=f, want_weights=True)
    w = w.cpu().numpy()
    assert (w > 0).all()
    np.testing.assert_allclose(w.sum(1), 1.0, atol=2e-6)
    want, w_want = oracle.nafs_combine(feats, return_weights=True)
    # Tolerance (north star: 1e-5 relative / 1e-6 absolute for propagated features; this is the aggregator on top).
    # The softmax weights come from float32 dot products over up to 150 columns reduced in a different order than
    # the oracle's (butterfly over lanes vs sequential): relative error ~sqrt(f) * eps32 ~ 7e-7 on the cosine, which
    # exp() passes through unamplified (|score| <= 1), and up to 9 hops are normalised by their sum => a few 1e-6
    # relative on a weight; standard-normal inputs make the combination's terms cancel, so the output carries the
    # same few 1e-6 as an ABSOLUTE error (|x| ~ 1).  2e-5 / 2e-6 is ~4x the estimate; the reference's own torch
    # kernels (vectorised exp / norm) are not bit-reproducible across builds either.
    np.testing.assert_allclose(w, w_want, rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(out[:, :f].cpu().numpy(), want, rtol=2e-5, atol=2e-6)
    same = [devs[0]] * hops
    _, wu = nafs_combine_device(same, f=f, want_weights=True)
    np.testing.assert_allclose(wu.cpu().numpy(), 1.0 / hops, rtol=1e-6)
