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
