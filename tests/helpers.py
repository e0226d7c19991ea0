"""Shared test helpers: golden CSR unpacking, synthetic graphs, ulp distance."""
import numpy as np
import scipy.sparse as sp


def golden_csr(g, prefix, n=None):
    indptr = g[prefix + "_indptr"]
    n = len(indptr) - 1 if n is None else n
    return sp.csr_matrix((g[prefix + "_data"], g[prefix + "_indices"], indptr), shape=(n, n))


def ulp_diff64(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64).view(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float64).view(np.int64)
    return np.abs(a - b)


def ulp_diff32(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def sym_graph(n, m, seed, weighted=False, dtype=np.float64):
    """G(N, nnz, seed, uniform): m random pairs, self pairs dropped, symmetrised, coalesced."""
    rng = np.random.default_rng(seed)
    u, v = rng.integers(0, n, m), rng.integers(0, n, m)
    keep = u != v
    u, v = u[keep], v[keep]
    w = rng.random(len(u)) + 0.5 if weighted else np.ones(len(u))
    a = sp.coo_matrix((w, (u, v)), shape=(n, n)).tocsr()
    a = a.maximum(a.T).tocsr()
    a.sort_indices()
    return a.astype(dtype)


def assert_same_structure(a, b):
    a, b = a.tocsr(), b.tocsr()
    assert a.shape == b.shape
    np.testing.assert_array_equal(a.indptr, b.indptr)
    np.testing.assert_array_equal(a.indices, b.indices)
