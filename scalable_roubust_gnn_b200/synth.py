"""Synthetic inputs of the named shapes — host-side generators, fixed seeds.

G(N, nnz, seed, uniform): nnz/2 random pairs, self pairs dropped, symmetrised, duplicates
coalesced, all weights 1.0.  `nnz` is the stored-entry count of the symmetric adjacency WITHOUT
self loops (Cora 10 556 = 2 x 5 278, PubMed 88 648 = 2 x 44 324).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

#: name -> (N, nnz(A) symmetric without loops, F, K)   (BASELINE.json configs 1-5)
SHAPES = {
    "cora": (2708, 10556, 1433, 3),
    "pubmed": (19717, 88648, 500, 5),
    "arxiv": (169343, 1166243, 128, 3),
    "products": (2449029, 61859140, 100, 3),
    "papers100M": (111059956, 1615685872, 128, 3),
    "papers100M_16th": (111059956 // 16, 1615685872 // 16, 128, 3),   # 1/16-scale stand-in
}


def _csr_from_sorted_keys(key, n, dtype):
    rows = key // n
    cols = (key - rows * n).astype(np.int32)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=indptr[1:])
    a = sp.csr_matrix((np.ones(len(cols), dtype=dtype), cols, indptr.astype(np.int32)), shape=(n, n))
    a.has_sorted_indices = True
    return a


def _unique_sorted(key):
    key.sort()
    if len(key) == 0:
        return key
    keep = np.empty(len(key), dtype=bool)
    keep[0] = True
    np.not_equal(key[1:], key[:-1], out=keep[1:])
    return key[keep]


def uniform_graph(n, nnz, seed=0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    m = nnz // 2
    u = rng.integers(0, n, m, dtype=np.int64)
    v = rng.integers(0, n, m, dtype=np.int64)
    keep = u != v
    u, v = u[keep], v[keep]
    key = _unique_sorted(np.concatenate([u * n + v, v * n + u]))
    return _csr_from_sorted_keys(key, n, dtype)


def rmat_graph(n, nnz, seed=0, abcd=(0.57, 0.19, 0.19, 0.05), dtype=np.float64):
    """Power-law R-MAT graph (a,b,c,d), ids >= n rejected, symmetrised + coalesced."""
    rng = np.random.default_rng(seed)
    scale = int(np.ceil(np.log2(max(n, 2))))
    m = nnz // 2
    a, b, c, _ = abcd
    u = np.zeros(m, dtype=np.int64)
    v = np.zeros(m, dtype=np.int64)
    for _ in range(scale):
        p = rng.random(m)
        right = (p >= a) & (p < a + b) | (p >= a + b + c)      # quadrant b or d: column bit set
        down = p >= a + b                                       # quadrant c or d: row bit set
        u = (u << 1) | down
        v = (v << 1) | right
    keep = (u < n) & (v < n) & (u != v)
    u, v = u[keep], v[keep]
    key = _unique_sorted(np.concatenate([u * n + v, v * n + u]))
    return _csr_from_sorted_keys(key, n, dtype)


def features(n, f, seed=1):
    """X = default_rng(seed).random((N, F), float32)  — U[0, 1)."""
    return np.random.default_rng(seed).random((n, f), dtype=np.float32)
