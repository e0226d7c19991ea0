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


# ------------------------------------------------------------------------------------------------
# counter-based generators shared with the device (csrc/coo.cu: srg_synth_rmat_shard_csr,
# srg_synth_hash_features_f32).  Pure integer arithmetic: the numpy forms below reproduce the device
# output bit for bit and define the graph of BASELINE.json's configs 4 (power-law) / 5 at any scale.
# ------------------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return x ^ (x >> np.uint64(31))


def _scramble(x, scale, seed):
    mask = np.uint64((1 << scale) - 1)
    h = np.uint64(scale // 2 + 1)
    with np.errstate(over="ignore"):
        x = (x * np.uint64(0x9E3779B97F4A7C15) + _splitmix64(np.uint64(seed))) & mask
        x = x ^ (x >> h)
        x = (x * np.uint64(0xD6E8FEB86659FD93)) & mask
        x = x ^ (x >> h)
        x = (x * np.uint64(0xCA5A826395121157)) & mask
        x = x ^ (x >> h)
    return x


def rmat_scale(n):
    return max(1, int(np.ceil(np.log2(max(int(n), 2)))))


def rmat_draws(n, nnz, factor=1.08):
    """Edge ids to draw so that about nnz/2 undirected pairs survive the id rejection (ids >= n) and
    the duplicate collapse of R-MAT."""
    acc = (n / float(1 << rmat_scale(n))) ** 2
    return int(nnz / 2 / acc * factor)


def rmat_scrambled_edges_host(n, m_draw, seed=0, abc=(0.57, 0.19, 0.19), e0=0):
    """(u, v) of edge ids e0 .. e0+m_draw-1 before rejection (uint64 arrays)."""
    scale = rmat_scale(n)
    a, b, c = abc
    ta, tab, tabc = (np.uint64(int(p * 4294967296.0)) for p in (a, a + b, a + b + c))
    e = np.arange(e0, e0 + m_draw, dtype=np.uint64)
    with np.errstate(over="ignore"):
        base = _splitmix64(np.uint64(seed) ^ (e * np.uint64(0xA24BAED4963EE407)))
    u = np.zeros(m_draw, dtype=np.uint64)
    v = np.zeros(m_draw, dtype=np.uint64)
    bits = None
    for level in range(scale):
        if level % 2 == 0:
            with np.errstate(over="ignore"):
                bits = _splitmix64(base + np.uint64(level >> 1))
        r = (bits >> np.uint64(32)) if level % 2 else (bits & np.uint64(0xFFFFFFFF))
        right = ((r >= ta) & (r < tab)) | (r >= tabc)
        down = r >= tab
        u = (u << np.uint64(1)) | down.astype(np.uint64)
        v = (v << np.uint64(1)) | right.astype(np.uint64)
    return _scramble(u, scale, seed), _scramble(v, scale, seed)


def rmat_scrambled_host(n, m_draw, seed=0, abc=(0.57, 0.19, 0.19), dtype=np.float64):
    """The whole symmetric, loop-free, duplicate-free graph of the device generator as a scipy CSR."""
    u, v = rmat_scrambled_edges_host(n, m_draw, seed, abc)
    keep = (u < n) & (v < n) & (u != v)
    u, v = u[keep].astype(np.int64), v[keep].astype(np.int64)
    key = _unique_sorted(np.concatenate([u * n + v, v * n + u]))
    return _csr_from_sorted_keys(key, n, dtype)


def hash_features_host(seed, row0, n_rows, col0, f, f_total, rows=None):
    """float32 U[0,1) features of rows [row0, row0+n_rows) (or the explicit global ``rows``), columns
    [col0, col0+f) of the N x f_total matrix."""
    rows = (np.arange(row0, row0 + n_rows, dtype=np.uint64) if rows is None else np.asarray(rows, dtype=np.uint64))[:, None]
    cols = np.arange(col0, col0 + f, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        h = _splitmix64(np.uint64(seed) ^ ((rows * np.uint64(f_total) + cols) * np.uint64(0x9E3779B97F4A7C15)))
    return ((h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def rmat_shard_device(n, m_draw, row0, row1, seed=0, abc=(0.57, 0.19, 0.19), cap=None, device="cuda"):
    """Rows [row0, row1) of rmat_scrambled_host(n, m_draw, seed) built on the GPU: DeviceCSR with
    all-ones values (data=None), global column ids."""
    import ctypes as C

    import torch

    from . import _lib
    from .device import DeviceCSR
    lib = _lib.load()
    n_loc = row1 - row0
    if cap is None:        # expected share of the 2 * m_draw directed entries, with head-room
        acc = (n / float(1 << rmat_scale(n))) ** 2
        cap = int(2.0 * m_draw * acc * n_loc / max(n, 1) * 1.4) + 4096
    cap = min(cap, 2**31 - 1)
    indptr = torch.empty(n_loc + 1, dtype=torch.int32, device=device)
    indices = torch.empty(max(cap, 1), dtype=torch.int32, device=device)
    nnz = C.c_int64(0)
    a, b, c = abc
    stream = C.c_void_p(torch.cuda.current_stream(indptr.device).cuda_stream)
    _lib.check(lib.srg_synth_rmat_shard_csr(C.c_uint64(seed), rmat_scale(n), int(m_draw), a, b, c, int(n), int(row0),
                                            int(row1), int(cap), C.c_void_p(indptr.data_ptr()),
                                            C.c_void_p(indices.data_ptr()), C.byref(nnz), stream))
    return DeviceCSR(indptr, indices, None, n_loc, int(nnz.value))


def hash_features_device(n_rows, f, row0=0, col0=0, f_total=None, seed=1, ld=None, device="cuda"):
    """Device twin of hash_features_host in the padded layout (n_rows x ld, pad columns zero)."""
    import ctypes as C

    import torch

    from . import _lib
    from .device import pad_ld
    lib = _lib.load()
    ld = pad_ld(f) if ld is None else ld
    f_total = f if f_total is None else f_total
    out = torch.empty((n_rows, ld), dtype=torch.float32, device=device)
    stream = C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)
    _lib.check(lib.srg_synth_hash_features_f32(C.c_uint64(seed), int(row0), int(n_rows), int(col0), int(f), int(f_total),
                                               C.c_void_p(out.data_ptr()), ld, stream))
    return out
