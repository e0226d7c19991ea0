"""oracle — CPU restatement of the reference's propagation path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package, and only as the checker or the timed CPU baseline.  The product
(``scalable_roubust_gnn_b200``) never imports it and has no CPU fallback.

What it restates (file:line in /root/reference, SSRG = "Scalable Spectral Robust GNN"):
  sym_norm          SSRG/operators/utils.py:81-93  adj_to_symmetric_norm, driven as in
                    SSRG/operators/graph_operator/symmetrical_simgraph_laplacian_operator.py:12-15
                    and .../symmetrical_simgraph_ppr_operator.py:13-21
  spmm_hop          SSRG/operators/utils.py:17-47 + SSRG/operators/csrc/matmul.c:23-40
  propagate         SSRG/operators/base_operator.py:19-36
  feature_mask / edge_mask / symmetrize_edges
                    SSRG/data_process.py:35-41, :43-67, SSRG/data_augument.py:28, :99-102
  cheby_*           pygsp 0.5.1 (PyPI "PyGSP", un-pinned and absent from /root/reference):
                    call sites wavelet/src/utils.py:83,95,131-133 and
                    SSRG/models/base_scalable/base_model.py:184-189,243  -- PARITY UNPINNED
  two_dir_norm      SSRG/operators/utils.py:195-260
  fast_ppr_norm / two_order_ppr_norm   SSRG/operators/utils.py:262-335, :337-424 (device paths not built yet)
  mag_norm / com_propagate  SSRG/operators/utils.py:95-138, SSRG/operators/base_operator.py:152-208, :316-345
  spectral_preprocess  SSRG/models/base_scalable/base_model.py:180-221 (on top of cheby_*: PARITY UNPINNED)
  nafs_combine      SSRG/operators/message_operator/over_smooth_distance_op.py:11-33
  row_partition     new functionality (no reference code): the bit-exact partition map

Pinning: tests/test_oracle.py checks these against tests/golden/*.npz (outputs of the real
reference package imported in the build container by tests/golden/make_golden.py) and against
oracle/_ref/libmatmul_ref.so (the reference's matmul.c compiled in place by oracle/Makefile).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libmatmul_ref.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, ndim=1, flags="CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, ndim=1, flags="CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, ndim=1, flags="CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", _HERE] + ([] if ref else ["liboracle.so"]), check=True, capture_output=True)


_oracle_lib = None
_ref_lib = None


def _lib():
    global _oracle_lib
    if _oracle_lib is None:
        if not os.path.exists(_ORACLE_SO):
            build(ref=False)
        lib = C.CDLL(_ORACLE_SO)
        lib.oracle_spmm_csr_f32.argtypes = [_f32p, C.c_int64, _f32p, _i32p, _i32p, _f32p, C.c_int64, C.c_int64, C.c_int32]
        lib.oracle_spmm_csr_f32.restype = None
        lib.oracle_spmm_csr_f64.argtypes = [_f64p, C.c_int64, _f64p, _i32p, _i32p, _f64p, C.c_int64, C.c_int64, C.c_int32]
        lib.oracle_spmm_csr_f64.restype = None
        _oracle_lib = lib
    return _oracle_lib


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


def ref_lib():
    """The reference's own matmul.c, compiled (oracle/Makefile).  None if it was never built."""
    global _ref_lib
    if _ref_lib is None and have_ref():
        lib = C.CDLL(_REF_SO)
        lib.FloatCSRMulDenseOMP.argtypes = [_f32p, _f32p, _i32p, _i32p, _f32p, C.c_int, C.c_int]
        lib.FloatCSRMulDenseOMP.restype = None
        _ref_lib = lib
    return _ref_lib


# ------------------------------------------------------------------------------------------------
# normalisation
# ------------------------------------------------------------------------------------------------
def sym_norm(adj, r, ppr_alpha=None):
    """R = D^(r-1) (A+I)^T D^(-r) as a canonical CSR (int32 indices, float64 data).

    Restated entry-wise instead of as the reference's chain of diagonal products:
        A~ = A + I (duplicates summed, exact zeros dropped);  d = row sums of A~ (fp64)
        dl = d**(r-1), dr = d**(-r), inf -> 0
        R[a, b] = (A~[b, a] * dl[a]) * dr[b]          (this multiply order, fp64)
        exact zeros dropped (scipy's csr_matmat omits them)
        PPR:  R <- (1 - alpha) * R;  R[a, a] += alpha
    """
    n = adj.shape[0]
    a_tilde = (sp.csr_matrix(adj, dtype=np.float64) + sp.identity(n, dtype=np.float64, format="csr")).tocsr()
    a_tilde.sum_duplicates()
    d = np.asarray(a_tilde.sum(axis=1)).reshape(-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        dl = np.power(d, r - 1)
        dr = np.power(d, -r)
    dl[np.isinf(dl)] = 0.0
    dr[np.isinf(dr)] = 0.0
    t = a_tilde.transpose().tocsr()  # t[a, b] = A~[b, a]; structure-only operation
    t.sort_indices()
    rows = np.repeat(np.arange(n), np.diff(t.indptr))
    vals = (t.data * dl[rows]) * dr[t.indices]
    out = sp.csr_matrix((vals, t.indices.astype(np.int32), t.indptr.astype(np.int32)), shape=(n, n))
    out.eliminate_zeros()
    if ppr_alpha is not None:
        out.data = (1 - ppr_alpha) * out.data
        diag_pos = out.indices == np.repeat(np.arange(n), np.diff(out.indptr))
        # A^ has a full diagonal whenever d > 0; rows whose diagonal vanished get a fresh entry
        has_diag = np.zeros(n, dtype=bool)
        has_diag[out.indices[diag_pos]] = True
        out.data[diag_pos] = out.data[diag_pos] + ppr_alpha
        if not has_diag.all():
            missing = np.flatnonzero(~has_diag)
            extra = sp.csr_matrix((np.full(len(missing), float(ppr_alpha)), (missing, missing)), shape=(n, n))
            out = (out + extra).tocsr()
        out.eliminate_zeros()
        out.sort_indices()
    out.indices = out.indices.astype(np.int32)
    out.indptr = out.indptr.astype(np.int32)
    return out


def selfloop_structure(adj):
    """(indptr, indices, degree) of A~ = A + I: the integer part of the normalisation."""
    n = adj.shape[0]
    a_tilde = (sp.csr_matrix(adj, dtype=np.float64) + sp.identity(n, dtype=np.float64, format="csr")).tocsr()
    a_tilde.sum_duplicates()
    a_tilde.sort_indices()
    d = np.asarray(a_tilde.sum(axis=1)).reshape(-1)
    return a_tilde.indptr.astype(np.int32), a_tilde.indices.astype(np.int32), d


# ------------------------------------------------------------------------------------------------
# propagation
# ------------------------------------------------------------------------------------------------
def spmm_hop(adj_norm, feature, lib="oracle"):
    """One hop in the reference's fp32 arithmetic (utils.py:38-47 marshalling + matmul.c:23-40)."""
    feature = np.ascontiguousarray(feature, dtype=np.float32)
    n, f = feature.shape
    data = adj_norm.data.astype(np.float32)
    indices = np.ascontiguousarray(adj_norm.indices, dtype=np.int32)
    indptr = np.ascontiguousarray(adj_norm.indptr, dtype=np.int32)
    answer = np.zeros(adj_norm.shape[0] * f, dtype=np.float32)   # rectangular row slices allowed
    mat = feature.reshape(-1)
    if lib == "ref":
        rl = ref_lib()
        if rl is None:
            raise RuntimeError("oracle/_ref/libmatmul_ref.so not built")
        assert adj_norm.shape[0] == n, "the reference's matmul.c assumes a square adjacency"
        rl.FloatCSRMulDenseOMP(answer, data, indices, indptr, mat, n, f)
    else:
        _lib().oracle_spmm_csr_f32(answer, f, data, indices, indptr, mat, f, adj_norm.shape[0], f)
    return answer.reshape(adj_norm.shape[0], f)


def propagate(adj, feature, prop_steps, r=0.5, ppr_alpha=None, lib="oracle"):
    """[X, A^X, ..., A^^K X] as float32 arrays (base_operator.py:31-36)."""
    adj_norm = sym_norm(adj, r, ppr_alpha)
    out = [np.ascontiguousarray(feature, dtype=np.float32)]
    for _ in range(prop_steps):
        out.append(spmm_hop(adj_norm, out[-1], lib=lib))
    return out, adj_norm


# ------------------------------------------------------------------------------------------------
# magnetic-Laplacian operators of directed graphs
# ------------------------------------------------------------------------------------------------
def mag_norm(adj, r, q, ppr_alpha=None):
    """adj_to_directed_symmetric_mag_norm (SSRG/operators/utils.py:95-138) in fp64 numpy, and the PPR blend
    of SymDirMagComPprGraphOp.construct_adj (symmetrical_directed_magnetic_comppr_operator.py:33-38).
    torch_sparse.coalesce(.., "add") = sum of equal (row, col) keys in stored order, output sorted by key;
    torch_scatter.scatter_add = sequential sum in stored order (both packages absent from /root/reference:
    their documented behaviour is restated).  Returns (real_csr, imag_csr) with float64 data."""
    coo = sp.coo_matrix(adj)
    n = coo.shape[0]
    w = coo.data.astype(np.float64)
    row = np.concatenate([coo.row, coo.col]).astype(np.int64)          # :100
    col = np.concatenate([coo.col, coo.row]).astype(np.int64)
    sym = np.concatenate([w, w])                                       # :102
    theta = np.concatenate([w, -w])                                    # :103
    uniq, inv = np.unique(row * n + col, return_inverse=True)          # coalesce :105
    sym_c = np.zeros(len(uniq))
    th_c = np.zeros(len(uniq))
    np.add.at(sym_c, inv, sym)
    np.add.at(th_c, inv, theta)
    sym_c = sym_c / 2                                                  # :108
    rows = np.concatenate([uniq // n, np.arange(n)])                   # loops appended :109-115
    cols = np.concatenate([uniq % n, np.arange(n)])
    sym_all = np.concatenate([sym_c, np.ones(n)])
    th_all = np.concatenate([th_c, np.zeros(n)])                       # :117-119
    deg = np.zeros(n)
    np.add.at(deg, rows, sym_all)                                      # :122
    with np.errstate(divide="ignore"):
        dl = np.power(deg, r - 1)
        dr = np.power(deg, -r)
    dl[np.isinf(dl)] = 0
    dr[np.isinf(dr)] = 0
    angle = (1j * 2 * np.pi * q).imag * th_all                         # :124
    x = dl[rows] * sym_all * dr[cols]                                  # :130
    real = sp.csr_matrix((x * np.cos(angle), (rows, cols)), shape=(n, n))   # :133-136
    imag = sp.csr_matrix((x * np.sin(angle), (rows, cols)), shape=(n, n))
    if ppr_alpha is not None:
        real = ((1 - ppr_alpha) * real + ppr_alpha * sp.eye(n)).tocsr()
        imag = ((1 - ppr_alpha) * imag).tocsr()
    for m in (real, imag):
        m.sort_indices()
    return real, imag


def two_dir_norm(adj, r):
    """adj_to_un_in_out_dir_symmetric_norm (SSRG/operators/utils.py:195-260) in float32 numpy, dense products as
    in the reference (small graphs only).  add_self_loops = one appended (i, i) entry per node.
    Returns (un, in, out) float32 CSR matrices."""
    coo = sp.coo_matrix(adj)
    n = coo.shape[0]
    f32 = np.float32
    row = np.concatenate([coo.row, np.arange(n)]).astype(np.int64)
    col = np.concatenate([coo.col, np.arange(n)]).astype(np.int64)
    w = np.ones(len(row), dtype=f32)                                          # :197-201

    def pw(d, e):
        with np.errstate(divide="ignore"):
            out = np.power(d, f32(e), dtype=f32)
        out[np.isinf(out)] = 0
        return out

    def norm(rows, cols, vals):
        deg = np.zeros(n, dtype=f32)
        np.add.at(deg, rows, vals)                                            # scatter_add, stored order
        return (pw(deg, r - 1)[rows] * vals * pw(deg, -r)[cols]).astype(f32), deg

    un_w, deg = norm(row, col, w)                                             # :203-210
    un = sp.csr_matrix((un_w, (row, col)), shape=(n, n))
    p = pw(deg, -1)[row] * w                                                  # :212-214
    p_dense = np.zeros((n, n), dtype=f32)
    np.add.at(p_dense, (row, col), p)                                         # to_dense sums duplicates :215
    res = []
    for lmat in (p_dense.T @ p_dense, p_dense @ p_dense.T):                   # :216-217
        lmat = np.where(np.isnan(lmat), f32(0), lmat).astype(f32)
        rr, cc = np.nonzero(lmat)                                             # row-major, like torch.nonzero
        vals, _ = norm(rr, cc, lmat[rr, cc])
        res.append(sp.csr_matrix((vals, (rr, cc)), shape=(n, n)))
    return un, res[0], res[1]


def _f32_sym_scale(rows, cols, vals, n, r):
    """float32 `D^(r-1) L D^(-r)` with D = scatter_add row sums in stored order and inf -> 0
    (the block repeated at SSRG/operators/utils.py:204-210, :229-237, :324-332, :392-400, :413-421)."""
    f32 = np.float32
    vals = np.asarray(vals, dtype=f32)
    deg = np.zeros(n, dtype=f32)
    np.add.at(deg, rows, vals)
    with np.errstate(divide="ignore", invalid="ignore"):
        dl = np.power(deg, f32(r - 1), dtype=f32)
        dr = np.power(deg, f32(-r), dtype=f32)
    dl[np.isinf(dl)] = 0
    dr[np.isinf(dr)] = 0
    return (dl[rows] * vals * dr[cols]).astype(f32)


def _loops_appended(adj):
    """(row, col) of the adjacency pattern with one (i, i) entry appended per node (add_self_loops)."""
    coo = sp.coo_matrix(adj)
    n = coo.shape[0]
    return (np.concatenate([coo.row, np.arange(n)]).astype(np.int64),
            np.concatenate([coo.col, np.arange(n)]).astype(np.int64), n)


def fast_ppr_norm(adj, r, ppr_alpha, max_iter=100, tol=1e-6):
    """adj_to_fast_ppr_approx_symmetric_norm (SSRG/operators/utils.py:262-335): stationary distribution of the
    teleporting walk by fixed-point iteration (fp64, :278-296), L = (Pi^1/2 P Pi^-1/2 + Pi^-1/2 P^T Pi^1/2) / 2
    (:297-301), values cast to float32 and degree-normalised (:304-334)."""
    row, col, n = _loops_appended(adj)
    a1 = sp.csr_matrix((np.ones(len(row)), (row, col)), shape=(n, n))            # duplicates (old loops) summed
    deg = np.asarray(a1.sum(axis=1)).reshape(-1)
    nz = deg.nonzero()[0]
    d_inv = sp.csr_matrix((1 / deg[nz], (nz, nz)), shape=(n, n))
    s = np.full((n, 1), 1 / (1 + ppr_alpha) / n)
    z = ((ppr_alpha * (1 + ppr_alpha)) * (deg != 0)
         + ((1 - ppr_alpha) / (1 + ppr_alpha) + ppr_alpha * (1 + ppr_alpha)) * (deg == 0))[None, :]
    w = (1 - ppr_alpha) * a1.T @ d_inv
    x, old, it = s, np.zeros((n, 1)), 0
    while np.sqrt(((x - old) ** 2).sum()) > tol:
        old = x
        x = w @ x + s @ (z @ x)
        it += 1
        if it >= max_iter:
            break
    pi = (x / x.sum()).reshape(-1)
    p = d_inv @ a1
    with np.errstate(divide="ignore", invalid="ignore"):
        half, inv_half = sp.diags(np.power(pi, 0.5)), sp.diags(np.power(pi, -0.5))
        lap = ((half @ p @ inv_half + inv_half @ p.T @ half) / 2.0).tocsr()
    lap.data[np.isnan(lap.data)] = 0.0
    lap = lap.tocoo()
    vals = _f32_sym_scale(lap.row.astype(np.int64), lap.col.astype(np.int64), lap.data.astype(np.float32), n, r)
    out = sp.csr_matrix((vals, (lap.row, lap.col)), shape=(n, n))
    out.sort_indices()
    return out


def two_order_ppr_norm(adj, r, ppr_alpha):
    """adj_to_slow_first_second_ppr_approx_symmetric_norm (SSRG/operators/utils.py:337-424), dense as there:
    P = D^-1 (A + I) in float32 (:345-352), stationary vector = left eigenvector of the (N+1) x (N+1) teleport
    matrix by LAPACK (:353-369), first-order L from pi (:374-386), second-order L from P^T P and P P^T masked by each
    other (:402-411), both degree-normalised in float32.  Returns (one_order, two_order) float32 CSR."""
    import scipy.linalg
    f32 = np.float32
    row, col, n = _loops_appended(adj)
    deg = np.zeros(n, dtype=f32)
    np.add.at(deg, row, f32(1))
    with np.errstate(divide="ignore"):
        dinv = np.power(deg, f32(-1), dtype=f32)
    dinv[np.isinf(dinv)] = 0
    p = np.zeros((n, n), dtype=f32)
    np.add.at(p, (row, col), dinv[row])
    pv = np.zeros((n + 1, n + 1), dtype=f32)
    pv[:n, :n] = f32(1 - ppr_alpha) * p
    pv[n, :n] = f32(1.0 / n)
    pv[:n, n] = f32(ppr_alpha)
    ev, left = scipy.linalg.eig(pv, left=True, right=False)
    pi = left.real[:, np.argsort(-ev.real, kind="stable")[0]][:n]
    pi = pi / pi.sum()
    assert not (pi < 0).any()
    pi = pi.astype(f32)                       # sgeev on the float32 matrix: everything below stays float32
    with np.errstate(divide="ignore"):
        inv_half = np.power(pi, f32(-0.5), dtype=f32)
        half = np.power(pi, f32(0.5), dtype=f32)
    inv_half[np.isinf(inv_half)] = 0
    half[np.isinf(half)] = 0
    one = ((np.diag(half) @ p) @ np.diag(inv_half) + (np.diag(inv_half) @ p.T) @ np.diag(half)) / f32(2.0)
    one[np.isnan(one)] = 0

    def finish(lmat):
        rr, cc = np.nonzero(lmat)
        vals = _f32_sym_scale(rr, cc, lmat[rr, cc].astype(f32), n, r)
        m = sp.csr_matrix((vals, (rr, cc)), shape=(n, n))
        m.sort_indices()
        return m

    l_in, l_out = p.T @ p, p @ p.T
    m_in, m_out = l_in == 0, l_out == 0       # the reference masks in place: L_in by (L_out == 0), then L_out by the
    l_in[m_out] = 0                           # ALREADY MASKED L_in == 0 (:405-406, same arrays)
    l_out[l_in == 0] = 0
    two = (l_in + l_out) / f32(2.0)
    two[np.isnan(two)] = 0
    return finish(one), finish(two)


def com_propagate(real_adj, imag_adj, feature, prop_steps, lib="oracle"):
    """ComGraphOp.propagate (SSRG/operators/base_operator.py:152-208) including its aliasing: the
    in-place `+=` of calculate_real_imag_feat (:316-345) turns the first real and the first imaginary
    term of every step into the running totals, and those arrays are the inputs of the next step.
    Returns (real_list, imag_list) of float32 arrays, prop_steps + 1 entries each."""
    x = np.ascontiguousarray(feature, dtype=np.float32)
    real_list, imag_list = [x], [x]
    terms = []                                   # (value, r_step, i_step)
    for step in range(prop_steps):
        if step == 0:
            tr = spmm_hop(real_adj, x, lib=lib)
            ti = spmm_hop(imag_adj, x, lib=lib)
            terms = [[tr, 1, 0], [ti, 0, 1]]
            real_list.append(tr)
            imag_list.append(ti)
            continue
        out = []
        for val, rs, is_ in terms:
            out.append([spmm_hop(real_adj, val, lib=lib), rs + 1, is_])
        for val, rs, is_ in terms:
            v = spmm_hop(imag_adj, val, lib=lib)
            if (is_ + 1) & 1 == 0:               # reversal(): i_step even and non-zero
                v = -v
            out.append([v, rs, is_ + 1])
        reals = [t for t in out if t[2] & 1 == 0]
        imags = [t for t in out if t[2] & 1 == 1]
        assert len(reals) == len(imags)
        for k in range(1, len(reals)):
            reals[0][0] += reals[k][0]           # in place: out[...] now holds the running total
            imags[0][0] += imags[k][0]
        real_list.append(reals[0][0])
        imag_list.append(imags[0][0])
        terms = out
    return real_list, imag_list


# ------------------------------------------------------------------------------------------------
# sparsity masks (torch CPU RNG defines them)
# ------------------------------------------------------------------------------------------------
def feature_mask(shape, rate):
    """`(torch.rand(shape) > r).int()` — data_process.py:38-39 (first RNG draw after seeding)."""
    import torch
    return (torch.rand(shape) > rate).int()


def edge_mask(num_upper_edges, rate):
    """`torch.randperm(E)[int(E*rate):]` — data_process.py:55,65."""
    import torch
    return torch.randperm(num_upper_edges)[int(num_upper_edges * rate):]


def upper_edges(adj):
    """canonical (row-major sorted) edges with col > row — data_process.py:48-53."""
    coo = sp.coo_matrix(adj)
    keep = coo.col > coo.row
    return np.stack([coo.row[keep], coo.col[keep]]).astype(np.int64)


def symmetrize_edges(edge_index, n):
    """undirected, duplicate-free adjacency of an edge list (data_augument.py:99-102) as CSR of ones."""
    e = np.asarray(edge_index, dtype=np.int64)
    both = np.concatenate([e, e[::-1]], axis=1)
    key = np.unique(both[0] * n + both[1])
    row, col = key // n, key % n
    return sp.csr_matrix((np.ones(len(key), dtype=np.float64), (row, col)), shape=(n, n))


# ------------------------------------------------------------------------------------------------
# Chebyshev heat filter — restated from pygsp 0.5.1 (parity unpinned, see module docstring)
# ------------------------------------------------------------------------------------------------
def combinatorial_laplacian(adj):
    """L = D - W with W the symmetrised adjacency nx.Graph(adj) gives (max of the two directions
    for an unweighted graph), pygsp Graph(W).L, lap_type='combinatorial'."""
    w = sp.csr_matrix(adj, dtype=np.float64)
    w = w.maximum(w.T).tocsr()
    d = np.asarray(w.sum(axis=1)).reshape(-1)
    lap = (sp.diags(d) - w).tocsr()
    lap.sort_indices()
    return lap


def estimate_lmax(lap):
    """pygsp Graph.estimate_lmax: 1.01 * largest eigenvalue by ARPACK (tol 5e-3, ncv=min(N,10))."""
    from scipy.sparse.linalg import eigsh
    n = lap.shape[0]
    lmax = eigsh(lap, k=1, tol=5e-3, ncv=min(n, 10), return_eigenvectors=False)[0]
    return float(lmax) * 1.01


def cheby_coeff_heat(tau, lmax, order):
    """compute_cheby_coeff(Heat(G, tau), m=order): N = order + 1 quadrature points."""
    n_q = order + 1
    a1 = a2 = lmax / 2.0
    j = np.arange(n_q)
    theta = np.pi * (j + 0.5) / n_q
    g = np.exp(-tau * (a1 * np.cos(theta) + a2) / lmax)
    return np.array([2.0 / n_q * np.sum(g * np.cos(np.pi * o * (j + 0.5) / n_q)) for o in range(order + 1)])


def cheby_op(lap, coeffs, signal, lmax):
    """pygsp cheby_op for a list of coefficient vectors sharing the T_k (fp64)."""
    coeffs = np.atleast_2d(np.asarray(coeffs, dtype=np.float64))
    a1 = a2 = lmax / 2.0
    x = np.asarray(signal, dtype=np.float64)
    t_old = x
    t_cur = (lap.dot(x) - a2 * x) / a1
    res = [0.5 * c[0] * t_old + c[1] * t_cur for c in coeffs]
    for k in range(2, coeffs.shape[1]):
        t_new = (2.0 / a1) * (lap.dot(t_cur) - a2 * t_cur) - t_old
        res = [rs + c[k] * t_new for rs, c in zip(res, coeffs)]
        t_old, t_cur = t_cur, t_new
    return res


def wavelet_threshold(coeffs, tol):
    """`coeffs[coeffs < tol] = 0` then float32 CSR (wavelet/src/utils.py:98-103)."""
    c = np.array(coeffs, dtype=np.float64, copy=True)
    c[c < tol] = 0
    return sp.csr_matrix(c, dtype=np.float32)


def l1_normalize_rows(m):
    """sklearn.preprocessing.normalize(m, norm='l1', axis=1) for a float32 CSR (wavelet/src/utils.py:106-112).
    sklearn's _inplace_csr_row_normalize_l1 accumulates sum(|v|) sequentially in a C double and stores
    float32(v / sum) when the sum is non-zero (checked against sklearn in tests/test_oracle.py)."""
    m = sp.csr_matrix(m, dtype=np.float32, copy=True)
    for i in range(m.shape[0]):
        s, e = m.indptr[i], m.indptr[i + 1]
        tot = 0.0
        for v in m.data[s:e]:
            tot += abs(float(v))
        if tot != 0.0:
            m.data[s:e] = (m.data[s:e].astype(np.float64) / tot).astype(np.float32)
    return m


def spectral_preprocess(adj, feature, scale, order, tol, lmax, return_phis=False):
    """SpectralModel.preprocess (SSRG/models/base_scalable/base_model.py:180-221): wavelets for the scales
    (-scale, +scale) (:185-192, impulse blocks :236-265 are the same arithmetic as the full identity),
    L1 normalisation (:287-290), then [X | relu((Psi_0 Psi_1) X)] with the sparse-sparse product formed
    first in float32 (spspmm :208-214, spmm :215-219) and the concat of :221."""
    lap = combinatorial_laplacian(adj)
    n = lap.shape[0]
    coeffs = [cheby_coeff_heat(t, lmax, order) for t in (-scale, scale)]
    res = cheby_op(lap, coeffs, np.eye(n), lmax)
    phis = [l1_normalize_rows(wavelet_threshold(r, tol)) for r in res]
    x = np.asarray(feature, dtype=np.float32)
    prod = (phis[0] @ phis[1]).astype(np.float32)
    loc = np.maximum((prod @ x).astype(np.float32), np.float32(0))
    out = np.concatenate([x, loc], axis=1)
    return (out, phis) if return_phis else out


# ------------------------------------------------------------------------------------------------
# NAFS over-smoothing-distance aggregator (SSRG/operators/message_operator/over_smooth_distance_op.py)
# ------------------------------------------------------------------------------------------------
def nafs_combine(feat_list, return_weights=False):
    """OverSmoothDistanceWeightedOp.combine (:11-33) in float32 numpy: cosine score of every hop row
    against the input row (:13-19), softmax over the hops (:22), weighted sum in hop order (:27-31;
    the reference's per-row Python loop is the same arithmetic as these whole-array statements)."""
    feats = [np.asarray(f, dtype=np.float32) for f in feat_list]
    x0 = feats[0]
    eps = np.float32(1e-10)
    norm_fea = np.sqrt((x0 * x0).sum(1, dtype=np.float32)) + eps
    scores = []
    for fea in feats:
        norm_cur = np.sqrt((fea * fea).sum(1, dtype=np.float32)) + eps
        scores.append(((x0 * fea).sum(1, dtype=np.float32) / norm_cur) / norm_fea)
    s = np.stack(scores, axis=1)
    e = np.exp(s - s.max(1, keepdims=True), dtype=np.float32)
    w = e / e.sum(1, keepdims=True, dtype=np.float32)
    out = np.zeros_like(x0)
    for j, fea in enumerate(feats):
        out = out + w[:, j:j + 1] * fea
    return (out, w) if return_weights else out


# ------------------------------------------------------------------------------------------------
# row partition (new functionality; the oracle is its 3-line definition)
# ------------------------------------------------------------------------------------------------
def row_partition(n, world):
    rows_per = -(-n // world)
    starts = np.minimum(np.arange(world + 1, dtype=np.int64) * rows_per, n)
    return rows_per, starts


# ---- dataset augmentation: edge_augument (SSRG/data_augument.py:73-103) ------------------------------------------
def edge_augument(edge_row, edge_col, n, soft_label, degree_level=1, seed=None):
    """CPU restatement of edge_augument (test infrastructure; pinned to outputs of the reference's own function by
    tests/golden/reference_augment.npz).  Follows the reference line by line:
      :75-80  counts = Counter(cat(row, col)); nodes that never occur are appended with count 0
      :81     visiting order = stable sort of the Counter's items by count
      :83-85  stop at the first node whose count reaches degree_level
      :87     100 * deficit candidates: generate_numbers (SSRG/utils.py:29-33) = remove node, random.sample, append node
      :88-92  L2 distance of the soft labels (SSRG/utils.py:35-38, float32), ascending sort
      :93-96  the `deficit` closest candidates become edges (node -> candidate)
      :97-102 cat both directions, torch.unique(dim=1)  (= lexicographic (row, col) order, duplicates dropped)
    `seed`: random.seed(seed) before the first draw (seed_everything), None = use the stream as it is."""
    import random
    from collections import Counter
    edge_row = np.asarray(edge_row, dtype=np.int64)
    edge_col = np.asarray(edge_col, dtype=np.int64)
    soft = np.asarray(soft_label, dtype=np.float32)
    if seed is not None:
        random.seed(seed)
    counts = Counter(np.concatenate([edge_row, edge_col]).tolist())
    for i in range(n):
        if i not in counts:
            counts.update({i: 0})
    src, dst = [edge_row], [edge_col]
    numbers = list(range(n))
    for node, degree in sorted(counts.items(), key=lambda kv: kv[1]):
        if degree >= degree_level:
            break
        deficit = degree_level - degree
        numbers.remove(node)
        cand = random.sample(numbers, deficit * 100)
        numbers.append(node)
        diff = soft[node][None, :] - soft[cand]                        # float32, as repeat_feature - candidates
        dist = np.sqrt((diff.astype(np.float64) ** 2).sum(1)).astype(np.float32)
        order = np.argsort(dist, kind="stable")
        src.append(np.full(deficit, node, dtype=np.int64))
        dst.append(np.asarray(cand, dtype=np.int64)[order[:deficit]])
    r, c = np.concatenate(src), np.concatenate(dst)
    both = np.stack([np.concatenate([r, c]), np.concatenate([c, r])])
    return np.unique(both, axis=1)
