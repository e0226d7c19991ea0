"""Host-buffer helpers of the propagation path, mirroring SSRG/operators/utils.py.

Reference functions mirrored (same names, argument meaning and return layout):
  * ``csr_sparse_dense_matmul(adj, feature)``  — SSRG/operators/utils.py:17-47 (ctypes call into
    libmatmul.so:FloatCSRMulDenseOMP).  Here the same symbol of libsrgnn_b200.so runs the hop on
    the GPU.
  * ``adj_to_symmetric_norm(adj, r)``          — SSRG/operators/utils.py:81-93 (scipy fp64).
    Here: CSR kernels on the GPU; the result comes back as a scipy matrix.
  * ``adj_to_directed_symmetric_mag_norm(adj, r, q)`` — SSRG/operators/utils.py:95-138 (torch +
    torch_sparse + torch_scatter on the CPU).  Here: one key sort + segment kernels (csrc/magnetic.cu).
  * ``propagate_host``                         — the body of GraphOp.propagate
    (SSRG/operators/base_operator.py:31-36) as ONE library call.

Nothing in this module computes on the CPU: numpy/scipy only hold the host buffers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp
import torch

from .. import _lib

__all__ = ["csr_sparse_dense_matmul", "adj_to_symmetric_norm", "propagate_host", "propagate_aggregate_host",
           "csr_host_parts", "adj_to_directed_symmetric_mag_norm", "adj_to_un_in_out_dir_symmetric_norm", "adj_to_fast_ppr_approx_symmetric_norm", "adj_to_slow_first_second_ppr_approx_symmetric_norm", "DeviceHopRunner"]


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.data_ptr()
    return a.ctypes.data


def csr_host_parts(adj):
    """Return (indptr int32, indices int32, data|None, val_dtype, n, nnz) of a scipy CSR without
    copying when the dtypes already match what the C ABI takes."""
    if not isinstance(adj, sp.csr_matrix):
        raise TypeError("The adjacency matrix must be a scipy csr sparse matrix!")
    n = adj.shape[0]
    if adj.shape[0] != adj.shape[1]:
        raise ValueError("The adjacency matrix must be square!")
    nnz = int(adj.indptr[-1]) if n > 0 else 0
    if nnz + n > np.iinfo(np.int32).max:
        raise _lib.SrgError(_lib.SRG_ERR_RANGE, "nnz + n exceeds the int32 CSR range")
    indptr = np.ascontiguousarray(adj.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(adj.indices[:nnz], dtype=np.int32)
    data = adj.data[:nnz]
    if data.dtype == np.float64:
        vt = _lib.SRG_VAL_F64
    elif data.dtype == np.float32:
        vt = _lib.SRG_VAL_F32
    else:  # bool / integer adjacency: scipy's `adj + eye` would upcast to float64
        data = data.astype(np.float64)
        vt = _lib.SRG_VAL_F64
    data = np.ascontiguousarray(data)
    return indptr, indices, data, vt, n, nnz


def _empty(shape, dtype, pin):
    t = torch.empty(shape, dtype=dtype, pin_memory=bool(pin))
    return t


def _use_pinned(pin):
    return bool(pin) and torch.cuda.is_available()


def adj_to_symmetric_norm(adj, r, ppr_alpha=None, device=0, pin=True):
    """D^(r-1) (A+I)^T D^(-r) (and the PPR blend when ``ppr_alpha`` is given) on the GPU.

    Returns a ``scipy.sparse.csr_matrix`` with int32 indices, sorted rows and float64 data — the
    matrix the reference obtains after ``.tocsr()`` (the reference function itself returns CSC).
    """
    lib = _lib.load()
    if sp.issparse(adj) and not isinstance(adj, sp.csr_matrix):
        adj = adj.tocsr()
    return _construct(lib, adj, r, ppr_alpha, device, pin)


def _construct(lib, adj, r, ppr_alpha, device, pin):
    indptr, indices, data, vt, n, nnz = csr_host_parts(adj)
    pin = _use_pinned(pin)
    cap = nnz + n
    o_indptr = _empty((n + 1,), torch.int32, pin)
    o_indices = _empty((max(cap, 1),), torch.int32, pin)
    o_data = _empty((max(cap, 1),), torch.float64, pin)
    o_nnz = C.c_int64(0)
    alpha = -1.0 if ppr_alpha is None else float(ppr_alpha)
    rc = lib.srg_construct_adj_host(_ptr(indptr), _ptr(indices), _ptr(data), vt, n, nnz, float(r), alpha,
                                    _ptr(o_indptr), _ptr(o_indices), _ptr(o_data), C.byref(o_nnz), int(device))
    _lib.check(rc)
    m = int(o_nnz.value)
    out = sp.csr_matrix((o_data.numpy()[:m], o_indices.numpy()[:m], o_indptr.numpy()), shape=(n, n), copy=False)
    out.has_sorted_indices = True
    return out


def csr_sparse_dense_matmul(adj, feature):
    """One propagation hop ``adj @ feature`` in fp32 on the GPU (SSRG/operators/utils.py:17-47).

    ``adj`` is the (already normalised) scipy CSR, ``feature`` a float32 ndarray N x F.  Exactly as
    the reference wrapper, the weights are rounded to float32 (`utils.py:39`) and the result is a
    fresh float32 array of ``feature.shape``.
    """
    lib = _lib.load()
    if not isinstance(feature, np.ndarray) or feature.dtype != np.float32:
        raise C.ArgumentError("feature must be a float32 numpy.ndarray")
    if feature.ndim != 2:
        raise ValueError("feature must be a 2-D array (N x F)")
    # the native symbol reads indptr[mat_row] and mat[indices[j] * mat_col ...]: check what it will index
    if adj.shape[0] != adj.shape[1] or adj.shape[1] != feature.shape[0]:
        raise ValueError("Dimension mismatch detected for the adjacency and the feature matrix!")
    answer = np.zeros(feature.shape, dtype=np.float32).reshape(-1)
    data = np.ascontiguousarray(adj.data, dtype=np.float32)
    indices = np.ascontiguousarray(adj.indices, dtype=np.int32)
    indptr = np.ascontiguousarray(adj.indptr, dtype=np.int32)
    mat = np.ascontiguousarray(feature).reshape(-1)
    mat_row, mat_col = feature.shape
    if lib.srg_device_count() <= 0:
        raise _lib.SrgError(_lib.SRG_ERR_NODEV, "no CUDA device visible: libsrgnn_b200 has no CPU fallback")
    # FloatCSRMulDense is the same shim with an error channel (the literal FloatCSRMulDenseOMP symbol has none and
    # must abort on failure, matmul.h:5); its diagnostic goes through srg_last_error
    if lib.FloatCSRMulDense(_ptr(answer), int(data.size), _ptr(data), _ptr(indices), _ptr(indptr), _ptr(mat), mat_row, mat_col):
        raise _lib.SrgError(_lib.SRG_ERR_CUDA, _lib.last_error() or "FloatCSRMulDense failed")
    return answer.reshape(feature.shape)


def propagate_host(adj, feature, prop_steps, r, ppr_alpha=None, feature_mask=None, device=0, pin=True,
                   return_adj=False):
    """K-hop propagation from host buffers in one library call.

    Returns ``(hops, adj_norm)``: ``hops`` is the list of K float32 CPU tensors (hop 1..K, pinned
    when a GPU is present); ``adj_norm`` is the normalised scipy CSR when ``return_adj`` else None.
    """
    lib = _lib.load()
    return _propagate(lib, adj, feature, prop_steps, r, ppr_alpha, feature_mask, device, pin, return_adj)


def _propagate(lib, adj, feature, K, r, ppr_alpha, feature_mask, device, pin, return_adj):
    indptr, indices, data, vt, n, nnz = csr_host_parts(adj)
    if feature.dtype != np.float32:
        raise C.ArgumentError("feature must be float32 (the reference's ctypes signature rejects other dtypes)")
    feature = np.ascontiguousarray(feature)
    if feature.ndim != 2 or feature.shape[0] != n:
        raise ValueError("Dimension mismatch detected for the adjacency and the feature matrix!")
    F = feature.shape[1]
    mask = None
    if feature_mask is not None:
        mask = feature_mask.numpy() if isinstance(feature_mask, torch.Tensor) else np.asarray(feature_mask)
        if mask.shape != feature.shape:
            raise ValueError("feature_mask must have the shape of the feature matrix")
        mask = np.ascontiguousarray(mask, dtype=np.int32)
    pin = _use_pinned(pin)
    hops = [_empty((n, F), torch.float32, pin) for _ in range(K)]
    hop_ptrs = (C.c_void_p * max(K, 1))(*[h.data_ptr() for h in hops])
    o_indptr = o_indices = o_data = None
    if return_adj:
        cap = nnz + n
        o_indptr = _empty((n + 1,), torch.int32, pin)
        o_indices = _empty((max(cap, 1),), torch.int32, pin)
        o_data = _empty((max(cap, 1),), torch.float64, pin)
    o_nnz = C.c_int64(0)
    alpha = -1.0 if ppr_alpha is None else float(ppr_alpha)
    rc = lib.srg_propagate_host(_ptr(indptr), _ptr(indices), _ptr(data), vt, n, nnz, _ptr(feature), F, _ptr(mask),
                                int(K), float(r), alpha, hop_ptrs, _ptr(o_indptr), _ptr(o_indices), _ptr(o_data),
                                C.byref(o_nnz), int(device))
    _lib.check(rc)
    adj_norm = None
    if return_adj:
        m = int(o_nnz.value)
        adj_norm = sp.csr_matrix((o_data.numpy()[:m], o_indices.numpy()[:m], o_indptr.numpy()), shape=(n, n),
                                 copy=False)
        adj_norm.has_sorted_indices = True
    return hops, adj_norm


def propagate_aggregate_host(adj, feature, prop_steps, r, ppr_alpha, spec, feature_mask=None, device=0, pin=True):
    """One library call: normalisation + K hops + message-operator aggregation; returns the aggregate as
    a CPU float32 tensor (pinned when a GPU is present).  ``spec`` = (agg_mode, start, end, weights)."""
    lib = _lib.load()
    mode, lo, hi, weights = spec
    indptr, indices, data, vt, n, nnz = csr_host_parts(adj)
    feature = np.ascontiguousarray(feature)
    F = feature.shape[1]
    mask = None
    if feature_mask is not None:
        mask = feature_mask.numpy() if isinstance(feature_mask, torch.Tensor) else np.asarray(feature_mask)
        mask = np.ascontiguousarray(mask, dtype=np.int32)
    f_out = F * (hi - lo) if mode == _lib.SRG_AGG_CONCAT else F
    out = _empty((n, f_out), torch.float32, _use_pinned(pin))
    w = None if weights is None else weights.numpy().astype(np.float32)
    alpha = -1.0 if ppr_alpha is None else float(ppr_alpha)
    rc = lib.srg_propagate_aggregate_host(_ptr(indptr), _ptr(indices), _ptr(data), vt, n, nnz, _ptr(feature), F,
                                          _ptr(mask), int(prop_steps), float(r), alpha, int(mode), int(lo), int(hi),
                                          _ptr(w), _ptr(out), int(device))
    _lib.check(rc)
    return out


def adj_to_directed_symmetric_mag_norm(adj, r, q, ppr_alpha=None, device=0):
    """Magnetic-Laplacian normalisation of a directed adjacency on the GPU (SSRG/operators/utils.py:95-138).

    ``adj``: scipy sparse (the reference passes ``adj.tocoo()``).  Returns ``(real, imag)``: two
    ``scipy.sparse.csr_matrix`` with float64 data sharing one sorted pattern (A, A^T and the diagonal).
    ``ppr_alpha`` additionally applies the blend of SymDirMagComPprGraphOp.construct_adj."""
    lib = _lib.load()
    if lib.srg_device_count() <= 0:
        raise _lib.SrgError(_lib.SRG_ERR_NODEV, "no CUDA device visible: libsrgnn_b200 has no CPU fallback")
    if not sp.issparse(adj):
        raise TypeError("The adjacency matrix must be a scipy sparse matrix!")
    csr = adj.tocsr() if not isinstance(adj, sp.csr_matrix) else adj
    indptr, indices, data, vt, n, nnz = csr_host_parts(csr)
    if 2 * nnz + n > np.iinfo(np.int32).max:
        raise _lib.SrgError(_lib.SRG_ERR_RANGE, "2 nnz + n exceeds the int32 CSR range")
    dev = torch.device("cuda", int(device))
    cap = max(2 * nnz + n, 1)
    with torch.cuda.device(dev):
        d_indptr = torch.from_numpy(indptr).to(dev)
        d_indices = torch.from_numpy(indices).to(dev) if nnz else torch.zeros(1, dtype=torch.int32, device=dev)
        d_data = torch.from_numpy(data).to(dev) if nnz else torch.zeros(1, dtype=torch.float64, device=dev)
        o_indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        o_indices = torch.empty(cap, dtype=torch.int32, device=dev)
        o_re = torch.empty(cap, dtype=torch.float64, device=dev)
        o_im = torch.empty(cap, dtype=torch.float64, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        q_angle = (1j * 2 * np.pi * q).imag                 # the reference's own expression (utils.py:124)
        alpha = -1.0 if ppr_alpha is None else float(ppr_alpha)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = lib.srg_mag_norm_csr(_ptr(d_indptr), _ptr(d_indices), _ptr(d_data), vt, n, nnz, float(r), float(q_angle),
                                  alpha, _ptr(o_indptr), _ptr(o_indices), None, _ptr(o_re), _ptr(o_im), None, None,
                                  _ptr(flags), stream)
        _lib.check(rc)
        fl = int(flags.item())
        if fl & _lib.SRG_FLAG_BAD_INDEX:
            raise _lib.SrgError(_lib.SRG_ERR_INVALID, "column index out of range")
        if fl & _lib.SRG_FLAG_ZERO_PRODUCT:
            raise _lib.SrgUnsupported(_lib.SRG_ERR_UNSUPPORTED,
                                      "a blended real entry is exactly zero (scipy would drop it from the pattern)")
        h_indptr = o_indptr.cpu().numpy()
        m = int(h_indptr[-1]) if n else 0
        h_indices = o_indices[:m].cpu().numpy()
        real = sp.csr_matrix((o_re[:m].cpu().numpy(), h_indices, h_indptr), shape=(n, n), copy=False)
        imag = sp.csr_matrix((o_im[:m].cpu().numpy(), h_indices.copy(), h_indptr.copy()), shape=(n, n), copy=False)
    real.has_sorted_indices = imag.has_sorted_indices = True
    return real, imag


def adj_to_un_in_out_dir_symmetric_norm(adj, r, device=0):
    """Undirected / in / out normalised operators of a directed graph on the GPU
    (SSRG/operators/utils.py:195-260, the normaliser of TwoDirLaplacianGraphOp).

    The reference densifies: P = D^-1 (A + I) as an N x N float32 matrix, in_L = P^T P and out_L = P P^T by
    dense products, then torch.nonzero.  Here P stays a CSR and the two products are sparse x sparse
    (``srg_spgemm_csr_f32``); the float32 degree normalisation (row sums, pow, D^(r-1) L D^(-r)) follows the
    reference's operation order.  Edge weights are ignored exactly as there (``torch.ones``, :197).
    Returns three ``scipy.sparse.csr_matrix`` with float32 data: (un, in, out)."""
    from .. import device as sdev
    from ..sparse_mm import csr_sym_scale, csr_to_scipy, csr_transpose, spgemm
    lib = _lib.load()
    if lib.srg_device_count() <= 0:
        raise _lib.SrgError(_lib.SRG_ERR_NODEV, "no CUDA device visible: libsrgnn_b200 has no CPU fallback")
    if not sp.issparse(adj):
        raise TypeError("The adjacency matrix must be a scipy sparse matrix!")
    csr = adj.tocsr() if not isinstance(adj, sp.csr_matrix) else adj
    indptr, indices, _, _, n, nnz = csr_host_parts(csr)
    dev = torch.device("cuda", int(device))
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        d_indptr = torch.from_numpy(indptr).to(dev)
        d_indices = torch.from_numpy(indices).to(dev) if nnz else torch.zeros(1, dtype=torch.int32, device=dev)
        # A with one loop appended per node, then rows sorted and duplicates summed: the counts that
        # torch.sparse(...).to_dense() (:215) and scatter_add (:203) see
        cap = nnz + n
        l_indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        l_indices = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
        _lib.check(lib.srg_csr_append_diagonal(_ptr(d_indptr), _ptr(d_indices), n, _ptr(l_indptr), _ptr(l_indices), stream))
        c_indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        c_indices = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
        c_vals = torch.empty(max(cap, 1), dtype=torch.float64, device=dev)
        c_nnz = torch.zeros(1, dtype=torch.int32, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(lib.srg_csr_canonicalize(_ptr(l_indptr), _ptr(l_indices), None, _lib.SRG_VAL_ONES, n, cap,
                                            _ptr(c_indptr), _ptr(c_indices), _ptr(c_vals), _ptr(c_nnz), _ptr(flags), stream))
        if int(flags.item()) & _lib.SRG_FLAG_BAD_INDEX:
            raise _lib.SrgError(_lib.SRG_ERR_INVALID, "column index out of range")
        m = int(c_nnz.item())
        a_tilde = sdev.DeviceCSR(c_indptr, c_indices[:max(m, 1)], c_vals[:max(m, 1)].to(torch.float32), n, m)
        un = csr_sym_scale(a_tilde, float(r))                       # :204-210
        p = csr_sym_scale(a_tilde, 0.0)                             # deg^-1 * count  (:212-214)
        pt = csr_transpose(p)
        in_l = spgemm(pt, p, drop_zeros=True)                       # P^T P  (:216, :220-227)
        out_l = spgemm(p, pt, drop_zeros=True)                      # P P^T  (:217, :240-247)
        in_n = csr_sym_scale(in_l, float(r))                        # :229-237
        out_n = csr_sym_scale(out_l, float(r))                      # :249-257
        return csr_to_scipy(un), csr_to_scipy(in_n), csr_to_scipy(out_n)


def _pattern_plus_loops(lib, csr, dev, stream):
    """DeviceCSR of the adjacency pattern with one loop appended per node, rows sorted, duplicates summed into
    float32 counts (add_self_loops + the duplicate sum of the sparse constructor, utils.py:264-271)."""
    from .. import device as sdev
    indptr, indices, _, _, n, nnz = csr_host_parts(csr)
    d_indptr = torch.from_numpy(indptr).to(dev)
    d_indices = torch.from_numpy(indices).to(dev) if nnz else torch.zeros(1, dtype=torch.int32, device=dev)
    cap = nnz + n
    l_indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    l_indices = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    _lib.check(lib.srg_csr_append_diagonal(_ptr(d_indptr), _ptr(d_indices), n, _ptr(l_indptr), _ptr(l_indices), stream))
    c_indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    c_indices = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    c_vals = torch.empty(max(cap, 1), dtype=torch.float64, device=dev)
    c_nnz = torch.zeros(1, dtype=torch.int32, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.srg_csr_canonicalize(_ptr(l_indptr), _ptr(l_indices), None, _lib.SRG_VAL_ONES, n, cap,
                                        _ptr(c_indptr), _ptr(c_indices), _ptr(c_vals), _ptr(c_nnz), _ptr(flags), stream))
    if int(flags.item()) & _lib.SRG_FLAG_BAD_INDEX:
        raise _lib.SrgError(_lib.SRG_ERR_INVALID, "column index out of range")
    m = int(c_nnz.item())
    return sdev.DeviceCSR(c_indptr, c_indices[:max(m, 1)], c_vals[:max(m, 1)].to(torch.float32), n, m)


def adj_to_fast_ppr_approx_symmetric_norm(adj, r, ppr_alpha, max_iter=100, device=0, tol=1e-6):
    """Fast PPR-approximation normaliser of a directed graph on the GPU (SSRG/operators/utils.py:262-335, the
    normaliser of SymDirFastPprApproxGraphOp): stationary distribution of the teleporting walk by fixed-point sweeps
    (fp64, same stopping rule: |x - x_old|_2 <= 1e-6 or ``max_iter`` sweeps), the symmetrised Laplacian
    ``(Pi^1/2 P Pi^-1/2 + Pi^-1/2 P^T Pi^1/2) / 2`` and the float32 degree normalisation.  Returns a
    ``scipy.sparse.csr_matrix`` with float32 data.  Compared with the reference's own outputs on hardware
    (tests/test_fast_ppr.py, tests/golden/reference_ext.npz)."""
    from ..sparse_mm import csr_sym_scale, csr_to_scipy, csr_transpose
    from ..device import DeviceCSR
    lib = _lib.load()
    if lib.srg_device_count() <= 0:
        raise _lib.SrgError(_lib.SRG_ERR_NODEV, "no CUDA device visible: libsrgnn_b200 has no CPU fallback")
    if not sp.issparse(adj):
        raise TypeError("The adjacency matrix must be a scipy sparse matrix!")
    csr = adj.tocsr() if not isinstance(adj, sp.csr_matrix) else adj
    if csr.shape[0] == 0:                                             # empty graph: nothing to normalise (and no 1 / n)
        return sp.csr_matrix((0, 0), dtype=np.float32)
    dev = torch.device("cuda", int(device))
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        a1 = _pattern_plus_loops(lib, csr, dev, stream)
        n = a1.n
        _, deg32 = csr_sym_scale(a1, 0.0, want_degree=True)           # row sums of the counts (exact small integers)
        deg = deg32.to(torch.float64)
        a1t = csr_transpose(a1)
        x = torch.full((n,), 1.0 / (1.0 + ppr_alpha) / n, dtype=torch.float64, device=dev)
        y = torch.empty_like(x)
        stats = torch.zeros(3, dtype=torch.float64, device=dev)
        err = float(np.sqrt(n) * (1.0 / (1.0 + ppr_alpha) / n))         # |s - 0|_2 of the first loop test
        sweeps = 0
        while err > tol:
            _lib.check(lib.srg_ppr_iterate_f64(_ptr(a1t.indptr), _ptr(a1t.indices), _ptr(a1t.data), _ptr(deg), n,
                                               float(ppr_alpha), _ptr(x), _ptr(y), _ptr(stats), stream))
            x, y = y, x
            err = float(np.sqrt(stats[1].item()))
            sweeps += 1
            if sweeps >= max_iter:
                break
        if sweeps == 0:                                                # tol above the start norm: pi = uniform
            stats[2] = x.sum()
        cap = max(2 * a1.nnz, 1)
        o_indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        o_indices = torch.empty(cap, dtype=torch.int32, device=dev)
        o_vals = torch.empty(cap, dtype=torch.float32, device=dev)
        _lib.check(lib.srg_ppr_symmetrize(_ptr(a1.indptr), _ptr(a1.indices), _ptr(a1.data), _ptr(deg), _ptr(x), _ptr(stats),
                                          n, a1.nnz, _ptr(o_indptr), _ptr(o_indices), _ptr(o_vals), stream))
        m = int(o_indptr[-1].item())
        lap = DeviceCSR(o_indptr, o_indices[:max(m, 1)], o_vals[:max(m, 1)], n, m)
        return csr_to_scipy(csr_sym_scale(lap, float(r)))


def adj_to_slow_first_second_ppr_approx_symmetric_norm(adj, r, ppr_alpha, device=0, max_sweeps=2000, tol=1e-13):
    """First- and second-order PPR-approximation operators of a directed graph on the GPU
    (SSRG/operators/utils.py:337-424, the normaliser of SymDirTwoOrderPprApproxGraphOp).

    The reference works on dense N x N matrices and takes the stationary vector from a dense LAPACK eigen-
    decomposition of the (N+1) x (N+1) teleport matrix in float32.  Here P = D^-1 (A + I) stays a CSR, the stationary
    vector is the fixed point of the same matrix by power iteration in fp64 (it agrees with LAPACK's float32 result
    to float32 accuracy), the first-order Laplacian is the pi-weighted symmetrisation, the second-order one is
    ``(P^T P + P P^T) / 2`` on the entries where both products are non-zero (sparse x sparse products), both
    degree-normalised in float32.  Returns two ``scipy.sparse.csr_matrix`` (float32).  Compared with the reference's own
    outputs on hardware (tests/test_fast_ppr.py)."""
    from ..device import DeviceCSR
    from ..sparse_mm import csr_sym_scale, csr_to_scipy, csr_transpose, spgemm
    lib = _lib.load()
    if lib.srg_device_count() <= 0:
        raise _lib.SrgError(_lib.SRG_ERR_NODEV, "no CUDA device visible: libsrgnn_b200 has no CPU fallback")
    if not sp.issparse(adj):
        raise TypeError("The adjacency matrix must be a scipy sparse matrix!")
    csr = adj.tocsr() if not isinstance(adj, sp.csr_matrix) else adj
    if csr.shape[0] == 0:
        return sp.csr_matrix((0, 0), dtype=np.float32), sp.csr_matrix((0, 0), dtype=np.float32)
    dev = torch.device("cuda", int(device))
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        a1 = _pattern_plus_loops(lib, csr, dev, stream)
        n = a1.n
        p, deg32 = csr_sym_scale(a1, 0.0, want_degree=True)           # P = D^-1 A1 (float32, :345-352)
        pt = csr_transpose(p)
        # stationary vector of the teleport chain (:353-369)
        x = torch.full((n + 1,), 1.0 / (n + 1), dtype=torch.float64, device=dev)
        y = torch.empty_like(x)
        stats = torch.zeros(3, dtype=torch.float64, device=dev)
        prev = None
        for sweep in range(max_sweeps):
            _lib.check(lib.srg_teleport_iterate_f64(_ptr(pt.indptr), _ptr(pt.indices), _ptr(pt.data), n, float(ppr_alpha),
                                                    _ptr(x), _ptr(y), _ptr(stats), stream))
            x, y = y, x
            # the float32 rows of P do not sum to exactly 1, so the dominant eigenvalue is 1 + O(1e-8): keep the
            # iterate on the simplex and test the DIRECTION for convergence (every 8 sweeps, one sync)
            x.div_(x.sum())
            if sweep % 8 == 7:
                if prev is not None and float((x - prev).abs().sum().item()) <= tol:
                    break
                prev = x.clone()
        stats[2] = x[:n].sum()                      # pi = x[:n] / sum(x[:n]) inside the symmetrisation
        deg = deg32.to(torch.float64)
        cap = max(2 * a1.nnz, 1)
        o_indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        o_indices = torch.empty(cap, dtype=torch.int32, device=dev)
        o_vals = torch.empty(cap, dtype=torch.float32, device=dev)
        _lib.check(lib.srg_ppr_symmetrize(_ptr(a1.indptr), _ptr(a1.indices), _ptr(a1.data), _ptr(deg), _ptr(x), _ptr(stats),
                                          n, a1.nnz, _ptr(o_indptr), _ptr(o_indices), _ptr(o_vals), stream))
        m1 = int(o_indptr[-1].item())
        one = csr_sym_scale(DeviceCSR(o_indptr, o_indices[:max(m1, 1)], o_vals[:max(m1, 1)], n, m1), float(r))   # :374-400
        # second order (:402-421)
        l_in = spgemm(pt, p, drop_zeros=True)
        l_out = spgemm(p, pt, drop_zeros=True)
        t_indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        t_indices = torch.empty(max(l_in.nnz, 1), dtype=torch.int32, device=dev)
        t_vals = torch.empty(max(l_in.nnz, 1), dtype=torch.float32, device=dev)
        _lib.check(lib.srg_csr_intersect_mean_f32(_ptr(l_in.indptr), _ptr(l_in.indices), _ptr(l_in.data), _ptr(l_out.indptr),
                                                  _ptr(l_out.indices), _ptr(l_out.data), n, l_in.nnz, _ptr(t_indptr),
                                                  _ptr(t_indices), _ptr(t_vals), stream))
        m2 = int(t_indptr[-1].item())
        two = csr_sym_scale(DeviceCSR(t_indptr, t_indices[:max(m2, 1)], t_vals[:max(m2, 1)], n, m2), float(r))
        return csr_to_scipy(one), csr_to_scipy(two)


class DeviceHopRunner:
    """Hops with several (already normalised) adjacency matrices over device-resident features: the
    per-hop body of ComGraphOp / TwoDirGraphOp / TwoOrderPprApproxGraphOp.propagate
    (SSRG/operators/base_operator.py:62-306), each hop being ``csr_sparse_dense_matmul`` (utils.py:17-47:
    weights rounded to float32, fp32 FMA chain in stored order) without the host round trip."""

    def __init__(self, adjs, feature, device=0):
        from .. import device as sdev
        self.lib = _lib.load()
        if self.lib.srg_device_count() <= 0:
            raise _lib.SrgError(_lib.SRG_ERR_NODEV, "no CUDA device visible: libsrgnn_b200 has no CPU fallback")
        self.dev = torch.device("cuda", int(device))
        self.sdev = sdev
        self.n, self.f = feature.shape
        self.adjs = []
        with torch.cuda.device(self.dev):
            for a in adjs:
                if not isinstance(a, sp.csr_matrix):
                    a = sp.csr_matrix(a)
                nnz = int(a.indptr[-1])
                self.adjs.append(sdev.DeviceCSR(
                    torch.from_numpy(np.ascontiguousarray(a.indptr, dtype=np.int32)).to(self.dev),
                    torch.from_numpy(np.ascontiguousarray(a.indices[:nnz], dtype=np.int32)).to(self.dev)
                    if nnz else torch.zeros(1, dtype=torch.int32, device=self.dev),
                    torch.from_numpy(np.ascontiguousarray(a.data[:nnz]).astype(np.float32)).to(self.dev)
                    if nnz else torch.zeros(1, dtype=torch.float32, device=self.dev),
                    a.shape[0], nnz))
            x = torch.from_numpy(np.ascontiguousarray(feature, dtype=np.float32)).to(self.dev)
            self.x0 = sdev.pack_features(x) if self.n * self.f else x
        self.ld = self.x0.shape[1] if self.n * self.f else self.f

    def hop(self, which, x):
        """A[which] @ x on the device (x and the result in the padded layout)."""
        if self.n * self.f == 0:
            return x.clone()
        with torch.cuda.device(self.dev):
            return self.sdev.spmm(self.adjs[which], x, self.f)

    def _update(self, acc, x, mode, w, first):
        stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        _lib.check(self.lib.srg_aggregate_update_f32(_ptr(acc), self.ld, 0, _ptr(x), self.ld, self.n, self.f, mode,
                                                     C.c_float(w), first, stream))

    def add_(self, acc, x):
        """acc += x (elementwise fp32 add, numpy's in-place add)."""
        if self.n * self.f:
            with torch.cuda.device(self.dev):
                self._update(acc, x, _lib.SRG_AGG_SUM, 1.0, 0)
        return acc

    def neg(self, x):
        """-x as a fresh matrix."""
        out = torch.empty_like(x)
        if self.n * self.f:
            with torch.cuda.device(self.dev):
                self._update(out, x, _lib.SRG_AGG_WEIGHTED, -1.0, 1)
        return out

    def to_host(self, x):
        """padded device matrix -> torch.FloatTensor n x F on the CPU."""
        if self.n * self.f == 0:
            return torch.zeros((self.n, self.f), dtype=torch.float32)
        with torch.cuda.device(self.dev):
            return self.sdev.unpack_features(x, self.f).cpu()
