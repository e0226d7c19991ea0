"""Host-buffer helpers of the propagation path, mirroring SSRG/operators/utils.py.

Reference functions mirrored (same names, argument meaning and return layout):
  * ``csr_sparse_dense_matmul(adj, feature)``  — SSRG/operators/utils.py:17-47 (ctypes call into
    libmatmul.so:FloatCSRMulDenseOMP).  Here the same symbol of libsrgnn_b200.so runs the hop on
    the GPU.
  * ``adj_to_symmetric_norm(adj, r)``          — SSRG/operators/utils.py:81-93 (scipy fp64).
    Here: CSR kernels on the GPU; the result comes back as a scipy matrix.
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
           "csr_host_parts"]


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
    answer = np.zeros(feature.shape, dtype=np.float32).reshape(-1)
    data = np.ascontiguousarray(adj.data, dtype=np.float32)
    indices = np.ascontiguousarray(adj.indices, dtype=np.int32)
    indptr = np.ascontiguousarray(adj.indptr, dtype=np.int32)
    mat = np.ascontiguousarray(feature).reshape(-1)
    mat_row, mat_col = feature.shape
    if lib.srg_device_count() <= 0:
        raise _lib.SrgError(_lib.SRG_ERR_NODEV, "no CUDA device visible: libsrgnn_b200 has no CPU fallback")
    lib.FloatCSRMulDenseOMP(_ptr(answer), _ptr(data), _ptr(indices), _ptr(indptr), _ptr(mat), mat_row, mat_col)
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
