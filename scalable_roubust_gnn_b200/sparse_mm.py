"""Per-epoch sparse product of the GCN model on the device: ``torch.mm(adj, x)`` with autograd.

Reference: ``Layer2GraphConvolution.forward`` (SSRG/models/base_scalable/simple_models.py:225-240)
multiplies by ``self.adj`` twice per step; ``self.adj`` is the normalised adjacency converted by
``scipy_sparse_mat_to_torch_sparse_tensor`` (SSRG/models/utils.py:5-15: ``.tocoo().astype(np.float32)``
-> ``torch.sparse.FloatTensor``), set in ``BaseSGModel.preprocess`` (base_model.py:46-53).

``scipy_sparse_mat_to_device_adj`` is the drop-in for that conversion: it returns a ``DeviceAdj`` that
the *unmodified* reference layer can use — ``torch.mm(adj, x)``, ``torch.spmm``, ``torch.sparse.mm``,
``torch.matmul`` and ``adj @ x`` are intercepted through ``__torch_function__`` and run the
propagation hop kernel of libsrgnn_b200.so (forward: ``Y = A X``; backward: ``dX = A^T dY`` with the
transpose built once on the device, ``srg_csr_transpose_f32``).  Each output row is one sequential
fp32 FMA chain in CSR order (no atomics): results are deterministic run to run.
torch is plumbing (autograd graph, device memory, stream); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib
from .device import DeviceCSR, _p, _stream_ptr, spmm

__all__ = ["DeviceAdj", "scipy_sparse_mat_to_device_adj", "csr_transpose", "spgemm", "csr_sym_scale", "csr_to_scipy"]


def csr_transpose(a: DeviceCSR) -> DeviceCSR:
    """Transpose of a float32 DeviceCSR on the device (rows stay sorted)."""
    lib = _lib.load()
    dev = a.indptr.device
    cap = a.nnz_bound
    vals = a.data
    if vals is not None and vals.dtype != torch.float32:
        vals = vals.to(torch.float32)
    o_indptr = torch.empty(a.n + 1, dtype=torch.int32, device=dev)
    o_indices = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    o_vals = torch.empty(max(cap, 1), dtype=torch.float32, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.srg_csr_transpose_f32(_p(a.indptr), _p(a.indices), _p(vals), a.n, cap, _p(o_indptr), _p(o_indices),
                                         _p(o_vals), _p(flags), _stream_ptr(dev)))
    if int(flags.item()) & _lib.SRG_FLAG_BAD_INDEX:
        raise _lib.SrgError(_lib.SRG_ERR_INVALID, "csr_transpose: column index out of range")
    return DeviceCSR(o_indptr, o_indices, o_vals, a.n, a.nnz)


class _SpmmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, adj):
        ctx.adj = adj
        return adj._hop(adj.csr, x)

    @staticmethod
    def backward(ctx, grad_out):
        return ctx.adj._hop(ctx.adj.transposed().csr, grad_out), None


class DeviceAdj:
    """A square float32 sparse matrix resident on the GPU that behaves as the left operand of
    ``torch.mm`` (see the module docstring)."""

    _MM_FUNCS = None

    def __init__(self, csr: DeviceCSR, transpose: "DeviceAdj | None" = None):
        if csr.data is None:
            ones = torch.ones(max(csr.nnz_bound, 1), dtype=torch.float32, device=csr.indptr.device)
            csr = DeviceCSR(csr.indptr, csr.indices, ones, csr.n, csr.nnz)
        elif csr.data.dtype != torch.float32:
            csr = DeviceCSR(csr.indptr, csr.indices, csr.data.to(torch.float32), csr.n, csr.nnz)
        self.csr = csr
        self._t = transpose

    # -- tensor-like surface the reference touches -------------------------------------------------
    @property
    def shape(self):
        return torch.Size((self.csr.n, self.csr.n))

    @property
    def device(self):
        return self.csr.indptr.device

    @property
    def is_sparse(self):
        return True

    def to(self, *args, **kwargs):      # `adj.to(device)` in training loops: already resident
        return self

    def transposed(self) -> "DeviceAdj":
        if self._t is None:
            self._t = DeviceAdj(csr_transpose(self.csr), transpose=self)
        return self._t

    def t(self):
        return self.transposed()

    # -- the product -----------------------------------------------------------------------------------
    @staticmethod
    def _hop(csr, x):
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 2):
            raise TypeError("DeviceAdj: the dense operand must be a 2-d CUDA tensor")
        if x.shape[0] != csr.n:
            raise RuntimeError(f"DeviceAdj: size mismatch, adj is {csr.n} x {csr.n}, dense is {tuple(x.shape)}")
        x = x.detach().to(torch.float32).contiguous()
        if x.shape[1] == 0 or csr.n == 0:
            return torch.zeros_like(x)
        return spmm(csr, x)

    def mm(self, x):
        return _SpmmFn.apply(x, self)

    def __matmul__(self, x):
        return self.mm(x)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        if cls._MM_FUNCS is None:
            cls._MM_FUNCS = {torch.mm, torch.spmm, torch.sparse.mm, torch.matmul, torch.Tensor.matmul}
        kwargs = kwargs or {}
        if func in cls._MM_FUNCS and len(args) == 2 and isinstance(args[0], DeviceAdj) and not kwargs:
            return args[0].mm(args[1])
        return NotImplemented


def scipy_sparse_mat_to_device_adj(sparse_mx, device="cuda") -> DeviceAdj:
    """Drop-in for ``scipy_sparse_mat_to_torch_sparse_tensor`` (SSRG/models/utils.py:5-15): the values
    are rounded to float32 exactly as ``.astype(np.float32)`` does there; duplicates are summed and rows
    sorted (what ``torch.sparse`` coalescing would do)."""
    if not sp.issparse(sparse_mx):
        raise TypeError("The adjacency matrix must be a scipy sparse matrix!")
    if sparse_mx.shape[0] != sparse_mx.shape[1]:
        raise ValueError("The adjacency matrix must be square!")
    m = sparse_mx.tocsr().astype(np.float32)
    m.sum_duplicates()
    m.sort_indices()
    nnz = int(m.indptr[-1])
    if nnz > np.iinfo(np.int32).max:
        raise _lib.SrgError(_lib.SRG_ERR_RANGE, "nnz exceeds the int32 CSR range")
    csr = DeviceCSR(torch.from_numpy(np.ascontiguousarray(m.indptr, dtype=np.int32)).to(device),
                    torch.from_numpy(np.ascontiguousarray(m.indices[:nnz], dtype=np.int32)).to(device),
                    torch.from_numpy(np.ascontiguousarray(m.data[:nnz], dtype=np.float32)).to(device),
                    m.shape[0], nnz)
    return DeviceAdj(csr)


def spgemm(a: DeviceCSR, b: DeviceCSR, drop_zeros: bool = False, cap: int | None = None) -> DeviceCSR:
    """C = A B for square float32 DeviceCSR operands (``srg_spgemm_csr_f32``: expand / sort / compress, sequential
    fp32 sums in ascending k).  ``data=None`` operands are all ones.  Rows of C are sorted; ``nnz`` is exact."""
    lib = _lib.load()
    dev = a.indptr.device
    if a.n != b.n:
        raise ValueError("spgemm: inner dimensions differ")
    for m in (a, b):
        if m.data is not None and m.data.dtype != torch.float32:
            raise TypeError("spgemm: float32 values expected")
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    stream = _stream_ptr(dev)
    o_indptr = torch.empty(a.n + 1, dtype=torch.int32, device=dev)
    nnz = C.c_int64(0)
    if cap is None:
        # count pass: capacity 0 reports the size in the error text; cheaper: bound by the product count
        cap = int(_product_count(a, b))
    o_indices = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    o_vals = torch.empty(max(cap, 1), dtype=torch.float32, device=dev)
    _lib.check(lib.srg_spgemm_csr_f32(_p(a.indptr), _p(a.indices), _p(a.data), a.n, a.n, _p(b.indptr), _p(b.indices),
                                      _p(b.data), b.n, 1 if drop_zeros else 0, _p(o_indptr), _p(o_indices), _p(o_vals),
                                      cap, C.byref(nnz), _p(flags), stream))
    if int(flags.item()) & _lib.SRG_FLAG_BAD_INDEX:
        raise _lib.SrgError(_lib.SRG_ERR_INVALID, "spgemm: column index out of range")
    m = int(nnz.value)
    return DeviceCSR(o_indptr, o_indices[:max(m, 1)], o_vals[:max(m, 1)], a.n, m)


def _product_count(a: DeviceCSR, b: DeviceCSR) -> int:
    """Upper bound of nnz(A B): the number of intermediate products (index arithmetic on the row lengths)."""
    m = int(a.indptr[-1].item())
    if m == 0:
        return 0
    blen = (b.indptr[1:] - b.indptr[:-1]).to(torch.int64)
    return int(blen[a.indices[:m].to(torch.int64)].sum().item())


def csr_sym_scale(a: DeviceCSR, r: float, want_degree: bool = False):
    """(deg^(r-1) * v) * deg_col^(-r) in float32 with deg = row sums of ``a`` (``srg_csr_sym_scale_f32``)."""
    lib = _lib.load()
    dev = a.indptr.device
    out = torch.empty(max(int(a.indices.numel()), 1), dtype=torch.float32, device=dev)
    deg = torch.empty(max(a.n, 1), dtype=torch.float32, device=dev) if want_degree else None
    _lib.check(lib.srg_csr_sym_scale_f32(_p(a.indptr), _p(a.indices), _p(a.data), a.n, C.c_float(r), _p(out), _p(deg),
                                         _stream_ptr(dev)))
    res = DeviceCSR(a.indptr, a.indices, out, a.n, a.nnz)
    return (res, deg) if want_degree else res


def csr_to_scipy(a: DeviceCSR, dtype=np.float32):
    indptr = a.indptr.cpu().numpy()
    m = int(indptr[-1]) if a.n else 0
    data = (a.data[:m].cpu().numpy() if a.data is not None else np.ones(m, dtype=dtype)).astype(dtype, copy=False)
    out = sp.csr_matrix((data, a.indices[:m].cpu().numpy(), indptr), shape=(a.n, a.n), copy=False)
    out.has_sorted_indices = True
    return out
