"""Device-resident entry points: the same stages as ``operators/`` on torch CUDA tensors.

torch is plumbing only (device memory + the current stream); every stage is a kernel of
libsrgnn_b200.so called through the C ABI with raw device pointers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib

__all__ = ["DeviceCSR", "upload_csr", "pad_ld", "pack_features", "unpack_features", "sym_norm",
           "spmm", "propagate"]


def _stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def pad_ld(f: int) -> int:
    """Leading dimension of the device feature layout: rows padded to whole 32-byte sectors."""
    return (int(f) + 7) // 8 * 8


@dataclass
class DeviceCSR:
    indptr: torch.Tensor            # int32 [n+1]
    indices: torch.Tensor           # int32 [>= nnz]
    data: torch.Tensor | None       # float32 / float64 [>= nnz] or None (all ones)
    n: int
    nnz: int                        # stored entries, or an upper bound (see nnz_bound)

    @property
    def nnz_bound(self) -> int:
        """Upper bound of indptr[n] that is known on the host (the capacity of the arrays)."""
        return int(self.nnz) if self.nnz >= 0 else int(self.indices.numel())

    @property
    def val_dtype(self) -> int:
        if self.data is None:
            return _lib.SRG_VAL_ONES
        return _lib.SRG_VAL_F64 if self.data.dtype == torch.float64 else _lib.SRG_VAL_F32


def upload_csr(adj, device="cuda", ones_as_null=False) -> DeviceCSR:
    """scipy CSR -> DeviceCSR (int32 structure; data keeps float32/float64)."""
    import numpy as np
    n = adj.shape[0]
    nnz = int(adj.indptr[-1])
    indptr = torch.from_numpy(np.ascontiguousarray(adj.indptr, dtype=np.int32)).to(device)
    indices = torch.from_numpy(np.ascontiguousarray(adj.indices[:nnz], dtype=np.int32)).to(device)
    data = None
    if not ones_as_null:
        d = adj.data[:nnz]
        if d.dtype not in (np.float32, np.float64):
            d = d.astype(np.float64)
        data = torch.from_numpy(np.ascontiguousarray(d)).to(device)
    return DeviceCSR(indptr, indices, data, n, nnz)


def pack_features(x: torch.Tensor, ld: int | None = None, mask: torch.Tensor | None = None) -> torch.Tensor:
    """n x F (contiguous, cuda) -> n x ld padded layout (pad columns zero); optional x * mask."""
    lib = _lib.load()
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    n, f = x.shape
    ld = pad_ld(f) if ld is None else ld
    out = torch.empty((n, ld), dtype=torch.float32, device=x.device)
    if mask is not None:
        assert mask.is_cuda and mask.dtype == torch.int32 and mask.is_contiguous() and mask.shape == x.shape
    _lib.check(lib.srg_pack_features_f32(_p(x), f, _p(out), ld, n, f, _p(mask), _stream_ptr(x.device)))
    return out


def unpack_features(xp: torch.Tensor, f: int) -> torch.Tensor:
    lib = _lib.load()
    n, ld = xp.shape
    out = torch.empty((n, f), dtype=torch.float32, device=xp.device)
    _lib.check(lib.srg_unpack_features_f32(_p(xp), ld, _p(out), f, n, f, _stream_ptr(xp.device)))
    return out


def sym_norm(a: DeviceCSR, r: float, ppr_alpha: float | None = None, want_f64=False, want_degree=False):
    """Device normalisation.  Returns (DeviceCSR with float32 data, flags tensor, extras dict).

    No host synchronisation: the caller checks ``flags`` (int32[1]) when convenient.
    """
    lib = _lib.load()
    dev = a.indptr.device
    n, nnz = a.n, a.nnz
    s = _stream_ptr(dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    o_indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    o_count = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    _lib.check(lib.srg_degree_selfloop_csr(_p(a.indptr), _p(a.indices), _p(a.data), a.val_dtype, n, nnz, _p(o_indptr),
                                           _p(o_count), _p(flags), s))
    cap = max(nnz + n, 1)
    o_indices = torch.empty(cap, dtype=torch.int32, device=dev)
    o_val32 = torch.empty(cap, dtype=torch.float32, device=dev)
    o_val64 = torch.empty(cap, dtype=torch.float64, device=dev) if want_f64 else None
    o_deg = torch.empty(max(n, 1), dtype=torch.float64, device=dev) if want_degree else None
    alpha = -1.0 if ppr_alpha is None else float(ppr_alpha)
    _lib.check(lib.srg_sym_norm_csr(_p(a.indptr), _p(a.indices), _p(a.data), a.val_dtype, n, nnz, _p(o_indptr),
                                    float(r), alpha, _p(o_indices), _p(o_deg), _p(o_val64), _p(o_val32), _p(flags), s))
    out = DeviceCSR(o_indptr, o_indices, o_val32, n, nnz + n)   # nnz: upper bound (exact count stays on the device)
    return out, flags, {"val64": o_val64, "degree": o_deg, "count": o_count}


def spmm(a: DeviceCSR, x: torch.Tensor, f: int | None = None, out: torch.Tensor | None = None,
         n_rows: int | None = None) -> torch.Tensor:
    """One hop Y = A X on padded (or plain) row-major device matrices; ``f`` = logical width."""
    lib = _lib.load()
    assert x.is_cuda and x.dtype == torch.float32 and x.stride(1) == 1
    f = x.shape[1] if f is None else f
    rows = a.n if n_rows is None else n_rows
    if out is None:
        out = torch.empty((rows, x.shape[1]), dtype=torch.float32, device=x.device)
        w4 = (f + 3) // 4 * 4
        if out.shape[1] > w4:
            out[:, w4:].zero_()      # the kernels write whole float4 columns up to F: keep the remaining pad columns zero
    _lib.check(lib.srg_spmm_csr_f32(_p(a.indptr), _p(a.indices), _p(a.data), rows, a.nnz_bound, _p(x), x.stride(0),
                                    _p(out), out.stride(0), f, _stream_ptr(x.device)))
    return out


def propagate(a_norm: DeviceCSR, x0: torch.Tensor, f: int, k: int, hops: list | None = None) -> list:
    """[x0, A x0, ..., A^k x0] on the device (x0 in the padded layout, all buffers n x ld)."""
    lib = _lib.load()
    n, ld = x0.shape
    if hops is None:
        hops = [x0] + [torch.empty_like(x0) for _ in range(k)]
        w4 = (f + 3) // 4 * 4
        if ld > w4:
            for h in hops[1:]:
                h[:, w4:].zero_()    # pad columns beyond the last float4 of F are never written by the hop kernels
    ptrs = (C.c_void_p * (k + 1))(*[h.data_ptr() for h in hops])
    _lib.check(lib.srg_propagate_khop_f32(_p(a_norm.indptr), _p(a_norm.indices), _p(a_norm.data), n, a_norm.nnz_bound,
                                          ptrs, ld, f, k, _stream_ptr(x0.device)))
    return hops
