"""Chebyshev heat-kernel wavelets on the GPU.

Mirrors ``WaveletSparsifier`` (wavelet/src/utils.py:70-138) and the batched twin
``SpectralModel.calculate_wavelet`` (SSRG/models/base_scalable/base_model.py:236-265).  In the
reference the arithmetic lives in pygsp (absent, un-pinned: PARITY UNPINNED); the recurrence that
``oracle.cheby_op`` restates runs here as fused SpMM + epilogue kernels in fp64
(csrc/cheby.cu), both scales sharing every T_k.

Host-side setup that stays on the host (scalars, a handful of flops): the m+1 Chebyshev
coefficients of the heat kernel and lambda_max (ARPACK in pygsp; an input of the kernel).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib
from .device import DeviceCSR, _p, _stream_ptr, upload_csr

__all__ = ["laplacian", "heat_cheby_coeffs", "estimate_lmax", "estimate_lmax_device", "cheby_filter", "WaveletSparsifier", "SpectralModel",
           "wavelet_localize"]


def laplacian(w: DeviceCSR):
    """L = diag(W 1) - W on the device.  Returns (DeviceCSR with float64 data, degree, flags)."""
    lib = _lib.load()
    dev = w.indptr.device
    n, nnz = w.n, w.nnz
    cap = max(nnz + n, 1)
    o_indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    o_indices = torch.empty(cap, dtype=torch.int32, device=dev)
    o_vals = torch.empty(cap, dtype=torch.float64, device=dev)
    o_deg = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.srg_laplacian_csr(_p(w.indptr), _p(w.indices), _p(w.data), w.val_dtype, n, _p(o_indptr),
                                     _p(o_indices), _p(o_vals), _p(o_deg), _p(flags), _stream_ptr(dev)))
    return DeviceCSR(o_indptr, o_indices, o_vals, n, -1), o_deg, flags


def heat_cheby_coeffs(tau: float, lmax: float, order: int) -> np.ndarray:
    """pygsp compute_cheby_coeff(Heat(G, tau), m=order): order+1 quadrature points on [0, lmax]."""
    n_q = order + 1
    a = lmax / 2.0
    j = np.arange(n_q)
    theta = np.pi * (j + 0.5) / n_q
    g = np.exp(-tau * (a * np.cos(theta) + a) / lmax)
    return np.array([2.0 / n_q * np.sum(g * np.cos(np.pi * o * (j + 0.5) / n_q)) for o in range(order + 1)])


def estimate_lmax(lap_host) -> float:
    """pygsp Graph.estimate_lmax: 1.01 x ARPACK's largest eigenvalue (tol 5e-3, ncv = min(N, 10))."""
    from scipy.sparse.linalg import eigsh
    n = lap_host.shape[0]
    return float(eigsh(lap_host, k=1, tol=5e-3, ncv=min(n, 10), return_eigenvectors=False)[0]) * 1.01


def estimate_lmax_device(lap: DeviceCSR, tol: float = 5e-3, max_steps: int = 100) -> float:
    """pygsp Graph.estimate_lmax on the device: 1.01 x the largest eigenvalue of the Laplacian by Lanczos
    (csrc/chebysp.cu), to the tolerance of the reference's ARPACK call (5e-3 relative)."""
    lib = _lib.load()
    lam, steps = C.c_double(), C.c_int32()
    _lib.check(lib.srg_lanczos_lambda_max_f64(_p(lap.indptr), _p(lap.indices), _p(lap.data), lap.n, float(tol),
                                              int(max_steps), C.byref(lam), C.byref(steps), _stream_ptr(lap.indptr.device)))
    return float(lam.value) * 1.01


def cheby_filter(lap: DeviceCSR, x: torch.Tensor, lmax: float, coeffs, tol: float | None = None, want_f32=False):
    """r_s = sum_k coeffs[s][k] T_k(L~) x for every coefficient vector (fp64, device resident).

    ``x``: float64 cuda tensor n x B with an even row stride.  Returns a list of float64 tensors
    (and a list of float32 tensors when ``want_f32``).
    """
    lib = _lib.load()
    coeffs = np.ascontiguousarray(np.atleast_2d(np.asarray(coeffs, dtype=np.float64)))
    n_scales, m1 = coeffs.shape
    order = m1 - 1
    assert x.is_cuda and x.dtype == torch.float64 and x.stride(1) == 1 and x.stride(0) % 2 == 0
    n, b = x.shape
    ld = x.stride(0)
    outs = [torch.empty((n, ld), dtype=torch.float64, device=x.device) for _ in range(n_scales)]
    outs32 = [torch.empty((n, ld), dtype=torch.float32, device=x.device) for _ in range(n_scales)] if want_f32 else None
    w0 = torch.empty((n, ld), dtype=torch.float64, device=x.device)
    w1 = torch.empty((n, ld), dtype=torch.float64, device=x.device)
    r_ptrs = (C.c_void_p * n_scales)(*[o.data_ptr() for o in outs])
    q_ptrs = (C.c_void_p * n_scales)(*[o.data_ptr() for o in outs32]) if want_f32 else None
    tol_v = float("nan") if tol is None else float(tol)
    _lib.check(lib.srg_cheby_filter_f64(_p(lap.indptr), _p(lap.indices), _p(lap.data), n, _p(x), ld, b, float(lmax),
                                        coeffs.ctypes.data_as(C.POINTER(C.c_double)), n_scales, order, tol_v, r_ptrs,
                                        q_ptrs, ld, _p(w0), _p(w1), _stream_ptr(x.device)))
    outs = [o[:, :b] for o in outs]
    if want_f32:
        return outs, [o[:, :b] for o in outs32]
    return outs


class WaveletSparsifier:
    """Sparsified heat-kernel wavelets Psi(-s), Psi(+s) of a graph (wavelet/src/utils.py:70-138).

    ``graph`` may be a scipy sparse adjacency or a networkx graph (the reference's argument); the
    adjacency is symmetrised the way ``nx.Graph`` does.  Impulse columns are processed in blocks on
    the device (the reference materialises the N x N identity, utils.py:94).
    """

    def __init__(self, graph, scale, approximation_order, tolerance, lmax=None, block=1024, device="cuda",
                 method="sparse"):
        if not sp.issparse(graph):
            import networkx as nx  # only when the caller already uses it
            graph = nx.adjacency_matrix(graph)
        w = sp.csr_matrix(graph, dtype=np.float64)
        w = w.maximum(w.T).tocsr()
        w.sort_indices()
        self.n = w.shape[0]
        self.device = device
        self.block = int(block)
        # "sparse": the recurrence on stored entries only (csrc/chebysp.cu); "blocks": dense impulse column blocks
        # (csrc/cheby.cu), the shape of the reference's own evaluation - both give the same bits
        self.method = method
        self.stats = {}
        self.scales = [-scale, scale]
        self.approximation_order = approximation_order
        self.tolerance = tolerance
        self.phi_matrices = []
        self._w_dev = upload_csr(w, device=device)
        self.lap, self.degree, flags = laplacian(self._w_dev)
        if lmax is None:
            # device Lanczos; estimate_lmax(host Laplacian) reproduces the reference's ARPACK call for comparison
            lmax = estimate_lmax_device(self.lap)
        self.lmax = float(lmax)

    def chebyshev_coefficients(self):
        return np.stack([heat_cheby_coeffs(s, self.lmax, self.approximation_order) for s in self.scales])

    def calculate_all_wavelets(self, normalize=True):
        """Psi(-s), Psi(+s) as scipy CSR float32 (reference: utils.py:125-138).  Chebyshev filter,
        threshold, sparsification, block merge and L1 normalisation all run on the device; only the
        final CSR arrays are copied back."""
        phis = self.calculate_all_wavelets_device(normalize=normalize)
        out = []
        for indptr, cols, vals in phis:
            m = int(indptr[-1].item())
            out.append(sp.csr_matrix((vals[:m].cpu().numpy(), cols[:m].cpu().numpy(), indptr.cpu().numpy()),
                                     shape=(self.n, self.n)))
        self.phi_matrices = out
        return self.phi_matrices

    def _wavelets_sparse(self, normalize):
        """The whole identity impulse at once, on stored entries only (csrc/chebysp.cu)."""
        lib = _lib.load()
        dev_ = self.device
        n = self.n
        s = _stream_ptr(torch.device(dev_) if isinstance(dev_, str) else dev_)
        coeffs = np.ascontiguousarray(self.chebyshev_coefficients(), dtype=np.float64)
        l_nnz = int(self.lap.indptr[-1].item())
        handle = C.c_void_p()
        tol = float("nan") if self.tolerance is None else float(self.tolerance)
        _lib.check(lib.srg_cheby_sparse_run(_p(self.lap.indptr), _p(self.lap.indices), _p(self.lap.data), n, l_nnz,
                                            float(self.lmax), coeffs.ctypes.data_as(C.POINTER(C.c_double)), len(self.scales),
                                            int(self.approximation_order), tol, C.byref(handle), s))
        out = []
        try:
            for sc in range(len(self.scales)):
                nnz, prod, pat = C.c_int64(), C.c_int64(), C.c_int64()
                _lib.check(lib.srg_cheby_sparse_info(handle, sc, C.byref(nnz), C.byref(prod), C.byref(pat)))
                self.stats = {"products": int(prod.value), "pattern_nnz": int(pat.value)}
                indptr = torch.empty(n + 1, dtype=torch.int32, device=dev_)
                cols = torch.empty(max(int(nnz.value), 1), dtype=torch.int32, device=dev_)
                vals = torch.empty(max(int(nnz.value), 1), dtype=torch.float32, device=dev_)
                _lib.check(lib.srg_cheby_sparse_fetch(handle, sc, _p(indptr), _p(cols), _p(vals), s))
                if normalize:
                    _lib.check(lib.srg_csr_row_normalize_l1_f32(n, _p(indptr), _p(vals), s))
                out.append((indptr, cols, vals))
        finally:
            lib.srg_cheby_sparse_free(handle)
        return out

    def calculate_all_wavelets_device(self, normalize=True):
        """Same, device resident: a list (one per scale) of (indptr int32[n+1], cols int32, vals float32)."""
        if self.method == "sparse" and self.approximation_order >= 1:
            try:
                return self._wavelets_sparse(normalize)
            except _lib.SrgUnsupported:
                pass          # isolated nodes (no stored Laplacian diagonal): the dense-block path handles them
        lib = _lib.load()
        dev_ = self.device
        n = self.n
        s = _stream_ptr(torch.device(dev_) if isinstance(dev_, str) else dev_)
        coeffs = self.chebyshev_coefficients()
        n_scales = len(self.scales)
        blocks = [[] for _ in range(n_scales)]          # per scale: (indptr, cols, vals) of every column block
        totals = [torch.zeros(n, dtype=torch.int32, device=dev_) for _ in range(n_scales)]
        for j0 in range(0, n, self.block):
            b = min(self.block, n - j0)
            ld = (b + 1) // 2 * 2
            x = torch.zeros((n, ld), dtype=torch.float64, device=dev_)
            idx = torch.arange(b, device=dev_)
            x[j0 + idx, idx] = 1.0
            xv = x[:, :b] if ld == b else x.as_strided((n, b), (ld, 1))
            _, r32 = cheby_filter(self.lap, xv, self.lmax, coeffs, tol=self.tolerance, want_f32=True)
            for sc, r in enumerate(r32):
                bptr = torch.empty(n + 1, dtype=torch.int32, device=dev_)
                _lib.check(lib.srg_dense_block_to_csr_f32(_p(r), r.stride(0), n, b, j0, _p(bptr), None, None, 0, None, s))
                cnt = int(bptr[-1].item())
                bcols = torch.empty(max(cnt, 1), dtype=torch.int32, device=dev_)
                bvals = torch.empty(max(cnt, 1), dtype=torch.float32, device=dev_)
                _lib.check(lib.srg_dense_block_to_csr_f32(_p(r), r.stride(0), n, b, j0, _p(bptr), _p(bcols), _p(bvals),
                                                          cnt, _p(totals[sc]), s))
                blocks[sc].append((bptr, bcols, bvals))
        out = []
        for sc in range(n_scales):
            indptr = torch.empty(n + 1, dtype=torch.int32, device=dev_)
            _lib.check(lib.srg_exclusive_scan_i32(_p(totals[sc]), n, _p(indptr), s))
            m = int(indptr[-1].item())
            cols = torch.empty(max(m, 1), dtype=torch.int32, device=dev_)
            vals = torch.empty(max(m, 1), dtype=torch.float32, device=dev_)
            cursor = indptr[:n].clone()
            for bptr, bcols, bvals in blocks[sc]:        # block order = ascending columns: rows come out sorted
                _lib.check(lib.srg_csr_block_scatter_f32(n, _p(bptr), _p(bcols), _p(bvals), _p(cursor), _p(cols),
                                                         _p(vals), s))
            if normalize:
                _lib.check(lib.srg_csr_row_normalize_l1_f32(n, _p(indptr), _p(vals), s))
            out.append((indptr, cols, vals))
        return out

    def normalize_matrices(self):
        """L1 row normalisation of ``self.phi_matrices`` (utils.py:106-112) on the device."""
        lib = _lib.load()
        out = []
        for phi in self.phi_matrices:
            phi = sp.csr_matrix(phi, dtype=np.float32)
            indptr = torch.from_numpy(phi.indptr.astype(np.int32)).to(self.device)
            vals = torch.from_numpy(phi.data.copy()).to(self.device)
            _lib.check(lib.srg_csr_row_normalize_l1_f32(self.n, _p(indptr), _p(vals), _stream_ptr(vals.device)))
            out.append(sp.csr_matrix((vals.cpu().numpy(), phi.indices, phi.indptr), shape=phi.shape))
        self.phi_matrices = out


def wavelet_localize(phi, phi_inverse, x, theta=None):
    """``Psi (theta * (Psi^-1 x))`` on the device, differentiable in ``x`` and ``theta``.

    The reference forms the sparse product first — ``spspmm(Psi [diag(theta)], Psi^-1)`` followed by
    ``spmm(product, x)`` (SSRG/models/base_scalable/base_model.py:208-219; wavelet/src/gwnn_layer.py:
    59-85, 111-128) — whose pattern is the 2m-hop neighbourhood.  The same linear map is applied here
    as two hops of the propagation kernel with the diagonal in between: no fill-in, no SpGEMM, and the
    result differs from the reference's association order by fp32 rounding only.
    ``phi`` / ``phi_inverse``: ``sparse_mm.DeviceAdj``; ``x``: cuda float32 n x F; ``theta``: n or n x 1.
    """
    y = phi_inverse.mm(x)
    if theta is not None:
        y = y * theta.reshape(-1, 1)
    return phi.mm(y)


class SpectralModel:
    """Pre-processing of the reference's wavelet model on the device.

    Mirror of ``SpectralModel.__init__`` / ``.preprocess``
    (SSRG/models/base_scalable/base_model.py:171-221): heat-kernel wavelets Psi(-s), Psi(+s) by the
    Chebyshev recurrence in 1000-column impulse blocks (:236-265), threshold, float32 CSR, L1 row
    normalisation (:287-290), then ``processed_feature = [X | relu(Psi(-s) Psi(+s) X)]`` (:208-221).
    Everything between the adjacency upload and the final feature matrix stays on the GPU.
    ``lmax``: pygsp's ARPACK estimate is an input here (``estimate_lmax`` reproduces its call).
    """

    def __init__(self, scale, approximation_order, tolerance, lmax=None, block=1000, device="cuda", method="sparse"):
        self.method = method
        self.scales = [-scale, scale]
        self.approximation_order = approximation_order
        self.tolerance = tolerance
        self.lmax = lmax
        self.block = block
        self.device = device
        self.phi_device = []          # per scale: (indptr, cols, vals) on the device
        self._phi_host = None
        self.processed_feature = None
        self.ncount = None

    @property
    def phi_matrices(self):
        """The two normalised wavelet matrices as scipy CSR float32 (copied back on first use)."""
        if self._phi_host is None:
            self._phi_host = []
            for indptr, cols, vals in self.phi_device:
                m = int(indptr[-1].item())
                self._phi_host.append(sp.csr_matrix((vals[:m].cpu().numpy(), cols[:m].cpu().numpy(),
                                                     indptr.cpu().numpy()), shape=(self.ncount, self.ncount)))
        return self._phi_host

    def density(self):
        """Fractions of stored entries (calculate_density, base_model.py:292-298)."""
        return [int(p[0][-1].item()) / float(self.ncount) ** 2 for p in self.phi_device]

    def preprocess(self, adj, feature):
        from .sparse_mm import DeviceAdj
        if isinstance(feature, torch.Tensor):
            feature = feature.numpy()
        feature = np.ascontiguousarray(feature, dtype=np.float32)
        ws = WaveletSparsifier(adj, self.scales[1], self.approximation_order, self.tolerance, lmax=self.lmax,
                               block=self.block, device=self.device, method=self.method)
        if feature.ndim != 2 or feature.shape[0] != ws.n:
            raise ValueError("Dimension mismatch detected for the adjacency and the feature matrix!")
        self.lmax = ws.lmax
        self.ncount = ws.n
        self.phi_device = ws.calculate_all_wavelets_device(normalize=True)
        self._phi_host = None
        adjs = []
        for indptr, cols, vals in self.phi_device:
            adjs.append(DeviceAdj(DeviceCSR(indptr, cols, vals, ws.n, -1)))
        x = torch.from_numpy(feature).to(self.device)
        localized = torch.relu(wavelet_localize(adjs[0], adjs[1], x))
        self.processed_feature = torch.cat((x, localized), dim=1).cpu()
        return self.processed_feature
