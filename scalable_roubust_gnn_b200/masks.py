"""Sparsity masks of the reference's dataset pipeline, applied on the GPU.

Mirrors SSRG/data_process.py:35-41 (featureMasked), :43-67 (edgeMasked), the application
``x * feature_mask`` (SSRG/data_augument.py:28) and the symmetrise + unique adjacency rebuild
(SSRG/data_augument.py:99-102).

The masks themselves are DEFINED by torch's CPU RNG stream (``torch.rand`` then
``torch.randperm`` after ``seed_everything(2023)``), so they are drawn on the host exactly as the
reference draws them — that is a definition, not a fallback; gathering the edges, rebuilding the
CSR and masking the features run as kernels of libsrgnn_b200.so.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib
from .device import DeviceCSR, _p, _stream_ptr, pack_features

__all__ = ["feature_mask", "edge_keep_index", "upper_edges", "edge_gather", "edges_to_sym_csr",
           "apply_feature_mask", "masked_graph"]


def feature_mask(shape, rate) -> torch.Tensor:
    """``(torch.rand(shape) > rate).int()`` — data_process.py:38-39 (consumes the RNG stream)."""
    return (torch.rand(shape) > rate).int()


def edge_keep_index(num_upper_edges: int, rate: float) -> torch.Tensor:
    """``torch.randperm(E)[int(E * rate):]`` — data_process.py:55,65."""
    return torch.randperm(num_upper_edges)[int(num_upper_edges * rate):]


def upper_edges(adj) -> torch.Tensor:
    """edges with col > row of a scipy adjacency in row-major order (data_process.py:48-53), int64 2 x E."""
    coo = sp.coo_matrix(adj)
    keep = coo.col > coo.row
    return torch.from_numpy(np.stack([coo.row[keep], coo.col[keep]]).astype(np.int64))


def edge_gather(edge_index: torch.Tensor, keep: torch.Tensor) -> torch.Tensor:
    """``edge_index[:, keep]`` on the device (int64, bit-exact)."""
    lib = _lib.load()
    assert edge_index.is_cuda and keep.is_cuda and edge_index.dtype == torch.int64 and keep.dtype == torch.int64
    edge_index = edge_index.contiguous()
    keep = keep.contiguous()
    e, ek = edge_index.shape[1], keep.numel()
    out = torch.empty((2, ek), dtype=torch.int64, device=edge_index.device)
    flags = torch.zeros(1, dtype=torch.int32, device=edge_index.device)
    _lib.check(lib.srg_edge_gather_i64(_p(edge_index), e, _p(keep), ek, _p(out), _p(flags), _stream_ptr(edge_index.device)))
    if int(flags.item()) & _lib.SRG_FLAG_BAD_INDEX:
        raise IndexError("edge mask index out of range")
    return out


def edges_to_sym_csr(edge_index: torch.Tensor, n: int) -> DeviceCSR:
    """Undirected duplicate-free adjacency of an edge list as a pattern CSR (all values 1.0)."""
    lib = _lib.load()
    assert edge_index.is_cuda and edge_index.dtype == torch.int64
    edge_index = edge_index.contiguous()
    e = edge_index.shape[1]
    dev = edge_index.device
    indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    indices = torch.empty(max(2 * e, 1), dtype=torch.int32, device=dev)
    nnz_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.srg_edges_to_sym_csr(_p(edge_index), e, n, _p(indptr), _p(indices), _p(nnz_dev), _p(flags),
                                        _stream_ptr(dev)))
    if int(flags.item()) & _lib.SRG_FLAG_BAD_INDEX:
        raise IndexError("edge endpoint outside [0, n)")
    return DeviceCSR(indptr, indices, None, n, int(nnz_dev.item()))


def apply_feature_mask(x: torch.Tensor, mask: torch.Tensor, ld: int | None = None) -> torch.Tensor:
    """``x * feature_mask`` into the padded device layout (one pass)."""
    return pack_features(x, ld=ld, mask=mask)


def masked_graph(adj, feature_shape, feature_rate, edge_rate, device="cuda"):
    """The reference's dataset sparsification (data_process.py:108-117 order: feature mask first,
    then the edge permutation) with the edge gather + CSR rebuild on the GPU.

    Returns (feature_mask int32 CPU tensor, keep index, gathered edge_index (cuda), DeviceCSR).
    Seed the torch RNG (``torch.manual_seed(2023)``) before calling, as seed_everything does.
    """
    fmask = feature_mask(feature_shape, feature_rate)
    up = upper_edges(adj)
    keep = edge_keep_index(up.shape[1], edge_rate)
    gathered = edge_gather(up.to(device), keep.to(device))
    csr = edges_to_sym_csr(gathered, adj.shape[0])
    return fmask, keep, gathered, csr
