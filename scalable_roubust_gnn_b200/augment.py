"""Dataset augmentation that precedes the propagation path, on the GPU.

Mirror of ``edge_augument`` (SSRG/data_augument.py:73-103): every node whose endpoint count is below
``degree_level`` gets ``degree_level - count`` new neighbours — the closest in soft-label L2 distance among
``100 x deficit`` random candidates — and the edge list is symmetrised and de-duplicated.

The candidate draws are DEFINED by Python's ``random`` stream (``random.sample`` on a list the reference mutates
between draws, SSRG/utils.py:29-33), so they are made on the host exactly as the reference makes them — a
definition, not a fallback (the same rule as the torch-RNG sparsity masks, masks.py).  Counting the endpoints,
the distances, the top-k choice, symmetrise + unique and the final edge list are kernels of libsrgnn_b200.so
(csrc/augment.cu, csrc/coo.cu).
"""
from __future__ import annotations

import random

import numpy as np
import torch

from . import _lib
from .device import _p, _stream_ptr
from .masks import edges_to_sym_csr

__all__ = ["endpoint_counts", "low_degree_order", "draw_candidates", "edge_augument"]

_I64_MAX = np.iinfo(np.int64).max


def endpoint_counts(edge_row: torch.Tensor, edge_col: torch.Tensor, n: int):
    """``Counter(cat(edge_row, edge_col))`` on the device: (counts int32[n], first position int64[n])."""
    lib = _lib.load()
    assert edge_row.is_cuda and edge_col.is_cuda and edge_row.dtype == torch.int64 and edge_col.dtype == torch.int64
    edge_row, edge_col = edge_row.contiguous(), edge_col.contiguous()
    dev = edge_row.device
    counts = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    first = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.srg_endpoint_counts_i64(_p(edge_row), _p(edge_col), edge_row.numel(), n, _p(counts), _p(first), _p(flags),
                                           _stream_ptr(dev)))
    if int(flags.item()) & _lib.SRG_FLAG_BAD_INDEX:
        raise IndexError("edge endpoint outside [0, n)")
    return counts[:n], first[:n]


def low_degree_order(counts: torch.Tensor, first_pos: torch.Tensor, degree_level: int):
    """Nodes below ``degree_level`` in the order ``sorted(counts.items(), key=degree)`` visits them
    (data_augument.py:78-85): ascending count, ties in the Counter's insertion order — first appearance in
    cat(row, col), then the never-seen nodes in index order.  Returns (nodes int64 ndarray, their counts)."""
    low = torch.nonzero(counts < degree_level).flatten()
    nodes = low.cpu().numpy()
    cnt = counts[low].cpu().numpy().astype(np.int64)
    fp = first_pos[low].cpu().numpy()
    # insertion key: position for seen nodes; unseen nodes are appended after every seen one, by index
    key = np.where(fp == _I64_MAX, np.int64(2) ** 62 + nodes, fp)
    order = np.lexsort((key, cnt))
    return nodes[order], cnt[order]


def draw_candidates(nodes, counts, n: int, degree_level: int, rng=random):
    """The reference's candidate draws (generate_numbers, SSRG/utils.py:29-33) for the nodes in visiting order:
    ``numbers.remove(node); random.sample(numbers, 100 * deficit); numbers.append(node)`` on ONE list that keeps
    the re-ordering of earlier draws.  Returns a list of int arrays."""
    numbers = list(range(n))
    out = []
    for node, c in zip(nodes.tolist(), counts.tolist()):
        numbers.remove(node)
        out.append(np.asarray(rng.sample(numbers, (degree_level - c) * 100), dtype=np.int32))
        numbers.append(node)
    return out


def edge_augument(dataset, soft_label, degree_level: int = 1, device="cuda", rng=random) -> torch.Tensor:
    """``edge_augument(dataset, soft_label)`` (data_augument.py:73-103) -> int64 ``edge_index`` 2 x E, on the CPU like
    the reference's.  ``dataset.edge.row`` / ``.col``: int64 tensors; ``dataset.x.shape[0]``: node count;
    ``degree_level``: data_augument_args.degree_level (configs/data_augument_config.py:17, default 1)."""
    lib = _lib.load()
    n = int(dataset.x.shape[0])
    row = torch.as_tensor(dataset.edge.row, dtype=torch.int64).to(device)
    col = torch.as_tensor(dataset.edge.col, dtype=torch.int64).to(device)
    soft = torch.as_tensor(soft_label, dtype=torch.float32).to(device).contiguous()
    s = _stream_ptr(row.device)
    counts, first = endpoint_counts(row, col, n)
    nodes, cnts = low_degree_order(counts, first, degree_level)
    new_src = new_dst = None
    if len(nodes):
        cands = draw_candidates(nodes, cnts, n, degree_level, rng)
        deficit = (degree_level - cnts).astype(np.int32)
        c_max = int(max(len(c) for c in cands))
        cand_mat = np.zeros((len(nodes), c_max), dtype=np.int32)
        for i, c in enumerate(cands):
            cand_mat[i, :len(c)] = c
        cand_cnt = np.asarray([len(c) for c in cands], dtype=np.int32)
        off = np.concatenate([[0], np.cumsum(deficit)]).astype(np.int32)
        total = int(off[-1])
        d_nodes = torch.from_numpy(nodes.astype(np.int32)).to(device)
        d_cand = torch.from_numpy(cand_mat).to(device)
        d_cnt = torch.from_numpy(cand_cnt).to(device)
        d_k = torch.from_numpy(deficit).to(device)
        d_off = torch.from_numpy(off).to(device)
        new_src = torch.empty(max(total, 1), dtype=torch.int64, device=device)
        new_dst = torch.empty(max(total, 1), dtype=torch.int64, device=device)
        _lib.check(lib.srg_candidate_topk_f32(_p(soft), soft.stride(0), n, soft.shape[1], _p(d_nodes), _p(d_cand), _p(d_cnt),
                                              _p(d_k), _p(d_off), c_max, len(nodes), _p(new_src), _p(new_dst), s))
        new_src, new_dst = new_src[:total], new_dst[:total]
    if new_src is not None:
        edge_index = torch.stack([torch.cat([row, new_src]), torch.cat([col, new_dst])])
    else:
        edge_index = torch.stack([row, col])
    csr = edges_to_sym_csr(edge_index, n)                      # cat both directions + unique (:97-102)
    out = torch.empty((2, max(csr.nnz, 1)), dtype=torch.int64, device=device)
    _lib.check(lib.srg_csr_to_edge_index_i64(_p(csr.indptr), _p(csr.indices), n, csr.nnz, _p(out), s))
    return out[:, :csr.nnz].cpu()
