"""OverSmoothDistanceWeightedOp — the NAFS aggregator, evaluated on the GPU.

Mirror of SSRG/operators/message_operator/over_smooth_distance_op.py:6-33 (used by
SSRG/models/nafs.py:12): every node weights its hop features by the softmax of the cosine
similarity between the hop feature and the node's input feature.  The reference forms the weighted
sum in an O(N * hops) Python loop over rows; ``combine`` here uploads the hop list and runs one
kernel (libsrgnn_b200.so:srg_nafs_combine_f32, one warp per row).  ``fused_spec`` lets
``GraphOp.propagate_aggregate`` evaluate it on the device-resident hops without copying them back.
There is no CPU path: without a CUDA device ``combine`` raises ``SrgError``.
"""
from __future__ import annotations

import ctypes as C

import torch

from ... import _lib
from .simple_ops import MessageOp


def nafs_combine_device(hops, f=None, out=None, want_weights=False):
    """Device entry point: ``hops`` is a list of cuda float32 matrices n x ld with a common row stride
    (hops[0] = input features); ``f`` the logical width.  Returns ``out`` (n x ld, columns >= f
    untouched) or ``(out, weights)`` with the n x len(hops) softmax weights."""
    lib = _lib.load()
    x0 = hops[0]
    n, width = x0.shape
    ld = x0.stride(0)
    f = width if f is None else f
    for h in hops:
        if not (h.is_cuda and h.dtype == torch.float32 and h.shape == x0.shape and h.stride(1) == 1
                and h.stride(0) == ld):
            raise ValueError("nafs_combine_device: hop matrices must be cuda float32 with one common layout")
    if out is None:
        out = torch.zeros_like(x0) if f < width else torch.empty_like(x0)
    weights = torch.empty((n, len(hops)), dtype=torch.float32, device=x0.device) if want_weights else None
    ptrs = (C.c_void_p * len(hops))(*[h.data_ptr() for h in hops])
    stream = C.c_void_p(torch.cuda.current_stream(x0.device).cuda_stream)
    _lib.check(lib.srg_nafs_combine_f32(ptrs, len(hops), ld, n, f, C.c_void_p(out.data_ptr()), out.stride(0),
                                        None if weights is None else C.c_void_p(weights.data_ptr()), stream))
    return (out, weights) if want_weights else out


class OverSmoothDistanceWeightedOp(MessageOp):
    aggr_type = "over_smooth_dis_weighted"

    def __init__(self):
        super().__init__()

    def combine(self, feat_list):
        lib = _lib.load()
        if lib.srg_device_count() <= 0:
            raise _lib.SrgError(_lib.SRG_ERR_NODEV, "no CUDA device visible: libsrgnn_b200 has no CPU fallback")
        shape = feat_list[0].shape
        if any(f.shape != shape or f.dim() != 2 for f in feat_list):
            raise ValueError("The feature matrices must share one 2-d shape!")
        from ...device import pack_features, unpack_features
        width = shape[1]
        if shape[0] * width == 0:
            return torch.zeros(shape, dtype=torch.float32)
        # the padded device layout (ld % 8 == 0): the same kernel path propagate_aggregate takes
        dev = [pack_features(f.to(device="cuda", dtype=torch.float32).contiguous()) for f in feat_list]
        return unpack_features(nafs_combine_device(dev, f=width), width).cpu()

    def fused_spec(self, n_hops):
        return _lib.SRG_AGG_NAFS, 0, n_hops, None
