from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from ... import _lib


class MessageOp(nn.Module):
    """Base of the aggregators (SSRG/operators/base_operator.py:40-59)."""

    aggr_type = None

    def __init__(self, start=None, end=None):
        super().__init__()
        self.start, self.end = start, end

    def combine(self, feat_list):
        raise NotImplementedError

    def aggregate(self, feat_list):
        if not isinstance(feat_list, list):
            return TypeError("The input must be a list consists of feature matrices!")
        if any(not isinstance(feat, Tensor) for feat in feat_list):
            raise TypeError("The feature matrices must be tensors!")
        return self.combine(feat_list)

    # -- fused device path -----------------------------------------------------------------------
    def fused_spec(self, n_hops: int):
        """(agg_mode, start, end, weights|None) over a list of ``n_hops`` matrices, or None when the
        operator has no fused form."""
        return None

    def _window(self, n_hops):
        lo, hi, _ = slice(self.start, self.end).indices(n_hops)
        return lo, hi


class LastMessageOp(MessageOp):
    aggr_type = "last"

    def __init__(self):
        super().__init__()

    def combine(self, feat_list):
        return feat_list[-1]

    def fused_spec(self, n_hops):
        return _lib.SRG_AGG_LAST, n_hops - 1, n_hops, None


class SumMessageOp(MessageOp):
    aggr_type = "sum"

    def combine(self, feat_list):
        return sum(feat_list[self.start:self.end])

    def fused_spec(self, n_hops):
        lo, hi = self._window(n_hops)
        return _lib.SRG_AGG_SUM, lo, hi, None


class MeanMessageOp(MessageOp):
    aggr_type = "mean"

    def combine(self, feat_list):
        return sum(feat_list[self.start:self.end]) / (self.end - self.start)

    def fused_spec(self, n_hops):
        lo, hi = self._window(n_hops)
        if hi - lo != self.end - self.start:      # the reference divides by end - start whatever the slice holds
            return None
        return _lib.SRG_AGG_MEAN, lo, hi, None


class SimMaxMessageOp(MessageOp):
    aggr_type = "max"

    def combine(self, feat_list):
        return torch.stack(feat_list[self.start:self.end], dim=0).max(dim=0)[0]

    def fused_spec(self, n_hops):
        lo, hi = self._window(n_hops)
        return _lib.SRG_AGG_MAX, lo, hi, None


class SimMinMessageOp(MessageOp):
    aggr_type = "min"

    def combine(self, feat_list):
        return torch.stack(feat_list[self.start:self.end], dim=0).min(dim=0)[0]

    def fused_spec(self, n_hops):
        lo, hi = self._window(n_hops)
        return _lib.SRG_AGG_MIN, lo, hi, None


class ConcatMessageOp(MessageOp):
    aggr_type = "concat"

    def combine(self, feat_list):
        return torch.hstack(feat_list[self.start:self.end])

    def fused_spec(self, n_hops):
        lo, hi = self._window(n_hops)
        return _lib.SRG_AGG_CONCAT, lo, hi, None


class SimpleWeightedMessageOp(MessageOp):
    """alpha-decay (`alpha * (1-alpha)^k`) or hand-crafted hop weights
    (SSRG/operators/message_operator/simple_weighted_message_op.py:8-58)."""

    aggr_type = "simple_weighted"

    def __init__(self, start, end, combination_type, *args):
        super().__init__(start, end)
        if combination_type not in ("alpha", "hand_crafted"):
            raise ValueError("Invalid weighted combination type! Type must be 'alpha' or 'hand_crafted'.")
        if len(args) != 1:
            raise ValueError("Invalid parameter numbers for the simple weighted aggregator!")
        self.combination_type = combination_type
        self.alpha, self.weight_list = None, None
        if combination_type == "alpha":
            self.alpha = args[0]
            if not isinstance(self.alpha, float):
                raise TypeError("The alpha must be a float!")
            if self.alpha > 1 or self.alpha < 0:
                raise ValueError("The alpha must be a float in [0,1]!")
        else:
            wl = args[0]
            if isinstance(wl, list):
                wl = torch.FloatTensor(wl)
            elif not isinstance(wl, Tensor):
                raise TypeError("The input weight list must be a list or a tensor!")
            self.weight_list = wl

    def _weights(self, n_hops):
        if self.combination_type == "alpha":
            w = [self.alpha]
            for _ in range(n_hops - 1):
                w.append((1 - self.alpha) * w[-1])
            return torch.FloatTensor(w[self.start:self.end])
        return self.weight_list

    def combine(self, feat_list):
        feats = feat_list[self.start:self.end]
        w = self._weights(len(feat_list))
        if len(feats) != w.shape[0]:
            raise ValueError("The feature list and the weight list have different lengths!")
        if w.dim() != 1:
            raise ValueError("The weight list should be a 1d tensor!")
        stacked = torch.vstack([f.reshape(1, -1).squeeze(0) for f in feats])
        return (stacked * w.view(-1, 1)).sum(dim=0).view(feats[0].shape)

    def fused_spec(self, n_hops):
        lo, hi = self._window(n_hops)
        w = self._weights(n_hops)
        if w.dim() != 1 or w.shape[0] != hi - lo:
            return None
        return _lib.SRG_AGG_WEIGHTED, lo, hi, w.to(torch.float32).contiguous()
