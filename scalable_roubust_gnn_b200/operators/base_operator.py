"""GraphOp — the drop-in boundary of the propagation path.

Mirrors ``SSRG/operators/base_operator.py:11-36`` (class GraphOp) and ``:309-314``
(ada_platform_one_step_propagation).  Same constructor, attributes, method names, argument
meaning, return type (list of ``prop_steps + 1`` CPU float32 tensors, element 0 = the input)
and error messages; the work runs in libsrgnn_b200.so on the GPU instead of scipy + OpenMP.

Differences, all supersets of the reference behaviour:
  * arguments are validated BEFORE any work (the reference normalises first, then raises);
  * the dimension check also runs for Tensor inputs (the reference's elif chain skips it);
  * ``self.adj`` is materialised lazily: propagate keeps the normalised CSR on the device and
    only copies it to a scipy matrix when somebody reads ``op.adj``.
"""
from __future__ import annotations

import ctypes

import numpy as np
import scipy.sparse as sp
import torch
from torch import Tensor

from . import utils as _u


class GraphOp:
    #: subclasses set these; they select the normalisation the device kernels apply
    _r = 0.5
    _ppr_alpha = None

    def __init__(self, prop_steps):
        self.prop_steps = prop_steps
        self._adj = None
        self._adj_source = None
        #: CUDA device ordinal used by this operator
        self.device = 0

    # -- reference attribute `adj`: normalised adjacency of the last propagate ------------------
    @property
    def adj(self):
        if self._adj is None and self._adj_source is not None:
            self._adj = self.construct_adj(self._adj_source)
            self._adj_source = None
        return self._adj

    @adj.setter
    def adj(self, value):
        self._adj = value
        self._adj_source = None

    def construct_adj(self, adj):
        raise NotImplementedError

    def _norm_params(self):
        """(r, ppr_alpha) for the fused device pipeline, or None when a subclass overrides
        construct_adj with something the library does not know."""
        return None

    def propagate(self, adj, feature, device_output=False):
        """``device_output=True`` (placement opt-in) returns the K+1 matrices as CUDA tensors that
        stay on the GPU — no device->host copies, which are 2/3 of the end-to-end time at the products
        shape; consumers index them exactly like the CPU list (``feat[idx].to(device)``, base_model.py:84-87)."""
        if not isinstance(adj, sp.csr_matrix):
            raise TypeError("The adjacency matrix must be a scipy csr sparse matrix!")
        if not isinstance(feature, np.ndarray):
            if isinstance(feature, Tensor):
                feature = feature.numpy()
            else:
                raise TypeError("The feature matrix must be a numpy.ndarray!")
        if feature.ndim != 2 or adj.shape[1] != feature.shape[0]:
            raise ValueError("Dimension mismatch detected for the adjacency and the feature matrix!")
        if feature.dtype != np.float32:
            # the reference fails inside ctypes for anything but float32 (utils.py:34,45)
            raise ctypes.ArgumentError("The feature matrix must be float32!")

        if device_output:
            return self._propagate_device(adj, feature)
        params = self._norm_params()
        if params is not None:
            r, alpha = params
            hops, _ = _u.propagate_host(adj, feature, self.prop_steps, r, alpha, device=self.device)
            self._adj, self._adj_source = None, adj
        else:
            # custom construct_adj: normalise through it, then hop by hop on the GPU
            self.adj = self.construct_adj(adj)
            hops, cur = [], np.ascontiguousarray(feature)
            for _ in range(self.prop_steps):
                cur = ada_platform_one_step_propagation(self._adj, cur)
                hops.append(torch.from_numpy(cur))
        return [torch.FloatTensor(feature)] + hops


    def _propagate_device(self, adj, feature):
        """propagate with the result left on the GPU: list of K+1 CUDA float32 tensors n x F (views of the
        padded device buffers; element 0 = the input)."""
        from .. import _lib, device as sdev
        if _lib.load().srg_device_count() <= 0:
            raise _lib.SrgError(_lib.SRG_ERR_NODEV, "no CUDA device visible: libsrgnn_b200 has no CPU fallback")
        n, f = feature.shape
        k = int(self.prop_steps)
        dev = torch.device("cuda", int(self.device))
        params = self._norm_params()
        with torch.cuda.device(dev):
            if n * f == 0:
                return [torch.zeros((n, f), dtype=torch.float32, device=dev) for _ in range(k + 1)]
            if params is None:
                # custom construct_adj: normalise through it, hop chain on the device
                self.adj = self.construct_adj(adj)
                run = _u.DeviceHopRunner([self._adj], feature, device=self.device)
                cur, hops = run.x0, [run.x0]
                for _ in range(k):
                    cur = run.hop(0, cur)
                    hops.append(cur)
                return [h[:, :f] for h in hops]
            r, alpha = params
            a_dev = sdev.upload_csr(adj, device=dev)
            x0 = sdev.pack_features(torch.from_numpy(np.ascontiguousarray(feature)).to(dev))
            norm, flags, _ = sdev.sym_norm(a_dev, r, alpha)
            fl = int(flags.item()) & ~_lib.SRG_FLAG_WEIGHTED
            if fl:
                # unsorted / directed / explicit-zero input: the host pipeline owns the retry logic
                self.adj = self.construct_adj(adj)
                run = _u.DeviceHopRunner([self._adj], feature, device=self.device)
                cur, hops = run.x0, [run.x0]
                for _ in range(k):
                    cur = run.hop(0, cur)
                    hops.append(cur)
                return [h[:, :f] for h in hops]
            hops = sdev.propagate(norm, x0, f, k)
            self._adj, self._adj_source = None, adj
            return [h[:, :f] for h in hops]

    def propagate_aggregate(self, adj, feature, msg_op):
        """``msg_op.aggregate(self.propagate(adj, feature))`` with the aggregation folded into the device
        pipeline: only the aggregate is copied back (SSRG/models/base_scalable/base_model.py:36-43 is the
        call pair this replaces).  Falls back to the two-step form for operators without a fused spec."""
        spec = msg_op.fused_spec(self.prop_steps + 1) if hasattr(msg_op, "fused_spec") else None
        params = self._norm_params()
        if spec is None or params is None:
            return msg_op.aggregate(self.propagate(adj, feature))
        if not isinstance(adj, sp.csr_matrix):
            raise TypeError("The adjacency matrix must be a scipy csr sparse matrix!")
        if isinstance(feature, Tensor):
            feature = feature.numpy()
        if not isinstance(feature, np.ndarray):
            raise TypeError("The feature matrix must be a numpy.ndarray!")
        if feature.ndim != 2 or adj.shape[1] != feature.shape[0]:
            raise ValueError("Dimension mismatch detected for the adjacency and the feature matrix!")
        if feature.dtype != np.float32:
            raise ctypes.ArgumentError("The feature matrix must be float32!")
        r, alpha = params
        out = _u.propagate_aggregate_host(adj, feature, self.prop_steps, r, alpha, spec, device=self.device)
        self._adj, self._adj_source = None, adj
        return out


def _check_inputs(adj, feature):
    """The reference's three checks (base_operator.py:22-30 and twins), run before any work."""
    if not isinstance(adj, sp.csr_matrix):
        raise TypeError("The adjacency matrix must be a scipy csr sparse matrix!")
    if not isinstance(feature, np.ndarray):
        if isinstance(feature, Tensor):
            feature = feature.numpy()
        else:
            raise TypeError("The feature matrix must be a numpy.ndarray!")
    if feature.ndim != 2 or adj.shape[1] != feature.shape[0]:
        raise ValueError("Dimension mismatch detected for the adjacency and the feature matrix!")
    if feature.dtype != np.float32:
        raise ctypes.ArgumentError("The feature matrix must be float32!")
    return feature


class _MultiAdjGraphOp:
    """Shared body of TwoOrderPprApproxGraphOp / TwoDirGraphOp: K independent hop chains, one per
    normalised adjacency, features resident on the device."""

    _adj_attrs = ()

    def __init__(self, prop_steps):
        self.prop_steps = prop_steps
        for name in self._adj_attrs:
            setattr(self, name, None)
        self.device = 0

    def construct_adj(self, adj):
        raise NotImplementedError

    def propagate(self, adj, feature):
        feature = _check_inputs(adj, feature)
        adjs = self.construct_adj(adj)
        for name, a in zip(self._adj_attrs, adjs):
            setattr(self, name, a)
        run = _u.DeviceHopRunner(adjs, feature, device=self.device)
        lists = []
        for which in range(len(adjs)):
            cur, hops = run.x0, [torch.FloatTensor(feature)]
            for _ in range(self.prop_steps):
                cur = run.hop(which, cur)
                hops.append(run.to_host(cur))
            lists.append(hops)
        return tuple(lists)


class TwoOrderPprApproxGraphOp(_MultiAdjGraphOp):
    """SSRG/operators/base_operator.py:62-96: returns (one_prop_feat_list, two_prop_feat_list)."""
    _adj_attrs = ("one_adj", "two_adj")


class TwoDirGraphOp(_MultiAdjGraphOp):
    """SSRG/operators/base_operator.py:244-284: returns (un_list, in_list, out_list)."""
    _adj_attrs = ("un_adj", "in_adj", "out_adj")


class ComGraphOp:
    """Complex (magnetic) propagation, SSRG/operators/base_operator.py:145-208.

    ``construct_adj`` yields the real and the imaginary part of the normalised operator; step k expands
    ``(R + iI)`` applied to every term of step k-1 (2^k sparse products), negates a term whenever its
    count of imaginary factors becomes even (``calculator.reversal``, :136-138) and sums the real /
    imaginary terms (``calculate_real_imag_feat``, :316-345).  ``faithful=True`` (default) reproduces the
    reference bit for bit, including the aliasing of its in-place ``+=``: the first real and the first
    imaginary term of a step become the running totals and feed the next step.  ``faithful=False``
    evaluates the recurrence the expansion stands for, ``Z_k = (R + iI) Z_{k-1}``, with 4 products per step.
    """

    def __init__(self, prop_steps, faithful=True):
        self.prop_steps = prop_steps
        self.real_adj = None
        self.imag_adj = None
        self.faithful = faithful
        self.device = 0

    def construct_adj(self, adj):
        raise NotImplementedError

    def propagate(self, adj, feature):
        feature = _check_inputs(adj, feature)
        self.real_adj, self.imag_adj = self.construct_adj(adj)
        run = _u.DeviceHopRunner([self.real_adj, self.imag_adj], feature, device=self.device)
        x0 = torch.FloatTensor(feature)
        real_list, imag_list = [x0], [x0]
        if not self.faithful:
            re, im = run.x0, None
            for step in range(self.prop_steps):
                if im is None:                                   # Z_0 = x is real
                    re, im = run.hop(0, re), run.hop(1, re)
                else:
                    re, im = run.add_(run.hop(0, re), run.neg(run.hop(1, im))), run.add_(run.hop(0, im), run.hop(1, re))
                real_list.append(run.to_host(re))
                imag_list.append(run.to_host(im))
            return real_list, imag_list
        terms = []                                               # [value, r_step, i_step]
        for step in range(self.prop_steps):
            if step == 0:
                terms = [[run.hop(0, run.x0), 1, 0], [run.hop(1, run.x0), 0, 1]]
                real_list.append(run.to_host(terms[0][0]))
                imag_list.append(run.to_host(terms[1][0]))
                continue
            out = [[run.hop(0, v), rs + 1, is_] for v, rs, is_ in terms]
            for v, rs, is_ in terms:
                t = run.hop(1, v)
                if (is_ + 1) & 1 == 0:
                    t = run.neg(t)
                out.append([t, rs, is_ + 1])
            reals = [t for t in out if t[2] & 1 == 0]
            imags = [t for t in out if t[2] & 1 == 1]
            if len(reals) != len(imags):
                raise RuntimeError("Something wrong!")
            for k in range(1, len(reals)):
                run.add_(reals[0][0], reals[k][0])
                run.add_(imags[0][0], imags[k][0])
            real_list.append(run.to_host(reals[0][0]))
            imag_list.append(run.to_host(imags[0][0]))
            terms = out
        return real_list, imag_list


def _typed_lists(names, lists):
    for name, feats in zip(names, lists):
        if any(not isinstance(feat, Tensor) for feat in feats):
            raise TypeError(f"The {name} feature matrices must be tensors!")


class ComMessageOp(torch.nn.Module):
    """SSRG/operators/base_operator.py:212-241."""

    def __init__(self, start=None, end=None):
        super().__init__()
        self.aggr_type = None
        self.start, self.end = start, end

    def combine(self, real_feat_list, imag_feat_list):
        return NotImplementedError

    def aggregate(self, real_feat_list, imag_feat_list):
        if not isinstance(real_feat_list, list) or not isinstance(imag_feat_list, list):
            return TypeError("The input must be a list consists of feature matrices!")
        _typed_lists(("real", "imag"), (real_feat_list, imag_feat_list))
        return self.combine(real_feat_list, imag_feat_list)


class TwoOrderPprApproxMessageOp(torch.nn.Module):
    """SSRG/operators/base_operator.py:99-124."""

    def __init__(self, start=None, end=None):
        super().__init__()
        self.aggr_type = None
        self.start, self.end = start, end

    def combine(self, one_feat_list, two_feat_list):
        return NotImplementedError

    def aggregate(self, one_feat_list, two_feat_list):
        if not isinstance(one_feat_list, list) or not isinstance(two_feat_list, list):
            return TypeError("The input must be a list consists of feature matrices!")
        _typed_lists(("one order", "two order"), (one_feat_list, two_feat_list))
        return self.combine(one_feat_list, two_feat_list)


class TwoDirMessageOp(torch.nn.Module):
    """SSRG/operators/base_operator.py:288-306."""

    def __init__(self, start=None, end=None):
        super().__init__()
        self.aggr_type = None
        self.start, self.end = start, end

    def combine(self, un_feat_list, in_feat_list, out_feat_list):
        return NotImplementedError

    def aggregate(self, un_feat_list, in_feat_list, out_feat_list):
        if not all(isinstance(x, list) for x in (un_feat_list, in_feat_list, out_feat_list)):
            return TypeError("The input must be a list consists of feature matrices!")
        _typed_lists(("un direction", "in direction", "out direction"), (un_feat_list, in_feat_list, out_feat_list))
        return self.combine(un_feat_list, in_feat_list, out_feat_list)


def ada_platform_one_step_propagation(adj, x):
    """One hop ``adj @ x`` (SSRG/operators/base_operator.py:309-314).  The reference switches on
    the platform between its OpenMP library and scipy; here there is one path: the GPU."""
    return _u.csr_sparse_dense_matmul(adj, x)
