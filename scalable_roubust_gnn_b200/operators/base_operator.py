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

    def propagate(self, adj, feature):
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


def ada_platform_one_step_propagation(adj, x):
    """One hop ``adj @ x`` (SSRG/operators/base_operator.py:309-314).  The reference switches on
    the platform between its OpenMP library and scipy; here there is one path: the GPU."""
    return _u.csr_sparse_dense_matmul(adj, x)
