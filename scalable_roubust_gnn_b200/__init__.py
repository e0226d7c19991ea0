"""scalable_roubust_gnn_b200 — B200 (sm_100a) implementation of the K-hop feature-propagation
path of yyysyyy/Scalable-Roubust-GNN behind the reference's ``operators/`` API.

Layout (only what the path needs):
  csrc/        hand-written CUDA kernels + the C ABI (libsrgnn_b200.so, include/srgnn_b200.h)
  operators/   host-side mirror of SSRG/operators (GraphOp, SymLaplacianGraphOp, PprGraphOp, utils)
  device.py    the same stages on device-resident torch tensors (no host copies)
"""
from . import _lib  # noqa: F401
from ._lib import SrgError, SrgUnsupported, build, load  # noqa: F401

__version__ = "0.1.0"
