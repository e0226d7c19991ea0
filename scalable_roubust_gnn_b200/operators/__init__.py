"""Mirror of the reference package ``SSRG/operators`` for the propagation path."""
from .base_operator import GraphOp, ada_platform_one_step_propagation  # noqa: F401
from .graph_operator import PprGraphOp, SymLaplacianGraphOp  # noqa: F401
from .utils import adj_to_symmetric_norm, csr_sparse_dense_matmul  # noqa: F401
from .message_operator import (ConcatMessageOp, LastMessageOp, MeanMessageOp, MessageOp,  # noqa: F401
                               OverSmoothDistanceWeightedOp, SimMaxMessageOp, SimMinMessageOp,
                               SimpleWeightedMessageOp, SumMessageOp)
