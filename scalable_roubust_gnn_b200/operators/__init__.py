"""Mirror of the reference package ``SSRG/operators`` for the propagation path."""
from .base_operator import (ComGraphOp, ComMessageOp, GraphOp, TwoDirGraphOp, TwoDirMessageOp,  # noqa: F401
                            TwoOrderPprApproxGraphOp, TwoOrderPprApproxMessageOp, ada_platform_one_step_propagation)
from .graph_operator import (PprGraphOp, SymDirMagComPprGraphOp, SymDirMagLaplacianGraphOp,  # noqa: F401
                             SymLaplacianGraphOp, TwoDirLaplacianGraphOp)
from .utils import (adj_to_directed_symmetric_mag_norm, adj_to_symmetric_norm,  # noqa: F401
                    adj_to_un_in_out_dir_symmetric_norm, csr_sparse_dense_matmul)
from .message_operator import (ConcatMessageOp, LastMessageOp, MeanMessageOp, MessageOp,  # noqa: F401
                               OverSmoothDistanceWeightedOp, SimMaxMessageOp, SimMinMessageOp,
                               SimpleWeightedMessageOp, SumMessageOp)
