"""Non-learnable message operators (hop-list aggregators) with a fused device path.

Mirror of SSRG/operators/message_operator/{last,sum,mean,max,min,concat,simple_weighted}_message_op.py,
over_smooth_distance_op.py (NAFS)
and of the ``MessageOp`` base (SSRG/operators/base_operator.py:40-59): same class names, constructor
arguments, ``aggr_type`` strings and ``aggregate(feat_list)`` / ``combine(feat_list)`` contract on a
list of CPU tensors.  In addition every operator describes itself to the library (``fused_spec``)
so ``GraphOp.propagate_aggregate(adj, feature, msg_op)`` can fold the hops into the result on the
GPU and copy back only the aggregate.
"""
from .simple_ops import (ConcatMessageOp, LastMessageOp, MeanMessageOp, MessageOp, SimMaxMessageOp,  # noqa: F401
                         SimMinMessageOp, SimpleWeightedMessageOp, SumMessageOp)
from .over_smooth_distance_op import OverSmoothDistanceWeightedOp, nafs_combine_device  # noqa: F401
