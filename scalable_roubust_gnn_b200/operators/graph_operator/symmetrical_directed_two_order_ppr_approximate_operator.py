"""SymDirTwoOrderPprApproxGraphOp — first- and second-order PPR-approximation operators, normalised on the GPU.

Mirror of SSRG/operators/graph_operator/symmetrical_directed_two_order_ppr_approximate_operator.py:7-16.
Round-1 status: the device normaliser has not run on hardware yet (its test is opt-in).
"""
from ..base_operator import TwoOrderPprApproxGraphOp
from ..utils import adj_to_slow_first_second_ppr_approx_symmetric_norm


class SymDirTwoOrderPprApproxGraphOp(TwoOrderPprApproxGraphOp):
    def __init__(self, prop_steps, r=0.5, ppr_alpha=0.1):
        super().__init__(prop_steps)
        self.r = r
        self.ppr_alpha = ppr_alpha

    def construct_adj(self, adj):
        return adj_to_slow_first_second_ppr_approx_symmetric_norm(adj, self.r, self.ppr_alpha, device=self.device)
