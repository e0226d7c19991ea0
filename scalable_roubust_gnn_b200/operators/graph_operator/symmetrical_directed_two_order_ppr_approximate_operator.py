"""SymDirTwoOrderPprApproxGraphOp — first- and second-order PPR-approximation operators, normalised on the GPU.

Mirror of SSRG/operators/graph_operator/symmetrical_directed_two_order_ppr_approximate_operator.py:7-16.
The device normaliser is compared with the reference's own outputs in tests/test_fast_ppr.py (-m gpu).
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
