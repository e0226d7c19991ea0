"""SymLaplacianGraphOp — ``A^ = D^(r-1) (A+I)^T D^(-r)`` on the GPU.

Mirror of SSRG/operators/graph_operator/symmetrical_simgraph_laplacian_operator.py:7-15; used by
every model of the reference (sgc.py:9, ssgc.py:11, sign.py:11, gamlp.py:10, gbp.py:10, nafs.py:10,
gcn.py:8).
"""
from ..base_operator import GraphOp
from ..utils import adj_to_symmetric_norm


class SymLaplacianGraphOp(GraphOp):
    def __init__(self, prop_steps, r=0.5):
        super().__init__(prop_steps)
        self.r = r

    def _norm_params(self):
        return float(self.r), None

    def construct_adj(self, adj):
        """scipy CSR in, normalised scipy CSR out (int32 indices, sorted rows, float64 data)."""
        return adj_to_symmetric_norm(adj, self.r, device=self.device)
