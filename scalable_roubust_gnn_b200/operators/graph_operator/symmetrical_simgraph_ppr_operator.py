"""PprGraphOp — ``(1 - alpha) * A^ + alpha * I`` on the GPU.

Mirror of SSRG/operators/graph_operator/symmetrical_simgraph_ppr_operator.py:7-21.  The blend is
applied inside the normalisation kernel (A^ always has a full diagonal, so the pattern is
unchanged).
"""
import scipy.sparse as sp

from ..base_operator import GraphOp
from ..utils import adj_to_symmetric_norm


class PprGraphOp(GraphOp):
    def __init__(self, prop_steps, r=0.5, alpha=0.15):
        super().__init__(prop_steps)
        self.r = r
        self.alpha = alpha

    def _norm_params(self):
        return float(self.r), float(self.alpha)

    def construct_adj(self, adj):
        if isinstance(adj, sp.coo_matrix):
            adj = adj.tocsr()
        elif not isinstance(adj, sp.csr_matrix):
            raise TypeError("The adjacency matrix must be a scipy.sparse.coo_matrix/csr_matrix!")
        return adj_to_symmetric_norm(adj, self.r, ppr_alpha=self.alpha, device=self.device)
