"""SymDirMagComPprGraphOp — complex personalised PageRank on the magnetic operator.

Mirror of SSRG/operators/graph_operator/symmetrical_directed_magnetic_comppr_operator.py:25-38:
``real = (1 - alpha) * real + alpha * I``, ``imag = (1 - alpha) * imag``; the blend runs inside the
normalisation pipeline on the GPU (the pattern always contains the diagonal).
"""
from ..base_operator import ComGraphOp
from ..utils import adj_to_directed_symmetric_mag_norm


class SymDirMagComPprGraphOp(ComGraphOp):
    def __init__(self, prop_steps, r=0.5, q=0.25, ppr_alpha=0.15, faithful=True):
        super().__init__(prop_steps, faithful=faithful)
        self.r = r
        self.q = q
        self.ppr_alpha = ppr_alpha

    def construct_adj(self, adj):
        return adj_to_directed_symmetric_mag_norm(adj, self.r, self.q, ppr_alpha=self.ppr_alpha, device=self.device)
