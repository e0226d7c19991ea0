"""SymDirMagLaplacianGraphOp — magnetic Laplacian of a directed graph, normalised on the GPU.

Mirror of SSRG/operators/graph_operator/symmetrical_directed_magnetic_laplacian_operator.py:7-16.
"""
from ..base_operator import ComGraphOp
from ..utils import adj_to_directed_symmetric_mag_norm


class SymDirMagLaplacianGraphOp(ComGraphOp):
    def __init__(self, prop_steps, r=0.5, q=0.25, faithful=True):
        super().__init__(prop_steps, faithful=faithful)
        self.r = r
        self.q = q

    def construct_adj(self, adj):
        return adj_to_directed_symmetric_mag_norm(adj, self.r, self.q, device=self.device)
