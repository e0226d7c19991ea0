"""TwoDirLaplacianGraphOp — undirected / in-direction / out-direction operators of a directed graph.

Mirror of SSRG/operators/graph_operator/in_out_directed_laplacian_operator.py:7-15; the normaliser runs on the
GPU with sparse x sparse products instead of the reference's dense N x N ones.
"""
from ..base_operator import TwoDirGraphOp
from ..utils import adj_to_un_in_out_dir_symmetric_norm


class TwoDirLaplacianGraphOp(TwoDirGraphOp):
    def __init__(self, prop_steps, r=0.5):
        super().__init__(prop_steps)
        self.r = r

    def construct_adj(self, adj):
        return adj_to_un_in_out_dir_symmetric_norm(adj, self.r, device=self.device)
