from .symmetrical_simgraph_laplacian_operator import SymLaplacianGraphOp  # noqa: F401
from .symmetrical_simgraph_ppr_operator import PprGraphOp  # noqa: F401
