from .symmetrical_simgraph_laplacian_operator import SymLaplacianGraphOp  # noqa: F401
from .symmetrical_simgraph_ppr_operator import PprGraphOp  # noqa: F401
from .symmetrical_directed_magnetic_laplacian_operator import SymDirMagLaplacianGraphOp  # noqa: F401
from .symmetrical_directed_magnetic_comppr_operator import SymDirMagComPprGraphOp  # noqa: F401
from .in_out_directed_laplacian_operator import TwoDirLaplacianGraphOp  # noqa: F401
from .symmetrical_directed_fast_ppr_approximate_operator import SymDirFastPprApproxGraphOp  # noqa: F401
from .symmetrical_directed_two_order_ppr_approximate_operator import SymDirTwoOrderPprApproxGraphOp  # noqa: F401
