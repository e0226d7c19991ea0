"""Drive the reference's own operator API from baseline/_ref (stock code path, nothing of this repository on it)."""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(DST, "operators/base_operator.py"))


def restore_openmp_threads() -> int:
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the reference's matmul.c is an OpenMP code that
    uses all host threads when run on its own.  Must be called before libgomp is loaded; also sets the count through
    the runtime in case some import got there first."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import ctypes
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(n)
        gomp.omp_get_max_threads.restype = ctypes.c_int
        return int(gomp.omp_get_max_threads())
    except OSError:
        return n


def import_reference():
    """The reference's SymLaplacianGraphOp / PprGraphOp and csr_sparse_dense_matmul, imported from baseline/_ref.
    torch_sparse / torch_scatter / torch_geometric are imported by operators/utils.py:10-14 but never touched on
    the SymLaplacian / Ppr path; they are absent from the image, so empty modules stand in for them."""
    for name in ["torch_sparse", "torch_scatter", "torch_geometric", "torch_geometric.utils"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["torch_sparse"].coalesce = None
    sys.modules["torch_scatter"].scatter_add = None
    sys.modules["torch_geometric.utils"].add_self_loops = None
    sys.modules["torch_geometric.utils"].to_scipy_sparse_matrix = None
    if DST not in sys.path:
        sys.path.insert(0, DST)
    from operators.graph_operator.symmetrical_simgraph_laplacian_operator import SymLaplacianGraphOp
    from operators.graph_operator.symmetrical_simgraph_ppr_operator import PprGraphOp
    from operators.utils import csr_sparse_dense_matmul
    return SymLaplacianGraphOp, PprGraphOp, csr_sparse_dense_matmul
