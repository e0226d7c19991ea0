"""Generate tests/golden/*.npz from the REAL reference (runs only in the build container).

    python tests/golden/make_golden.py

Imports the reference's own `operators` package from /root/reference (three absent third-party
modules that the SymLaplacian/Ppr path never touches are stubbed) and records, for small inputs,
exactly what the reference returns.  The fixtures are what pins `oracle/`; the GPU box never sees
/root/reference.  Also records the bundled real Cora / PubMed edge lists and sparsity masks
(data fixtures of the reference) that the mask tests reproduce from the torch seed.
"""
import hashlib
import os
import sys
import types

import numpy as np
import scipy.sparse as sp
import torch

REF = "/root/reference/Scalable Spectral Robust GNN"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name in ["torch_sparse", "torch_scatter", "torch_geometric", "torch_geometric.utils"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["torch_sparse"].coalesce = None
    sys.modules["torch_scatter"].scatter_add = None
    sys.modules["torch_geometric.utils"].add_self_loops = None
    sys.modules["torch_geometric.utils"].to_scipy_sparse_matrix = None
    sys.path.insert(0, REF)
    from operators.graph_operator.symmetrical_simgraph_laplacian_operator import SymLaplacianGraphOp
    from operators.graph_operator.symmetrical_simgraph_ppr_operator import PprGraphOp
    return SymLaplacianGraphOp, PprGraphOp


def sym_graph(n, m, seed, weighted=False):
    rng = np.random.default_rng(seed)
    u, v = rng.integers(0, n, m), rng.integers(0, n, m)
    keep = u != v
    u, v = u[keep], v[keep]
    w = rng.random(len(u)) + 0.5 if weighted else np.ones(len(u))
    a = sp.coo_matrix((w, (u, v)), shape=(n, n)).tocsr()
    a = a.maximum(a.T).tocsr()
    return a


def csr_pack(prefix, m, out):
    m = m.tocsr()
    m.sort_indices()
    out[prefix + "_indptr"] = m.indptr.astype(np.int32)
    out[prefix + "_indices"] = m.indices.astype(np.int32)
    out[prefix + "_data"] = m.data.astype(np.float64)


def main():
    Sym, Ppr = import_reference()
    out = {}
    cases = []

    # ---- real Cora topology (reference data fixture) ----------------------------------------
    pl = os.path.join(REF, "sparsity_datasets/simhomo/Planetoid")
    cora_e = torch.load(os.path.join(pl, "cora_0_0/raw/edge_index.pt")).numpy()
    n_cora = 2708
    out["cora_edges"] = cora_e.astype(np.int32)
    cora = sp.csr_matrix((np.ones(cora_e.shape[1]), (cora_e[0], cora_e[1])), shape=(n_cora, n_cora))
    cora = (cora + cora.T).tocsr()
    assert cora.nnz == 10556
    rng = np.random.default_rng(1)
    graphs = {
        "cora": (cora, rng.random((n_cora, 24), dtype=np.float32)),
        "rand_unw": (sym_graph(300, 1500, 2), rng.random((300, 37), dtype=np.float32)),
        "rand_w": (sym_graph(200, 900, 3, weighted=True), rng.random((200, 8), dtype=np.float32)),
    }
    # an undirected graph with an existing self loop and an isolated node
    g = sym_graph(64, 200, 4).tolil()
    g[5, 5] = 1.0
    g[9, :] = 0
    g[:, 9] = 0
    g = g.tocsr()
    g.eliminate_zeros()
    graphs["loop_iso"] = (g, rng.random((64, 5), dtype=np.float32))

    for name, (a, x) in graphs.items():
        csr_pack(f"{name}_adj", a, out)
        out[f"{name}_x"] = x
        for r in (0.5, 0.0, 0.3, 1.0):
            op = Sym(3, r=r)
            hops = op.propagate(a, x)
            tag = f"{name}_r{r}"
            csr_pack(f"{tag}_norm", op.adj, out)
            out[f"{tag}_hop3"] = hops[3].numpy()
            out[f"{tag}_hop1"] = hops[1].numpy()
            cases.append(tag)
        op = Ppr(2, r=0.5, alpha=0.15)
        hops = op.propagate(a, x)
        csr_pack(f"{name}_ppr_norm", op.adj, out)
        out[f"{name}_ppr_hop2"] = hops[2].numpy()

    # asymmetric adjacency with a duplicate entry (general transpose path)
    rows = np.array([0, 0, 1, 2, 2, 3, 3, 4, 0])
    cols = np.array([1, 2, 2, 0, 3, 3, 4, 1, 1])
    asym = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(6, 6))
    xa = rng.random((6, 4), dtype=np.float32)
    csr_pack("asym_adj", asym, out)
    out["asym_x"] = xa
    for r in (0.5, 0.3):
        op = Sym(2, r=r)
        hops = op.propagate(asym, xa)
        csr_pack(f"asym_r{r}_norm", op.adj, out)
        out[f"asym_r{r}_hop2"] = hops[2].numpy()

    np.savez_compressed(os.path.join(HERE, "reference_propagation.npz"), **out)

    # ---- message operators of the reference on a reference hop list ------------------------------
    from operators.message_operator.concat_message_op import ConcatMessageOp
    from operators.message_operator.last_message_op import LastMessageOp
    from operators.message_operator.max_message_op import SimMaxMessageOp
    from operators.message_operator.mean_message_op import MeanMessageOp
    from operators.message_operator.min_message_op import SimMinMessageOp
    from operators.message_operator.simple_weighted_message_op import SimpleWeightedMessageOp
    from operators.message_operator.sum_message_op import SumMessageOp
    mo = {}
    a, x = graphs["rand_unw"]
    hops = Sym(3, r=0.5).propagate(a, x)
    ops = {"last": LastMessageOp(), "mean": MeanMessageOp(0, 4), "sum": SumMessageOp(0, 4), "sum13": SumMessageOp(1, 3),
           "max": SimMaxMessageOp(0, 4), "min": SimMinMessageOp(0, 4), "concat": ConcatMessageOp(0, 4),
           "concat24": ConcatMessageOp(2, 4), "alpha": SimpleWeightedMessageOp(0, 4, "alpha", 0.5),
           "hand": SimpleWeightedMessageOp(0, 4, "hand_crafted", [0.1, 0.2, 0.3, 0.4])}
    for name, op in ops.items():
        mo[name] = op.aggregate(list(hops)).numpy()
    np.savez_compressed(os.path.join(HERE, "reference_message_ops.npz"), **mo)

    # ---- sparsity masks (reference data fixtures) ----------------------------------------------
    masks = {}
    for ds, n_feat_shape in [("cora_0_0.7", (2708, 1433)), ("cora_0.7_0.7", (2708, 1433)),
                             ("pubmed_0.6_0.6", (19717, 500)), ("citeseer_0.5_0.5", (3327, 3703))]:
        em = torch.load(os.path.join(pl, ds, "raw/edge_mask.pt")).numpy()
        ei = torch.load(os.path.join(pl, ds, "raw/edge_index.pt")).numpy()
        key = ds.replace(".", "p")
        masks[key + "_edge_mask_head"] = em[:64].astype(np.int64)
        masks[key + "_edge_mask_len"] = np.array([len(em)])
        masks[key + "_edge_mask_sha"] = np.frombuffer(hashlib.sha256(em.astype(np.int64).tobytes()).digest(), dtype=np.uint8)
        masks[key + "_edge_index_sha"] = np.frombuffer(hashlib.sha256(ei.astype(np.int64).tobytes()).digest(), dtype=np.uint8)
        masks[key + "_shape"] = np.array(n_feat_shape)
    pub = torch.load(os.path.join(pl, "pubmed_0.6_0.6/raw/edge_index.pt")).numpy()
    masks["pubmed_0p6_0p6_edge_index"] = pub.astype(np.int32)
    cite_full = None
    np.savez_compressed(os.path.join(HERE, "reference_masks.npz"), **masks)
    print("wrote", len(out), "+", len(masks), "arrays;", "cases:", cases)


if __name__ == "__main__":
    main()
