"""Generate tests/golden/reference_ext.npz from the REAL reference (build container only).

    python tests/golden/make_golden_ext.py

Second batch of fixtures (the "next" rows of SURVEY.md §8f), same recipe as make_golden.py: the
reference's own classes imported from /root/reference, run on small inputs, outputs recorded.
  * nafs_*   OverSmoothDistanceWeightedOp.aggregate on a reference hop list
             (SSRG/operators/message_operator/over_smooth_distance_op.py)
  * gcn_*    one forward + backward of the reference's Layer2GraphConvolution
             (SSRG/models/base_scalable/simple_models.py:214-240) on the adjacency produced by
             SymLaplacianGraphOp.construct_adj + scipy_sparse_mat_to_torch_sparse_tensor
             (SSRG/models/utils.py:5-15), dropout off: outputs, loss, parameter and input gradients
  * mag_*    SymDirMagLaplacianGraphOp / SymDirMagComPprGraphOp (SSRG/operators/graph_operator/
             symmetrical_directed_magnetic_*.py, ComGraphOp.propagate in base_operator.py:145-208) on small
             directed graphs.  torch_sparse.coalesce / torch_scatter.scatter_add are absent from this image:
             the script supplies pure-torch stand-ins with their documented semantics (sum of equal keys in
             stored order, output sorted by key / sequential scatter sum) — everything else is reference code.
  * twodir_* TwoDirLaplacianGraphOp (in_out_directed_laplacian_operator.py, utils.py:195-260), add_self_loops stand-in
  * fastppr_* / twoorder_*  SymDirFastPprApproxGraphOp, SymDirTwoOrderPprApproxGraphOp (utils.py:262-424): fixtures for
             the oracle restatement; the device implementations are not built yet
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import csr_pack, import_reference, sym_graph  # noqa: E402


def main():
    Sym, _ = import_reference()
    out = {}
    rng = np.random.default_rng(11)

    # ---- NAFS aggregator ------------------------------------------------------------------------
    from operators.message_operator.over_smooth_distance_op import OverSmoothDistanceWeightedOp
    a = sym_graph(300, 1500, 2)
    x = rng.random((300, 37), dtype=np.float32) - 0.3          # mixed signs
    x[7] = 0.0                                                 # an all-zero input row (norm + 1e-10 path)
    hops = Sym(3, r=0.5).propagate(a, x)
    out["nafs_x"] = x
    out["nafs_hop3"] = hops[3].numpy()
    out["nafs_out"] = OverSmoothDistanceWeightedOp().aggregate(list(hops)).numpy()
    # a second list that is not a propagation (weights far from uniform)
    feats = [rng.standard_normal((50, 9)).astype(np.float32) * s for s in (1.0, 3.0, 0.2, 1.5, 1.0)]
    import torch
    out["nafs2_feats"] = np.stack(feats)
    out["nafs2_out"] = OverSmoothDistanceWeightedOp().aggregate([torch.from_numpy(f) for f in feats]).numpy()

    # ---- per-epoch sparse product of the GCN model (forward + backward) ---------------------------
    sys.modules["torch_sparse"].spspmm = None
    sys.modules["torch_sparse"].spmm = None
    from models.base_scalable.simple_models import Layer2GraphConvolution
    from models.utils import scipy_sparse_mat_to_torch_sparse_tensor
    for tag, r in (("gcn", 0.5), ("gcn_r03", 0.3)):
        torch.manual_seed(7)
        a = sym_graph(400, 2400, 6)
        xg = rng.random((400, 20), dtype=np.float32)
        adj_n = Sym(None, r=r).construct_adj(a)
        layer = Layer2GraphConvolution(feat_dim=20, hidden_dim=16, output_dim=5, dropout=0.5)
        layer.eval()
        layer.adj = scipy_sparse_mat_to_torch_sparse_tensor(adj_n)
        xt = torch.from_numpy(xg).requires_grad_(True)
        y = layer(xt)
        target = torch.from_numpy(rng.integers(0, 5, 400))
        loss = torch.nn.functional.cross_entropy(y, target)
        loss.backward()
        out[f"{tag}_x"] = xg
        out[f"{tag}_target"] = target.numpy()
        out[f"{tag}_y"] = y.detach().numpy()
        out[f"{tag}_loss"] = np.array([loss.item()], dtype=np.float32)
        out[f"{tag}_grad_x"] = xt.grad.numpy()
        for k, v in layer.state_dict().items():
            out[f"{tag}_w_{k}"] = v.numpy()
        for k, v in layer.named_parameters():
            if v.grad is not None:
                out[f"{tag}_g_{k}"] = v.grad.numpy()

    # ---- magnetic operators of directed graphs ------------------------------------------------------
    import scipy.sparse as sp
    import operators.utils as U

    def coalesce(index, value, m, n, op="add"):
        key = index[0] * n + index[1]
        uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
        acc = torch.zeros((uniq.numel(),) + tuple(value.shape[1:]), dtype=value.dtype)
        acc.index_add_(0, inv, value)
        return torch.stack([uniq // n, uniq % n]), acc

    def scatter_add(src, index, dim=0, dim_size=None):
        return torch.zeros(dim_size, dtype=src.dtype).index_add_(0, index, src)

    U.coalesce, U.scatter_add = coalesce, scatter_add
    from operators.graph_operator.symmetrical_directed_magnetic_comppr_operator import SymDirMagComPprGraphOp
    from operators.graph_operator.symmetrical_directed_magnetic_laplacian_operator import SymDirMagLaplacianGraphOp
    rg = np.random.default_rng(21)

    def digraph(n, m, weighted=False, loops=False):
        rows, cols = rg.integers(0, n, m), rg.integers(0, n, m)
        if not loops:
            keep = rows != cols
            rows, cols = rows[keep], cols[keep]
        a = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n))
        a.data[:] = 1.0
        if weighted:
            a.data[:] = np.round(rg.random(a.nnz) + 0.25, 3)
        a.sort_indices()
        return a

    mag_cases = {"mag_unw": (digraph(120, 700), 0.5, 0.25, 3), "mag_w": (digraph(90, 500, weighted=True, loops=True), 0.3, 0.1, 2),
                 "mag_tiny": (digraph(7, 12, loops=True), 0.5, 0.25, 4)}
    for tag, (a, r, q, k) in mag_cases.items():
        xm = rg.random((a.shape[0], 6), dtype=np.float32) - 0.5
        csr_pack(f"{tag}_adj", a, out)
        out[f"{tag}_x"] = xm
        out[f"{tag}_params"] = np.array([r, q, k])
        op = SymDirMagLaplacianGraphOp(k, r=r, q=q)
        re, im = op.propagate(a, xm)
        csr_pack(f"{tag}_real", op.real_adj, out)
        csr_pack(f"{tag}_imag", op.imag_adj, out)
        out[f"{tag}_re_hops"] = np.stack([t.numpy() for t in re])
        out[f"{tag}_im_hops"] = np.stack([t.numpy() for t in im])
        op = SymDirMagComPprGraphOp(k, r=r, q=q, ppr_alpha=0.15)
        re, im = op.propagate(a, xm)
        csr_pack(f"{tag}_ppr_real", op.real_adj, out)
        csr_pack(f"{tag}_ppr_imag", op.imag_adj, out)
        out[f"{tag}_ppr_re_hops"] = np.stack([t.numpy() for t in re])
        out[f"{tag}_ppr_im_hops"] = np.stack([t.numpy() for t in im])

    # ---- undirected / in / out operators of a directed graph (TwoDirLaplacianGraphOp) ------------------
    def add_self_loops(edge_index, edge_attr=None, fill_value=1.0, num_nodes=None):
        # torch_geometric.utils.add_self_loops: one (i, i) entry appended per node, existing loops kept
        loop = torch.arange(num_nodes, dtype=torch.long)      # as in PyG: the cat promotes int32 indices to int64
        ei = torch.cat([edge_index, torch.stack([loop, loop])], dim=1)
        ea = torch.cat([edge_attr, torch.full((num_nodes,), fill_value, dtype=edge_attr.dtype)])
        return ei, ea

    U.add_self_loops = add_self_loops
    from operators.graph_operator.in_out_directed_laplacian_operator import TwoDirLaplacianGraphOp
    for tag, (a, r, k) in {"twodir": (digraph(150, 600), 0.5, 2), "twodir_loops": (digraph(60, 260, loops=True), 0.3, 2)}.items():
        xm = rg.random((a.shape[0], 5), dtype=np.float32)
        csr_pack(f"{tag}_adj", a, out)
        out[f"{tag}_x"] = xm
        out[f"{tag}_params"] = np.array([r, k])
        op = TwoDirLaplacianGraphOp(k, r=r)
        lists = op.propagate(a, xm)
        for nm, m, hops in zip(("un", "in", "out"), (op.un_adj, op.in_adj, op.out_adj), lists):
            csr_pack(f"{tag}_{nm}", m, out)
            out[f"{tag}_{nm}_hops"] = np.stack([t.numpy() for t in hops])

    # ---- the two PPR-approximation operators of directed graphs (oracle pinned now, device path next round) -----
    import scipy
    if not hasattr(scipy, "newaxis"):
        scipy.newaxis = None            # utils.py:283 uses the alias of numpy.newaxis that newer scipy removed
    from operators.graph_operator.symmetrical_directed_fast_ppr_approximate_operator import SymDirFastPprApproxGraphOp
    from operators.graph_operator.symmetrical_directed_two_order_ppr_approximate_operator import \
        SymDirTwoOrderPprApproxGraphOp
    a = digraph(60, 300)
    xm = rg.random((60, 4), dtype=np.float32)
    csr_pack("ppr_adj", a, out)
    out["ppr_x"] = xm
    op = SymDirFastPprApproxGraphOp(2, r=0.5, ppr_alpha=0.1)
    hops = op.propagate(a, xm)
    csr_pack("fastppr_norm", op.adj, out)
    out["fastppr_hops"] = np.stack([t.numpy() for t in hops])
    op = SymDirTwoOrderPprApproxGraphOp(2, r=0.5, ppr_alpha=0.1)
    h1, h2 = op.propagate(a, xm)
    csr_pack("twoorder_one", op.one_adj, out)
    csr_pack("twoorder_two", op.two_adj, out)
    out["twoorder_one_hops"] = np.stack([t.numpy() for t in h1])
    out["twoorder_two_hops"] = np.stack([t.numpy() for t in h2])

    np.savez_compressed(os.path.join(HERE, "reference_ext.npz"), **out)
    print("wrote", sorted(out))


if __name__ == "__main__":
    main()
