"""Generate tests/golden/reference_ext.npz from the REAL reference (build container only).

    python tests/golden/make_golden_ext.py

Second batch of fixtures (the "next" rows of SURVEY.md §8f), same recipe as make_golden.py: the
reference's own classes imported from /root/reference, run on small inputs, outputs recorded.
  * nafs_*   OverSmoothDistanceWeightedOp.aggregate on a reference hop list
             (SSRG/operators/message_operator/over_smooth_distance_op.py)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, sym_graph  # noqa: E402


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

    np.savez_compressed(os.path.join(HERE, "reference_ext.npz"), **out)
    print("wrote", sorted(out))


if __name__ == "__main__":
    main()
