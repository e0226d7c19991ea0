"""Generate tests/golden/reference_augment.npz from the REAL reference (build container only).

    python tests/golden/make_golden_augment.py

SURVEY.md §8f-4: `edge_augument` (SSRG/data_augument.py:73-103) gives every node below `degree_level` new
neighbours chosen among random candidates by soft-label distance, then symmetrises and de-duplicates the edge list.
data_augument.py cannot be imported (its module-level imports need the absent `datasets`, `rich`, `configs` ...
packages), so the function's OWN SOURCE is taken from the file by ast, compiled and executed unchanged, together with
the two helpers it calls (`generate_numbers`, `compute_distance`, SSRG/utils.py:29-38); only the names it reads
from the module namespace (`data_augument_args`, `Counter`, `torch`) are supplied.  Nothing is restated here.
"""
import ast
import os
import random
import types
from collections import Counter

import numpy as np
import torch

REF = "/root/reference/Scalable Spectral Robust GNN"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_function(path, name, namespace):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), namespace)
            return namespace[name]
    raise KeyError(name)


def reference_edge_augument(degree_level):
    ns = {"torch": torch, "Counter": Counter, "random": random,
          "data_augument_args": types.SimpleNamespace(degree_level=degree_level)}
    load_function(os.path.join(REF, "utils.py"), "generate_numbers", ns)
    load_function(os.path.join(REF, "utils.py"), "compute_distance", ns)
    return load_function(os.path.join(REF, "data_augument.py"), "edge_augument", ns)


def make_case(n, m, classes, degree_level, seed, isolated):
    rng = np.random.default_rng(seed)
    u, v = rng.integers(0, n - isolated, m), rng.integers(0, n - isolated, m)     # the last `isolated` nodes have no edge
    keep = u != v
    u, v = u[keep], v[keep]
    row = torch.from_numpy(np.concatenate([u, v]).astype(np.int64))              # both directions, as dataset.edge holds them
    col = torch.from_numpy(np.concatenate([v, u]).astype(np.int64))
    logits = torch.from_numpy(rng.standard_normal((n, classes)).astype(np.float32) * 2.0)
    soft = torch.nn.functional.softmax(logits, dim=1)
    dataset = types.SimpleNamespace(edge=types.SimpleNamespace(row=row, col=col), x=np.zeros((n, 3), np.float32))
    random.seed(seed)
    out = reference_edge_augument(degree_level)(dataset, soft)
    return {"row": row.numpy(), "col": col.numpy(), "soft": soft.numpy(), "n": n, "degree_level": degree_level,
            "seed": seed, "edge_index": out.numpy()}


def main():
    out = {}
    cases = [("a", make_case(600, 700, 7, 1, 2023, 25)),       # default degree_level = 1: only isolated nodes get an edge
             ("b", make_case(1500, 1400, 10, 3, 7, 40)),       # many low-degree nodes, 100 * deficit candidates each
             ("c", make_case(400, 3000, 5, 2, 11, 0))]         # dense enough that few nodes qualify
    for tag, c in cases:
        for k, v in c.items():
            out[f"{tag}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "reference_augment.npz"), **out)
    for tag, c in cases:
        print(tag, "edges in", len(c["row"]), "edge_index out", c["edge_index"].shape)


if __name__ == "__main__":
    main()
