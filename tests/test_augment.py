"""edge_augument (SSRG/data_augument.py:73-103): oracle and device mirror against outputs of the
reference's own function (tests/golden/reference_augment.npz, made by tests/golden/make_golden_augment.py)."""
import os
import random
import types

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN

CASES = ["a", "b", "c"]


def _case(tag):
    g = np.load(os.path.join(GOLDEN, "reference_augment.npz"))
    return {k[len(tag) + 1:]: g[k] for k in g.files if k.startswith(tag + "_")}


@pytest.mark.parametrize("tag", CASES)
def test_oracle_edge_augument_vs_reference_golden(tag):
    c = _case(tag)
    got = oracle.edge_augument(c["row"], c["col"], int(c["n"]), c["soft"], int(c["degree_level"]), seed=int(c["seed"]))
    np.testing.assert_array_equal(got, c["edge_index"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_device_edge_augument_vs_reference_golden(tag):
    """Bit-exact edge_index: endpoint counts, visiting order, the reference's candidate stream, device distances /
    top-k, symmetrise + unique."""
    from scalable_roubust_gnn_b200 import augment
    c = _case(tag)
    n = int(c["n"])
    ds = types.SimpleNamespace(edge=types.SimpleNamespace(row=torch.from_numpy(c["row"]), col=torch.from_numpy(c["col"])),
                               x=np.zeros((n, 1), np.float32))
    random.seed(int(c["seed"]))
    got = augment.edge_augument(ds, torch.from_numpy(c["soft"]), degree_level=int(c["degree_level"]))
    assert got.dtype == torch.int64 and not got.is_cuda
    np.testing.assert_array_equal(got.numpy(), c["edge_index"])


@pytest.mark.gpu
def test_endpoint_counts_and_visiting_order():
    from collections import Counter

    from scalable_roubust_gnn_b200 import augment
    rng = np.random.default_rng(3)
    n, m = 5000, 6000
    row, col = rng.integers(0, n - 50, m), rng.integers(0, n - 50, m)
    counts, first = augment.endpoint_counts(torch.from_numpy(row).cuda(), torch.from_numpy(col).cuda(), n)
    ref = Counter(np.concatenate([row, col]).tolist())
    want = np.array([ref.get(i, 0) for i in range(n)])
    np.testing.assert_array_equal(counts.cpu().numpy(), want)
    for i in range(n):
        if i not in ref:
            ref.update({i: 0})
    order = [k for k, v in sorted(ref.items(), key=lambda kv: kv[1]) if v < 3]
    nodes, cnts = augment.low_degree_order(counts, first, 3)
    assert nodes.tolist() == order
    np.testing.assert_array_equal(cnts, want[nodes])
    with pytest.raises(IndexError):
        augment.endpoint_counts(torch.tensor([0, n]).cuda(), torch.tensor([1, 2]).cuda(), n)
