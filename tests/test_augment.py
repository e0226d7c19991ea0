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


@pytest.mark.parametrize("tag", CASES)
def test_host_side_of_the_mirror_vs_reference_golden(tag):
    """The host logic of augment.edge_augument without a GPU: visiting order (low_degree_order) and the reference's
    candidate stream (draw_candidates) fed with numpy stand-ins for the three device stages reproduce the reference's
    edge_index, i.e. what stays on the host is exactly the reference's and the device stages are plain data-parallel
    work (counts, distances + top-k, symmetrise + unique)."""
    from scalable_roubust_gnn_b200 import augment
    c = _case(tag)
    n, level = int(c["n"]), int(c["degree_level"])
    both = np.concatenate([c["row"], c["col"]])
    counts = np.bincount(both, minlength=n).astype(np.int32)
    first = np.full(n, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(first, both, np.arange(len(both), dtype=np.int64))
    nodes, cnts = augment.low_degree_order(torch.from_numpy(counts), torch.from_numpy(first), level)
    random.seed(int(c["seed"]))
    cands = augment.draw_candidates(nodes, cnts, n, level)
    src, dst = [c["row"]], [c["col"]]
    for node, cnt, cand in zip(nodes, cnts, cands):
        diff = c["soft"][node][None, :] - c["soft"][cand]
        dist = np.sqrt((diff.astype(np.float64) ** 2).sum(1)).astype(np.float32)
        order = np.argsort(dist, kind="stable")[: level - cnt]
        src.append(np.full(level - cnt, node, dtype=np.int64))
        dst.append(cand[order].astype(np.int64))
    r, cc = np.concatenate(src), np.concatenate(dst)
    got = np.unique(np.stack([np.concatenate([r, cc]), np.concatenate([cc, r])]), axis=1)
    np.testing.assert_array_equal(got, c["edge_index"])
