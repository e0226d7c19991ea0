"""Counter-based synthetic generators: the numpy definition, and the device generator
against it bit for bit (integer work)."""
import numpy as np
import pytest
import torch

from scalable_roubust_gnn_b200 import synth


def test_scramble_is_a_bijection_and_graph_is_symmetric_loop_free():
    for scale in (5, 12, 13):
        s = synth._scramble(np.arange(1 << scale, dtype=np.uint64), scale, 7)
        assert len(np.unique(s)) == 1 << scale and int(s.max()) < 1 << scale
    n = 3000
    a = synth.rmat_scrambled_host(n, synth.rmat_draws(n, 40000), seed=5)
    assert (a != a.T).nnz == 0 and a.diagonal().sum() == 0 and a.has_sorted_indices
    assert np.all(a.data == 1.0)
    deg = np.diff(a.indptr)
    assert deg.max() > 20 * max(1.0, deg.mean())            # power law: hubs exist
    blocks = [a.indptr[(i + 1) * n // 4] - a.indptr[i * n // 4] for i in range(4)]
    assert max(blocks) < 1.6 * min(blocks)                   # scrambling spreads them over the row blocks


def test_hash_features_host_layout_independent():
    full = synth.hash_features_host(1, 0, 50, 0, 16, 16)
    part = synth.hash_features_host(1, 10, 7, 4, 8, 16)
    np.testing.assert_array_equal(part, full[10:17, 4:12])
    rows = synth.hash_features_host(1, 0, 0, 4, 8, 16, rows=[3, 40, 3])
    np.testing.assert_array_equal(rows, full[[3, 40, 3], 4:12])
    assert full.dtype == np.float32 and full.min() >= 0.0 and full.max() < 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("n,nnz", [(5000, 60000), (1 << 12, 30000), (777, 9000)])
def test_device_rmat_shards_equal_host_graph(n, nnz):
    m_draw = synth.rmat_draws(n, nnz)
    a = synth.rmat_scrambled_host(n, m_draw, seed=3)
    bounds = [0, n // 3, n // 3, (2 * n) // 3 + 5, n]          # includes an empty shard
    total = 0
    for r0, r1 in zip(bounds[:-1], bounds[1:]):
        d = synth.rmat_shard_device(n, m_draw, r0, r1, seed=3)
        want_ptr = a.indptr[r0:r1 + 1] - a.indptr[r0]
        np.testing.assert_array_equal(d.indptr.cpu().numpy(), want_ptr)
        assert d.nnz == int(want_ptr[-1]) and d.data is None
        np.testing.assert_array_equal(d.indices[:d.nnz].cpu().numpy(), a.indices[a.indptr[r0]:a.indptr[r1]])
        total += d.nnz
    assert total == a.nnz


@pytest.mark.gpu
def test_device_rmat_capacity_overflow_is_reported():
    from scalable_roubust_gnn_b200 import SrgError
    with pytest.raises(SrgError, match="exceed the capacity"):
        synth.rmat_shard_device(5000, 40000, 0, 5000, seed=3, cap=100)


@pytest.mark.gpu
def test_device_hash_features_equal_host():
    got = synth.hash_features_device(300, 21, row0=1000, col0=5, f_total=64, seed=9)
    want = synth.hash_features_host(9, 1000, 300, 5, 21, 64)
    assert got.shape == (300, 24)
    np.testing.assert_array_equal(got[:, :21].cpu().numpy(), want)
    assert float(got[:, 21:].abs().sum()) == 0.0


@pytest.mark.gpu
def test_propagation_on_device_generated_graph_vs_oracle():
    """The generated shard feeds the normalisation + hops like any CSR (all-ones values = NULL)."""
    import oracle
    from scalable_roubust_gnn_b200 import device as dev
    n, f = 4000, 20
    m_draw = synth.rmat_draws(n, 50000)
    a_host = synth.rmat_scrambled_host(n, m_draw, seed=1)
    a_dev = synth.rmat_shard_device(n, m_draw, 0, n, seed=1)
    x = synth.hash_features_device(n, f, seed=1)
    norm, flags, _ = dev.sym_norm(a_dev, 0.5)
    hops = dev.propagate(norm, x, f, 2)
    assert int(flags.item()) & ~16 == 0
    want, _ = oracle.propagate(a_host, synth.hash_features_host(1, 0, n, 0, f, f), 2, r=0.5)
    np.testing.assert_array_equal(hops[2][:, :f].cpu().numpy(), want[2])
