"""CPU suite for the N>1 path: partition map, row sharding and the per-hop exchange orchestration,
driven over gloo with world_size 2 (the local hop is the oracle; the GPU ops are covered by
tests/test_dist_gpu.py on a multi-GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from helpers import sym_graph
from scalable_roubust_gnn_b200 import dist as sdist


def test_row_partition_matches_oracle_definition():
    for n, w in [(10, 4), (8, 8), (3, 8), (2449029, 8), (111059956, 8), (1, 1), (0, 2)]:
        rp, st = sdist.row_partition(n, w)
        rp_o, st_o = oracle.row_partition(n, w) if n > 0 else (0, np.zeros(w + 1, dtype=np.int64))
        assert rp == rp_o
        np.testing.assert_array_equal(st, st_o)
        assert st[0] == 0 and st[-1] == n and (np.diff(st) >= 0).all()
        # owner-of-row and local offset are what the kernels assume: global = rank * rows_per + local
        if n > 0:
            rows = np.array([0, n // 2, n - 1])
            owner = rows // rp
            assert ((st[owner] <= rows) & (rows < st[owner + 1])).all()


def test_shard_rows_concatenate_to_the_whole():
    a = oracle.sym_norm(sym_graph(1000, 8000, 3), 0.5)
    rp, st = sdist.row_partition(1000, 3)
    parts = [sdist.shard_rows(a, int(st[r]), int(st[r + 1])) for r in range(3)]
    assert sum(p.nnz for p in parts) == a.nnz
    import scipy.sparse as sp
    assert (sp.vstack(parts) != a).nnz == 0
    assert all(p.shape[1] == 1000 and p.indptr[0] == 0 for p in parts)


class _GlooOps:
    """ops for propagate_sharded: numpy buffers, oracle hop, gloo all-gather."""

    def __init__(self, n, f, rows_per, world, rank, row0, n_local):
        self.n, self.f, self.rows_per, self.world, self.rank, self.row0, self.n_local = n, f, rows_per, world, rank, row0, n_local

    def new_full(self):
        return np.zeros((self.rows_per * self.world, self.f), dtype=np.float32)

    def load_local(self, full, x_local):
        full[self.row0:self.row0 + self.n_local] = x_local

    def exchange(self, full):
        mine = torch.from_numpy(full[self.rank * self.rows_per:(self.rank + 1) * self.rows_per].copy())
        out = torch.from_numpy(full)
        dist.all_gather_into_tensor(out, mine)

    def hop(self, local_norm, full_in, full_out):
        full_out[self.row0:self.row0 + self.n_local] = oracle.spmm_hop(local_norm, full_in)

    def snapshot_local(self, full):
        return full[self.row0:self.row0 + self.n_local].copy()


def _worker(rank, world, port, n, f, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        adj = sym_graph(n, 8 * n, 5)
        x = np.random.default_rng(1).random((n, f), dtype=np.float32)
        norm = oracle.sym_norm(adj, 0.5)
        rows_per, starts = sdist.row_partition(n, world)
        s, e = int(starts[rank]), int(starts[rank + 1])
        local = sdist.shard_rows(norm, s, e)
        # the local operator multiplies the padded full buffer: pad the column space
        import scipy.sparse as sp
        local = sp.csr_matrix((local.data, local.indices, local.indptr), shape=(e - s, rows_per * world))
        ops = _GlooOps(n, f, rows_per, world, rank, s, e - s)
        hops = sdist.propagate_sharded(ops, local, x[s:e], k, rows_per, world)
        np.save(os.path.join(out_dir, f"hops_{rank}.npy"), np.stack(hops))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 1001])
def test_sharded_propagation_bitwise_equals_single_process(tmp_path, n):
    world, f, k = 2, 12, 3
    port = 29500 + (os.getpid() % 500) + n % 7
    mp.spawn(_worker, args=(world, port, n, f, k, str(tmp_path)), nprocs=world, join=True)
    adj = sym_graph(n, 8 * n, 5)
    x = np.random.default_rng(1).random((n, f), dtype=np.float32)
    want, _ = oracle.propagate(adj, x, k)
    got = np.concatenate([np.load(tmp_path / f"hops_{r}.npy") for r in range(world)], axis=1)
    for h in range(k + 1):
        np.testing.assert_array_equal(got[h], want[h])


def test_grid_coordinates_and_feature_slices():
    """P_r x P_f grid used when the exchange is NVLink-bound: rank = ri * P_f + ci."""
    assert [sdist.grid_coords(r, 8, 2) for r in range(8)] == [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1), (3, 0), (3, 1)]
    assert sdist.push_peers(5, 8, 2) == [1, 3, 5, 7] and sdist.push_peers(2, 4, 1) == [0, 1, 2, 3]
    assert [sdist.feature_slice(100, 2, c) for c in range(2)] == [(0, 50), (50, 100)]
    assert [sdist.feature_slice(7, 3, c) for c in range(3)] == [(0, 3), (3, 6), (6, 7)]
    assert sdist.feature_slice(2, 4, 3) == (2, 2)
    cover = sorted(c for ci in range(4) for c in range(*sdist.feature_slice(129, 4, ci)))
    assert cover == list(range(129))
    with pytest.raises(ValueError):
        sdist.grid_coords(0, 6, 4)


def _grid_worker(rank, world, port, n, f, k, pf, out_dir):
    """2 x 2 grid (row blocks x feature slices): every rank propagates ITS feature slice over ITS row block and
    exchanges rows only with the ranks that hold the same slice (push_peers) - the scheme of the 8-GPU bench."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.sparse as sp
        groups = {}
        for ci in range(pf):                       # new_group is collective: every rank creates every group
            members = sdist.push_peers(ci, world, pf)
            groups[ci] = (members, dist.new_group(members))
        ri, ci = sdist.grid_coords(rank, world, pf)
        members, group = groups[ci]
        assert members == sdist.push_peers(rank, world, pf) and members[ri] == rank
        n_blocks = world // pf
        adj = sym_graph(n, 8 * n, 5)
        x = np.random.default_rng(1).random((n, f), dtype=np.float32)
        norm = oracle.sym_norm(adj, 0.5)
        rows_per, starts = sdist.row_partition(n, n_blocks)
        s, e = int(starts[ri]), int(starts[ri + 1])
        f0, f1 = sdist.feature_slice(f, pf, ci)
        local = sdist.shard_rows(norm, s, e)
        local = sp.csr_matrix((local.data, local.indices, local.indptr), shape=(e - s, rows_per * n_blocks))

        class _GroupOps(_GlooOps):
            def exchange(self, full):
                mine = torch.from_numpy(full[self.rank * self.rows_per:(self.rank + 1) * self.rows_per].copy())
                dist.all_gather_into_tensor(torch.from_numpy(full), mine, group=group)

        ops = _GroupOps(n, f1 - f0, rows_per, n_blocks, ri, s, e - s)      # "rank" inside the group = row block index
        hops = sdist.propagate_sharded(ops, local, np.ascontiguousarray(x[s:e, f0:f1]), k, rows_per, n_blocks)
        np.save(os.path.join(out_dir, f"grid_{rank}.npy"), np.stack(hops))
    finally:
        dist.destroy_process_group()


def test_row_by_feature_grid_bitwise_equals_single_process(tmp_path):
    world, pf, n, f, k = 4, 2, 1003, 13, 3
    port = 29400 + (os.getpid() % 500)
    mp.spawn(_grid_worker, args=(world, port, n, f, k, pf, str(tmp_path)), nprocs=world, join=True)
    adj = sym_graph(n, 8 * n, 5)
    x = np.random.default_rng(1).random((n, f), dtype=np.float32)
    want, _ = oracle.propagate(adj, x, k)
    rows_per, starts = sdist.row_partition(n, world // pf)
    for rank in range(world):
        ri, ci = sdist.grid_coords(rank, world, pf)
        f0, f1 = sdist.feature_slice(f, pf, ci)
        got = np.load(tmp_path / f"grid_{rank}.npy")
        for h in range(k + 1):
            # every output element is one FMA chain over the row's entries: slicing columns or rows changes no bit
            np.testing.assert_array_equal(got[h], want[h][int(starts[ri]):int(starts[ri + 1]), f0:f1])
