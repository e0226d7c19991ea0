"""Native multi-GPU handle (csrc/dist.cu): NCCL bound at run time, row partition, one all-gather per
hop, against the single-GPU path bit for bit."""
import os
import time

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from helpers import sym_graph


def test_unique_id_and_no_device_error():
    from scalable_roubust_gnn_b200 import SrgError, _lib, dist as sdist
    a, b = sdist.native_unique_id(), sdist.native_unique_id()
    assert len(a) == 128 and a != b and any(a)
    if not torch.cuda.is_available():
        with pytest.raises(SrgError) as e:
            sdist.NativeDist(a, 1, 0, 10, 4)
        assert e.value.code == _lib.SRG_ERR_NODEV


def _single_gpu(adj, x, k, r, alpha=None):
    from scalable_roubust_gnn_b200 import device as dev
    norm, flags, _ = dev.sym_norm(dev.upload_csr(adj), r, alpha)
    assert int(flags.item()) & ~16 == 0
    f = x.shape[1]
    hops = dev.propagate(norm, dev.pack_features(torch.from_numpy(x).cuda()), f, k)
    return np.stack([h[:, :f].cpu().numpy() for h in hops])


@pytest.mark.gpu
@pytest.mark.parametrize("weighted,f,alpha", [(False, 100, None), (True, 37, 0.15)])
def test_native_handle_world_1_equals_device_path(weighted, f, alpha):
    from scalable_roubust_gnn_b200 import device as dev, dist as sdist
    n, k = 20001, 3
    adj = sym_graph(n, 10 * n, 5, weighted=weighted)
    x = np.random.default_rng(1).random((n, f), dtype=np.float32)
    nd = sdist.NativeDist(sdist.native_unique_id(), 1, 0, n, f)
    assert (nd.row0, nd.n_local, nd.rows_per) == (0, n, n) and nd.ld % 8 == 0
    hops, flags = nd.propagate(dev.upload_csr(adj), torch.from_numpy(x).cuda(), k, 0.5, alpha)
    torch.cuda.synchronize()
    assert int(flags.item()) & ~16 == 0
    np.testing.assert_array_equal(np.stack([h.cpu().numpy() for h in hops]), _single_gpu(adj, x, k, 0.5, alpha))
    # the handle is reusable
    hops2, _ = nd.propagate(dev.upload_csr(adj), torch.from_numpy(x).cuda(), 1, 0.5, alpha)
    np.testing.assert_array_equal(hops2[1].cpu().numpy(), hops[1].cpu().numpy())
    nd.close()


def _worker(rank, world, n, f, k, out_dir):
    torch.cuda.set_device(rank)
    from scalable_roubust_gnn_b200 import device as dev, dist as sdist
    id_file = os.path.join(out_dir, "nccl_id.bin")
    if rank == 0:
        with open(id_file + ".tmp", "wb") as fh:
            fh.write(sdist.native_unique_id())
        os.replace(id_file + ".tmp", id_file)
    while not os.path.exists(id_file):
        time.sleep(0.05)
    uid = open(id_file, "rb").read()
    adj = sym_graph(n, 10 * n, 5)
    x = np.random.default_rng(1).random((n, f), dtype=np.float32)
    nd = sdist.NativeDist(uid, world, rank, n, f)
    s, e = nd.row0, nd.row0 + nd.n_local
    hops, flags = nd.propagate(dev.upload_csr(sdist.shard_rows(adj, s, e)), torch.from_numpy(x[s:e]).cuda(), k)
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), hops=np.stack([h.cpu().numpy() for h in hops]), flags=int(flags.item()))
    nd.close()


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_native_handle_two_gpus_bitwise_equal_one_gpu(tmp_path):
    world, n, f, k = 2, 50001, 100, 3
    mp.spawn(_worker, args=(world, n, f, k, str(tmp_path)), nprocs=world, join=True)
    adj = sym_graph(n, 10 * n, 5)
    x = np.random.default_rng(1).random((n, f), dtype=np.float32)
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    assert all(int(p["flags"]) & ~16 == 0 for p in parts)
    np.testing.assert_array_equal(np.concatenate([p["hops"] for p in parts], axis=1), _single_gpu(adj, x, k, 0.5))
