"""Where does the exposed part of the per-hop exchange come from?  (launched with torch.distributed.run, one rank per GPU)

Products-shaped graph, P x 1 row partition.  Every experiment is one hop's worth of work per rank, timed with CUDA
events on the rank's stream after a barrier; per-rank times are gathered so one-sided experiments show which SIDE
(sender or receiver) pays.

  A  local hop only (rows written to the local buffer)
  B  fused push hop (per-lane remote stores in the epilogue)             = the product path
  C  fused push hop, bulk-store (TMA) epilogue
  D  local hop  ||  push_rows kernel of an independent buffer on a second stream (SM-driven exchange, not fused)
  E  local hop  ||  cudaMemcpyAsync of the same rows to the peers on a second stream (copy engines)
  F  like D but ONLY rank 0 pushes (rank 0 = sender + compute, others = receiver + compute)
  G  like E but ONLY rank 0 copies
  X  exchange alone (push_rows, all ranks)          Y  exchange alone (copy engines, all ranks)

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/overlap_lab.py [--scale 1.0]
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from scalable_roubust_gnn_b200 import _lib, device as dev, dist as sdist, synth
from scalable_roubust_gnn_b200.device import _p, _stream_ptr

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--feat-groups", type=int, default=1)
args = ap.parse_args()

world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
local_rank = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
lib = _lib.load()
n, nnz, f, k = synth.SHAPES["products"]
n, nnz = int(n * args.scale), int(nnz * args.scale)
st = sdist.DistState(n, f, world, rank, mode="push", feat_groups=args.feat_groups)
a = synth.uniform_graph(n, nnz)
a_loc = dev.upload_csr(sdist.shard_rows(a, st.row0, st.row0 + st.n_local))
x_loc = dev.pack_features(torch.from_numpy(np.ascontiguousarray(synth.features(n, f)[st.row0:st.row0 + st.n_local, st.f0:st.f1])).cuda())
del a
sdist.start_input_exchange(st, x_loc)
norm, flags = sdist.dist_sym_norm(st, a_loc, 0.5)
sdist.propagate_device(st, norm, x_loc, 1, keep_hops=False)
torch.cuda.synchronize()
dist.barrier()

s_main = torch.cuda.current_stream()
side = torch.cuda.Stream()
xin = st.full[0]
out_local = st.full[1][st.row0:st.row0 + st.n_local]
dests1 = (C.c_void_p * len(st.peers))(*st.peer_ptrs[1])
dests0 = (C.c_void_p * len(st.peers))(*st.peer_ptrs[0])
row_bytes = st.ld * 4


def local_hop():
    _lib.check(lib.srg_spmm_csr_f32(_p(norm.indptr), _p(norm.indices), _p(norm.data), st.n_local, norm.nnz_bound,
                                    _p(xin), st.ld, _p(out_local), st.ld, st.f_loc, _stream_ptr()))


def fused(tma):
    def fn():
        _lib.set_tuning("push_tma", tma)
        _lib.check(lib.srg_spmm_csr_f32_push(_p(norm.indptr), _p(norm.indices), _p(norm.data), st.n_local, norm.nnz_bound,
                                             _p(xin), st.ld, dests1, len(st.peers), st.row0, st.ld, st.f_loc, _stream_ptr()))
        _lib.set_tuning("push_tma", 0)
    return fn


def push_side(only_rank0):
    def fn():
        if only_rank0 and rank != 0:
            return
        side.wait_stream(s_main)
        src = xin[st.row0:st.row0 + st.n_local]
        _lib.check(lib.srg_push_rows_f32(_p(src), st.n_local, st.ld, dests0, len(st.peers), st.row0, C.c_void_p(side.cuda_stream)))
    return fn


def copy_side(only_rank0):
    def fn():
        if only_rank0 and rank != 0:
            return
        side.wait_stream(s_main)
        src = xin.data_ptr() + st.row0 * row_bytes
        for blk, peer in enumerate(st.peers):
            if peer != rank:
                _lib.check(lib.srg_copy_async(C.c_void_p(st.peer_ptrs[0][blk] + st.row0 * row_bytes), C.c_void_p(src),
                                              st.n_local * row_bytes, C.c_void_p(side.cuda_stream)))
    return fn


def run(name, main_fn, side_fn=None):
    times = []
    for rep in range(args.reps + 1):
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if side_fn is not None:
            side_fn()
        if main_fn is not None:
            main_fn()
        if side_fn is not None:
            s_main.wait_stream(side)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            times.append(e0.elapsed_time(e1))
    t = torch.tensor([float(np.mean(times))], device="cuda")
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        print(json.dumps({"exp": name, "ms_per_rank": [round(float(v.item()), 4) for v in allt]}), flush=True)


if rank == 0:
    print(json.dumps({"world": world, "grid": f"{st.n_row_blocks}x{st.feat_groups}", "rows_per_rank": st.n_local, "ld": st.ld,
                      "exchange_MB_out_per_rank": round((len(st.peers) - 1) * st.n_local * row_bytes / 1e6, 1)}), flush=True)
run("A local hop only", local_hop)
run("B fused push (remote st.global epilogue)", fused(0))
run("C fused push (TMA bulk-store epilogue)", fused(1))
run("D local hop || push_rows kernel (2nd stream)", local_hop, push_side(False))
run("E local hop || cudaMemcpyAsync to peers (copy engines)", local_hop, copy_side(False))
run("F local hop || push_rows, ONLY rank 0 sends", local_hop, push_side(True))
run("G local hop || copy engine, ONLY rank 0 sends", local_hop, copy_side(True))
run("X exchange alone: push_rows, all ranks", None, push_side(False))
run("Y exchange alone: copy engines, all ranks", None, copy_side(False))
run("A2 local hop only (again)", local_hop)
torch.cuda.synchronize()
dist.barrier()
st.close()
dist.destroy_process_group()
