"""bench.py body for N > 1 GPUs (one rank per GPU, launched by torch.distributed.run).

Strong scaling: the same products-shaped graph is row-partitioned over the ranks; a step is the
sharded normalisation (one degree all-gather) + K hops with the per-hop exchange fused into the
SpMM epilogue (push over NVLink peer mappings) or, with SRG_DIST_MODE=allgather, NCCL all-gather.
Timed on the device (CUDA events) between barriers, max over ranks.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np


def _bind_to_gpu_numa_node(torch, local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs), so the pinned host buffers of the
    end-to-end leg are allocated next to the GPU's PCIe root and 8 ranks do not meet on one socket's memory.
    Best effort: any missing piece leaves the affinity alone.  SRG_NUMA_BIND=0 disables it."""
    if os.environ.get("SRG_NUMA_BIND", "1") == "0":
        return None
    try:
        p = torch.cuda.get_device_properties(local_rank)
        dev = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{dev}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node, len(cpus)
    except Exception:
        return None


def default_feat_groups(world, mode):
    """Row blocks x feature slices of the default grid.  A pure row partition makes every rank RECEIVE (P-1)/P of X
    per hop; at 8 GPUs that exchange (1.14 ms) is twice the local hop, so the features are split in two (the exchange
    volume halves, the gathered rows become 224 bytes wide - the shape the bulk-gather kernel is built for)."""
    if mode not in ("push", "copy", "push_tma"):
        return "1"
    # measured (products shape, ms per step): 4 GPUs 4 x 1 5.52 vs 2 x 2 6.12; 8 GPUs 4 x 2 3.6 vs 8 x 1 6.8 (round 1)
    return "2" if world >= 8 else "1"


def run(args, workloads, metric, unit, emit):
    import torch
    import torch.distributed as dist

    from . import _lib, device as dev, dist as sdist, synth
    lib = _lib.load()

    if os.environ.get("SRG_DEBUG_HANG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["SRG_DEBUG_HANG"]), exit=True)
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    numa = _bind_to_gpu_numa_node(torch, local_rank)
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    mode = os.environ.get("SRG_DIST_MODE", "push")
    n, nnz, f, k = workloads[args.workload]
    n, nnz = int(n * args.scale), int(nnz * args.scale)

    t0 = time.perf_counter()
    # 8 GPUs are NVLink-ingress bound with a pure row partition (every rank receives 7/8 of X per hop):
    # a 4 x 2 grid (row blocks x feature slices) halves the exchange volume
    pf = int(os.environ.get("SRG_FEAT_GROUPS", default_feat_groups(world, mode)))
    st = sdist.DistState(n, f, world, rank, mode=mode, feat_groups=pf)
    s, e = st.row0, st.row0 + st.n_local
    f_loc = st.f_loc
    gen = os.environ.get("SRG_GEN", "device" if args.workload.startswith("papers100M") else "host")
    a_loc_host = x_loc_host = ag_inputs = p1_inputs = None
    if gen == "device":
        # config 5: every rank builds ITS rows of the scrambled R-MAT graph on its GPU (counter-based
        # generator, csrc/coo.cu) and its slice of the procedural features; nothing of this size ever
        # exists on the host
        m_draw = synth.rmat_draws(n, nnz)
        a_loc = synth.rmat_shard_device(n, m_draw, s, e)
        x_loc = synth.hash_features_device(st.n_local, f_loc, row0=s, col0=st.f0, f_total=f)
        cnt = torch.tensor([a_loc.nnz if st.ci == 0 else 0], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt)
        nnz_hat = int(cnt.item()) + n
        if st.ci == 0:
            # column 0 := sqrt(degree): an eigenvector of D^-1/2 (A+I) D^-1/2 with eigenvalue 1, so every
            # hop must hand it back unchanged (size-independent check of the hops and of the exchange)
            deg = (a_loc.indptr[1:] - a_loc.indptr[:-1] + 1).to(torch.float32)
            x_loc[:, 0] = torch.sqrt(deg)
    else:
        gen_fn = synth.rmat_graph if args.workload == "products-rmat" else synth.uniform_graph
        a = gen_fn(n, nnz)                     # every rank regenerates the same graph (fixed seed) ...
        a_loc_host = sdist.shard_rows(a, s, e)  # ... and keeps its row slice
        x_full = synth.features(n, f)
        x_loc_host = np.ascontiguousarray(x_full[s:e, st.f0:st.f1])
        nnz_hat = a.nnz + n
        # the plain row partition + NCCL all-gather (the north-star collective) is timed beside the push grid
        ag_rows_per, ag_starts = sdist.row_partition(n, world)
        ag_s, ag_e = int(ag_starts[rank]), int(ag_starts[rank + 1])
        ag_inputs = (sdist.shard_rows(a, ag_s, ag_e), np.ascontiguousarray(x_full[ag_s:ag_e])) if mode != "allgather" else None
        p1_inputs = (a, x_full) if rank == 0 else None      # rank 0 re-runs the whole graph on ONE GPU for the bitwise check
        del a, x_full
        a_loc = dev.upload_csr(a_loc_host)
        x_loc = dev.pack_features(torch.from_numpy(np.ascontiguousarray(x_loc_host)).cuda())
    torch.cuda.synchronize()
    if rank == 0:
        print(f"[bench] world={world} mode={mode} grid={st.n_row_blocks}x{pf} N={n} nnz_hat={nnz_hat} F={f} K={k} rows/rank={st.rows_per} "
              f"gen={gen} numa={numa} setup {time.perf_counter() - t0:.1f}s", file=sys.stderr, flush=True)

    def step():
        sdist.start_input_exchange(st, x_loc)          # overlaps the normalisation
        norm, flags = sdist.dist_sym_norm(st, a_loc, 0.5)
        # keep_hops=True: the step produces what GraphOp.propagate returns - all K+1 matrices (the rank's rows)
        hops = sdist.propagate_device(st, norm, x_loc, k, keep_hops=True)
        return flags, hops

    from bench import ClockSampler  # noqa: E402  (bench.py is the entry script)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                        # before the warm-up: its start-up must not hit timed steps
    # the ranks finish their host-side set-up (graph generation, uploads) seconds apart; the flag fence of the first
    # step waits for a peer for 2 s at most, so line the ranks up before the first step
    dist.barrier()
    for _ in range(args.warmup):
        flags, _ = step()
    torch.cuda.synchronize()
    assert int(flags.item()) & ~_lib.SRG_FLAG_WEIGHTED == 0, f"normalisation flags {int(flags.item())}"
    step()
    torch.cuda.synchronize()
    st.check_fence()
    sampler.lines.clear()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    t_step = float(ms.item()) * 1e-3 / args.steps
    launches = _lib.launch_count() - launches0

    # ---- where the step time goes: one more step with an event after every stage (rank 0's clock) ----------------
    # (steady state: two unmarked steps are queued right before it, so the host runs ahead of the device as it does
    # in the timed loop and the figures are device time, not launch latency)
    torch.cuda.synchronize()
    dist.barrier()
    step()
    step()
    marks = []
    ev_s = torch.cuda.Event(enable_timing=True)
    ev_s.record()
    sdist.start_input_exchange(st, x_loc)
    norm_b, _ = sdist.dist_sym_norm(st, a_loc, 0.5, marks=marks)
    sdist.propagate_device(st, norm_b, x_loc, k, keep_hops=True, marks=marks)
    torch.cuda.synchronize()
    stages, prev = [], ev_s
    for label, ev in marks:
        stages.append([label, round(prev.elapsed_time(ev), 4)])
        prev = ev
    del norm_b

    # ---- overlap probe (SURVEY 8d multi-GPU reporting): the local hop alone, the exchange alone, the fused hop ---
    # collectives (normalisation all-gather, barrier, the hops' ticks) run unconditionally on every rank; only the
    # purely local timing sits inside the try, so a failure cannot leave another rank waiting
    sdist.start_input_exchange(st, x_loc)
    norm_p, _ = sdist.dist_sym_norm(st, a_loc, 0.5)
    sdist.propagate_device(st, norm_p, x_loc, 1, keep_hops=False)
    torch.cuda.synchronize()
    dist.barrier()
    # the fused hop (kernel with the push epilogue + the fence that orders it across ranks), back to back on
    # the two buffers - the input exchange and the normalisation are not part of this figure
    ops_p = sdist.DeviceOps(st)
    pushed = st.mode in ("push", "copy") and world > 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hop_reps = 6
    ops_p.hop(norm_p, 0, 1)
    ops_p.exchange(1, pushed=pushed)
    e0.record()
    for i in range(hop_reps):
        ops_p.hop(norm_p, i & 1 ^ 1, i & 1)
        ops_p.exchange(i & 1, pushed=pushed)
    e1.record()
    torch.cuda.synchronize()
    t_hops = torch.tensor([e0.elapsed_time(e1) / hop_reps], device="cuda")
    dist.all_reduce(t_hops, op=dist.ReduceOp.MAX)
    probe = {"hop_fused_ms": float(t_hops.item())}
    dist.barrier()
    try:
        probe.update(_overlap_probe(st, norm_p, lib, torch))
        probe["overlap_efficiency"] = max(probe["hop_local_only_ms"], probe["exchange_only_ms"]) / probe["hop_fused_ms"]
    except Exception as exc:                      # auxiliary figures must not take the bench line down
        probe["error"] = str(exc)
    torch.cuda.synchronize()
    dist.barrier()
    del norm_p

    verify = None
    allgather = None
    if gen == "device":
        verify = _verify_device_generated(st, a_loc, x_loc, k, f, sdist, synth, torch, dist)
        e2e_line = None
    else:
        verify = _verify_against_one_gpu(st, a_loc, x_loc, k, f, p1_inputs, dev, sdist, torch, dist)
        p1_inputs = None
        if ag_inputs is not None and os.environ.get("SRG_BENCH_ALLGATHER", "1") != "0":
            allgather = _allgather_baseline(args, n, f, k, world, rank, ag_inputs, dev, sdist, torch, dist)
        ag_inputs = None
        e2e_line = None
        if os.environ.get("SRG_BENCH_E2E", "1") != "0":
            e2e_line = _e2e(args, st, a_loc_host, x_loc_host, k, f_loc, lib, dev, sdist, torch, dist)
    st.check_fence()
    clocks = sampler.stop() if rank == 0 else None
    probe["stages_ms_rank0"] = stages
    if allgather is not None:
        probe["allgather_mode"] = allgather
    _emit_line(args, st, metric, unit, emit, mode, pf, world, rank, n, nnz_hat, f, f_loc, k, t_step, launches, e2e_line,
               clocks, verify, gen, probe)
    st.close()
    dist.destroy_process_group()


def _overlap_probe(st, norm, lib, torch):
    """This rank's hop WITHOUT the exchange (rows written to the local buffer only) and the exchange WITHOUT the
    hop (the finished slice pushed to every peer), CUDA events, no collectives.  All ranks run it at the same
    time (a barrier precedes it), so the exchange figure sees the real NVLink contention."""
    import ctypes as C

    from . import _lib
    from .device import _p, _stream_ptr
    s = _stream_ptr(st.device)
    xin, out = st.full[0], st.full[1][st.row0:st.row0 + st.n_local]

    def timed(fn, reps=5):
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    def local_hop():
        _lib.check(lib.srg_spmm_csr_f32(_p(norm.indptr), _p(norm.indices), _p(norm.data), st.n_local, norm.nnz_bound,
                                        _p(xin), st.ld, _p(out), st.ld, st.f_loc, s))

    res = {"hop_local_only_ms": timed(local_hop)}
    if st.p2p:
        dests = (C.c_void_p * len(st.peers))(*st.peer_ptrs[1])

        def exchange():
            _lib.check(lib.srg_push_rows_f32(_p(out), st.n_local, st.ld, dests, len(st.peers), st.row0, s))
        res["exchange_only_ms"] = timed(exchange)
        res["exchange_GBps_out"] = (len(st.peers) - 1) * st.n_local * st.ld * 4 / res["exchange_only_ms"] / 1e6
    else:
        res["exchange_only_ms"] = 0.0
    return res


def _verify_against_one_gpu(st, a_loc, x_loc, k, f, p1_inputs, dev, sdist, torch, dist):
    """Multi-GPU parity inside the bench run: every rank's rows of EVERY hop are gathered on rank 0 and compared,
    bit for bit, with the same propagation run on rank 0's GPU alone over the whole graph (the single-GPU path the
    parity tests pin to the oracle).  Returns the verify block on rank 0."""
    sdist.start_input_exchange(st, x_loc)
    norm, _ = sdist.dist_sym_norm(st, a_loc, 0.5)
    hops = sdist.propagate_device(st, norm, x_loc, k, keep_hops=True)
    torch.cuda.synchronize()
    want = None
    if st.rank == 0:
        a, x = p1_inputs
        norm1, flags1, _ = dev.sym_norm(dev.upload_csr(a), 0.5)
        want = dev.propagate(norm1, dev.pack_features(torch.from_numpy(x).cuda()), f, k)
        torch.cuda.synchronize()
    ok, rows_checked = True, 0
    slot = torch.zeros((st.rows_per, st.ld), dtype=torch.float32, device="cuda")
    gathered = torch.empty((st.world * st.rows_per, st.ld), dtype=torch.float32, device="cuda")
    for j in range(1, k + 1):
        slot.zero_()
        slot[:st.n_local].copy_(hops[j])
        dist.all_gather_into_tensor(gathered, slot)
        if st.rank == 0:
            for r in range(st.world):
                ri, ci = sdist.grid_coords(r, st.world, st.feat_groups)
                r0, r1 = int(st.starts[ri]), int(st.starts[ri + 1])
                c0, c1 = sdist.feature_slice(f, st.feat_groups, ci)
                got = gathered[r * st.rows_per: r * st.rows_per + (r1 - r0), :c1 - c0]
                ok = ok and bool(torch.equal(got, want[j][r0:r1, c0:c1]))
                rows_checked += r1 - r0
    del hops, gathered, slot, want
    if st.rank != 0:
        return None
    return {"bitwise_vs_p1": ok, "hops_compared": k, "rows_compared_per_hop": rows_checked // max(k, 1),
            "how": "every rank's rows of every hop gathered on rank 0 and compared with torch.equal against the same "
                   "propagation run on rank 0's GPU alone over the whole graph"}


def _allgather_baseline(args, n, f, k, world, rank, ag_inputs, dev, sdist, torch, dist):
    """The same step with the plain row partition and one NCCL all-gather per hop (SRG_DIST_MODE=allgather)."""
    a_host, x_host = ag_inputs
    st2 = sdist.DistState(n, f, world, rank, mode="allgather", feat_groups=1)
    a_d = dev.upload_csr(a_host)
    x_d = dev.pack_features(torch.from_numpy(x_host).cuda())

    def step2():
        sdist.start_input_exchange(st2, x_d)
        norm, _ = sdist.dist_sym_norm(st2, a_d, 0.5)
        return sdist.propagate_device(st2, norm, x_d, k, keep_hops=True)

    for _ in range(3):
        step2()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(1, min(args.steps, 5))
    e0.record()
    for _ in range(reps):
        step2()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    st2.close()
    return {"ms_per_step": float(ms.item()), "partition": f"{world} contiguous row blocks x 1",
            "exchange": "ncclAllGather of the previous hop's slices before each hop"}


def _verify_device_generated(st, a_loc, x_loc, k, f, sdist, synth, torch, dist):
    """Checks for the device-generated workload (no host copy of the graph exists):
      (1) 8 sampled local rows of hop 1 recomputed on the host from the procedural features of their
          neighbour rows (rows owned by every rank: exercises the exchange) and the normalised weights;
      (2) column 0 = sqrt(degree) must survive all K hops (eigenvalue 1).
    Returns a dict (max errors) on rank 0."""
    sdist.start_input_exchange(st, x_loc)
    norm, _ = sdist.dist_sym_norm(st, a_loc, 0.5)
    hops = sdist.propagate_device(st, norm, x_loc, k, keep_hops=True)
    torch.cuda.synchronize()
    errs = torch.zeros(2, dtype=torch.float64, device="cuda")
    rows = np.linspace(0, st.n_local - 1, 8).astype(np.int64) if st.n_local > 0 else np.zeros(0, np.int64)
    ip = norm.indptr.cpu().numpy()
    for rl in rows:
        lo, hi = int(ip[rl]), int(ip[rl + 1])
        cols = norm.indices[lo:hi].cpu().numpy().astype(np.int64)
        vals = norm.data[lo:hi].cpu().numpy().astype(np.float64)
        if hi - lo > 200000:
            continue                                   # a hub row: (2) covers it
        xin = synth.hash_features_host(1, 0, 0, st.f0, st.f_loc, f, rows=cols).astype(np.float64)
        if st.ci == 0:      # column 0 carries sqrt(degree) of the NEIGHBOUR: val = 1/sqrt(d_i d_j) gives it back
            d_i = float(hi - lo)                       # row length of A+I = degree incl. the loop
            xin[:, 0] = 1.0 / (vals * np.sqrt(d_i))    # sqrt(d_j) from the weight itself
        want = (vals[:, None] * xin).sum(0)
        got = hops[1][rl, :st.f_loc].double().cpu().numpy()
        c0 = 1 if st.ci == 0 else 0                    # column 0 is checked by (2)
        err = np.max(np.abs(got[c0:] - want[c0:]) / (np.abs(want[c0:]) + 1e-6))
        errs[0] = max(float(errs[0]), float(err))
    if st.ci == 0 and st.n_local > 0:
        rel = ((hops[k][:, 0] - x_loc[:, 0]).abs() / x_loc[:, 0]).max()
        errs[1] = rel.double()
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    del hops
    return {"hop1_sampled_rows_max_rel_err": float(errs[0]), "sqrt_degree_eigenvector_max_rel_err_after_K_hops": float(errs[1])}


def _e2e(args, st, a_loc_host, x_loc_host, k, f_loc, lib, dev, sdist, torch, dist):
    # end to end: host slices in (pinned), host slices of every hop out, per rank
    x_pin = torch.from_numpy(np.ascontiguousarray(x_loc_host)).pin_memory()
    ip = torch.from_numpy(a_loc_host.indptr).pin_memory()
    ii = torch.from_numpy(a_loc_host.indices).pin_memory()
    dd = torch.from_numpy(a_loc_host.data).pin_memory()
    outs = [torch.empty((st.n_local, f_loc), dtype=torch.float32).pin_memory() for _ in range(k)]

    # device staging allocated once: the timed region is copies + kernels, not cudaMalloc
    d_ip = torch.empty_like(ip, device="cuda")
    d_ii = torch.empty_like(ii, device="cuda")
    d_dd = torch.empty_like(dd, device="cuda")
    d_x = torch.empty_like(x_pin, device="cuda")
    d_flat = torch.empty((st.n_local, f_loc), dtype=torch.float32, device="cuda")
    copy_stream = torch.cuda.Stream()

    from . import _lib
    vt = _lib.SRG_VAL_F64 if dd.dtype == torch.float64 else _lib.SRG_VAL_F32
    state = {"h2d": 0}

    def e2e_step():
        d_ip.copy_(ip, non_blocking=True)
        d_ii.copy_(ii, non_blocking=True)
        d_x.copy_(x_pin, non_blocking=True)
        # the all-ones shortcut of the host pipeline (csrc/host_api.cu): scipy's float64 ones carry no information,
        # they are verified on the host in the shadow of the copies above and uploaded only if a value differs
        ones = lib.srg_host_all_ones(dd.data_ptr(), vt, dd.numel(), 4) == 1
        if not ones:
            d_dd.copy_(dd, non_blocking=True)
        state["h2d"] = int(ip.numel() * 4 + ii.numel() * 4 + x_pin.numel() * 4 + (0 if ones else dd.numel() * dd.element_size()))
        a_d = dev.DeviceCSR(d_ip, d_ii, None if ones else d_dd, st.n_local, int(a_loc_host.nnz))
        xp = dev.pack_features(d_x)
        sdist.start_input_exchange(st, xp)
        norm, _ = sdist.dist_sym_norm(st, a_d, 0.5)
        hops = sdist.propagate_device(st, norm, xp, k, keep_hops=True)
        for o, h in zip(outs, hops[1:]):
            lib.srg_unpack_features_f32(h.data_ptr(), st.ld, d_flat.data_ptr(), f_loc, st.n_local, f_loc,
                                        torch.cuda.current_stream().cuda_stream)
            o.copy_(d_flat, non_blocking=True)
        torch.cuda.synchronize()

    dist.barrier()                             # pinning the host buffers takes a different time on every rank
    e2e_step()
    dist.barrier()
    t0 = time.perf_counter()
    reps = max(1, min(args.steps, 5))
    for _ in range(reps):
        e2e_step()
    dist.barrier()
    t_e2e = torch.tensor([(time.perf_counter() - t0) / reps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    t_e2e = float(t_e2e.item())
    return {"t": t_e2e, "h2d": state["h2d"],
            "d2h": int(k * st.n_local * f_loc * 4)}


def _emit_line(args, st, metric, unit, emit, mode, pf, world, rank, n, nnz_hat, f, f_loc, k, t_step, launches, e2e_line,
               clocks, verify, gen, probe=None):
    if rank == 0:
        from bench import comp_bytes, gather_bytes, measured_peak, workload_config
        peak, peak_src = measured_peak()
        value = k * nnz_hat * f / t_step
        bg = gather_bytes(n, nnz_hat, f)
        # per-GPU roofline: each rank gathers nnz_hat/P rows; the exchange moves (P-1)/P * N*F*4 bytes in
        hop_s = t_step / k
        nvlink_bytes = (st.n_row_blocks - 1) / st.n_row_blocks * n * f_loc * 4
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(workload_config(args, n, nnz_hat, f, k), exchange=mode,
                               partition=f"{st.n_row_blocks} contiguous row blocks x {pf} feature slices"),
                "roofline": {"bound": "hbm", "kernel": ("spmm_bulk_kernel (TMA row gathers, push epilogue)" if (mode in ("push", "push_tma") and st.ld <= 64) else None) or {"push": "spmm_stream_kernel (push epilogue)", "push_tma": "spmm_stream_kernel (TMA bulk-store push epilogue)", "copy": "spmm hop in row chunks + copy-engine exchange"}.get(mode, "spmm_stream_kernel + ncclAllGather"),
                             "achieved": bg / world / hop_s / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": bg / world / hop_s / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                             "note": "per GPU, step time / K (includes the sharded normalisation and the exchange)",
                             "nvlink_GBps_in_per_gpu": nvlink_bytes / hop_s / 1e9,
                             "compulsory_bytes_per_launch": comp_bytes(n, nnz_hat, f) / world},
                "cpu_baseline": None,
                "e2e": None if e2e_line is None else
                {"value": k * nnz_hat * f / e2e_line["t"], "unit": unit, "ms_per_step": e2e_line["t"] * 1e3,
                 "h2d_bytes_per_step": e2e_line["h2d"], "d2h_bytes_per_step": e2e_line["d2h"],
                 "note": "per-rank bytes; max over ranks time"},
                "gpu_launches": int(launches), "clocks": clocks}
        if probe is not None:
            probe["note"] = ("rank 0: hop with the exchange fused (max over ranks), the same hop writing locally only, the "
                             "exchange alone (all ranks pushing at once); overlap_efficiency = max(local, exchange) / fused")
            line["overlap"] = probe
        if verify is not None:
            line["verify"] = verify
        if gen == "device":
            line["data"] = "synthetic (device-generated shards; no host copy exists, so no host end-to-end leg)"
        emit(line)
