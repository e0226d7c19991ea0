"""Row-partitioned multi-GPU propagation: one process per GPU, one exchange per hop.

The reference has no multi-GPU code; this is the scheme SURVEY.md §8e defines:

  * partition map: contiguous equal row blocks, ``rows_per = ceil(N / P)``, rank p owns
    ``[p*rows_per, min(N, (p+1)*rows_per))``; column ids stay global.
  * normalisation: stage 1/2a on the local rows, ONE all-gather of the degree vector (N x 8 B),
    power tables and values locally (``srg_*_rows_csr``).  Symmetry of the adjacency is the
    caller's promise here (the mirror rows live on other ranks).
  * hop k: every rank needs all of X_{k-1}.  Two exchange modes:
      - ``"allgather"``: NCCL all-gather of the slices (torch.distributed), then the local SpMM;
      - ``"push"``: the SpMM epilogue stores every finished row into the next-hop buffer of every
        rank over NVLink peer mappings (CUDA IPC), one stream-ordered tiny all-reduce orders the
        hops across ranks.  No separate all-gather kernel runs.
      - ``"copy"``: the hop runs in row chunks that write only the local buffer; as soon as a chunk is
        done the copy engines move it to every peer's buffer (cudaMemcpyAsync on peer mappings, a second
        stream), so the exchange costs no SM / load-store bandwidth and only the last chunk's copy is
        exposed.
  * every output row is owned by one rank and reduced in CSR order => bitwise equal to 1 GPU.

The orchestration is written against a small ``ops`` object so the CPU test-suite can drive the
same code over gloo with the oracle as the local hop (tests/test_dist_cpu.py).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

__all__ = ["row_partition", "grid_coords", "feature_slice", "push_peers", "shard_rows", "propagate_sharded",
           "DeviceOps", "DistState", "dist_sym_norm"]


def row_partition(n: int, world: int):
    """(rows_per, starts[world+1]) of the contiguous equal-block partition."""
    rows_per = -(-int(n) // int(world)) if n > 0 else 0
    starts = np.minimum(np.arange(world + 1, dtype=np.int64) * rows_per, n)
    return rows_per, starts


def grid_coords(rank: int, world: int, feat_groups: int = 1):
    """rank -> (row block index, feature slice index) of the P_r x P_f grid, rank = ri * P_f + ci.
    P_f = 1 is the plain row partition; P_f > 1 additionally splits the feature columns, which divides
    the per-hop exchange volume by P_f (the exchange is NVLink-ingress bound at 8 GPUs)."""
    if world % feat_groups:
        raise ValueError("world size must be a multiple of feat_groups")
    return rank // feat_groups, rank % feat_groups


def feature_slice(f: int, feat_groups: int, ci: int):
    """[f0, f1) of feature slice ci: contiguous blocks of ceil(F / P_f) columns."""
    per = -(-int(f) // int(feat_groups))
    f0 = min(ci * per, f)
    return f0, min(f, f0 + per)


def push_peers(rank: int, world: int, feat_groups: int = 1):
    """Ranks that hold the same feature slice (one per row block, in row-block order): the
    destinations of this rank's output rows."""
    _, ci = grid_coords(rank, world, feat_groups)
    return [r * feat_groups + ci for r in range(world // feat_groups)]


def shard_rows(adj: sp.csr_matrix, start: int, end: int) -> sp.csr_matrix:
    """Rows [start, end) of a CSR with GLOBAL column ids (shape (end-start) x N)."""
    lo, hi = int(adj.indptr[start]), int(adj.indptr[end])
    indptr = (adj.indptr[start:end + 1] - adj.indptr[start]).astype(np.int32)
    return sp.csr_matrix((adj.data[lo:hi], adj.indices[lo:hi], indptr), shape=(end - start, adj.shape[1]), copy=False)


def propagate_sharded(ops, local_norm, x_local, k, rows_per, world):
    """K hops over this rank's rows.  ``ops`` provides:
         ops.new_full()                 -> buffer holding all world*rows_per rows
         ops.local_view(full)           -> this rank's slice of a full buffer (rows_per rows)
         ops.load_local(full, x_local)  -> copy the rank's input rows into its slice
         ops.exchange(full)             -> make every rank's slice visible in `full` on all ranks
         ops.hop(local_norm, full_in, full_out) -> write this rank's rows of A^ X into full_out
                                           (push mode: into every rank's full_out, then ops.exchange
                                            only synchronises)
         ops.snapshot_local(full)       -> the rank's rows as an independent array/tensor
    Returns [hop_0_local, ..., hop_k_local]."""
    cur, nxt = ops.new_full(), ops.new_full()
    ops.load_local(cur, x_local)
    ops.exchange(cur)
    out = [ops.snapshot_local(cur)]
    for _ in range(k):
        ops.hop(local_norm, cur, nxt)
        ops.exchange(nxt)
        out.append(ops.snapshot_local(nxt))
        cur, nxt = nxt, cur
    return out


# ------------------------------------------------------------------------------------------------
# device implementation
# ------------------------------------------------------------------------------------------------
class DistState:
    """Per-rank device state: partition, the two full feature buffers (peer-mapped in push mode)."""

    def __init__(self, n, f, world, rank, mode="push", group=None, device=None, feat_groups=1):
        import torch
        import torch.distributed as dist

        from . import _lib
        from .device import pad_ld
        self.torch, self.dist, self._lib = torch, dist, _lib
        self.lib = _lib.load()
        if mode == "push_tma":      # push hop whose epilogue is one TMA bulk store per peer (csrc/spmm.cu)
            mode = "push"
            _lib.set_tuning("push_tma", 1)
        elif mode == "push":
            # rows wider than the bulk-gather kernel takes (LDGSTS stream kernel): the bulk-store tile epilogue is the
            # default (4 GPUs, 4 x 1: 5.52 vs 5.66 ms per step; 2 GPUs: 2.27 vs 2.30 ms per hop); SRG_PUSH_TMA=0 = st.global
            _lib.set_tuning("push_tma", int(os.environ.get("SRG_PUSH_TMA", "1")))
        self.n, self.f, self.world, self.rank, self.mode, self.group = n, f, world, rank, mode, group
        self.feat_groups = int(feat_groups)
        self.ri, self.ci = grid_coords(rank, world, self.feat_groups)
        self.n_row_blocks = world // self.feat_groups
        if mode not in ("push", "copy") and self.feat_groups != 1:
            raise ValueError("the all-gather exchange supports the plain row partition only (feat_groups = 1)")
        self.rows_per, self.starts = row_partition(n, self.n_row_blocks)
        self.row0 = int(self.starts[self.ri])
        self.n_local = int(self.starts[self.ri + 1] - self.starts[self.ri])
        self.f0, self.f1 = feature_slice(f, self.feat_groups, self.ci)
        self.f_loc = self.f1 - self.f0
        self.ld = pad_ld(self.f_loc)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.n_pad = self.rows_per * self.n_row_blocks
        self.peers = push_peers(rank, world, self.feat_groups)
        self._raw = []          # (ptr) owned IPC allocations
        self._opened = []       # peer mappings
        self.full = []          # torch views of the two local full buffers
        self.peer_ptrs = []     # per buffer: device pointers of the push peers (index = row block)
        self._tick = torch.zeros(1, dtype=torch.int32, device=self.device)
        # the input exchange runs on a side stream so that it overlaps the normalisation
        self.side = torch.cuda.Stream(device=self.device)
        self._x_event = None
        self.p2p = mode in ("push", "copy") and world > 1     # peer-mapped full buffers
        self.comm = torch.cuda.Stream(device=self.device)     # copy mode: the DMA exchange stream
        self.peer_streams = [torch.cuda.Stream(device=self.device) for _ in self.peers]   # input exchange: one per peer
        self.n_chunks = int(os.environ.get("SRG_COPY_CHUNKS", "4"))   # copy mode: row chunks per hop
        nbytes = self.n_pad * self.ld * 4
        for _ in range(2):
            if self.p2p:
                p = C.c_void_p()
                _lib.check(self.lib.srg_ipc_alloc(C.byref(p), max(nbytes, 256)))
                self._raw.append(p)
                t = self._wrap(p.value, nbytes)
            else:
                t = torch.zeros((self.n_pad, self.ld), dtype=torch.float32, device=self.device)
            self.full.append(t)
        # cross-rank ordering of the push hops: "flags" = peer-mapped epoch slots + one tiny kernel per hop
        # (srg_peer_barrier); "nccl" = a 4-byte all-reduce per hop (the round-1 form, kept for comparison)
        self.fence = os.environ.get("SRG_DIST_FENCE", "flags") if self.p2p else "nccl"
        self._epoch = 0
        self._flag_ptrs = None
        self._timeout = torch.zeros(1, dtype=torch.int32, device=self.device)
        if self.p2p:
            for b in range(2):
                self.full[b].zero_()
            fl = C.c_void_p()
            _lib.check(self.lib.srg_ipc_alloc(C.byref(fl), 256))
            self._raw.append(fl)
            self._flags_local = fl
            torch.cuda.synchronize()
            # zero the epoch slots before anybody can see the mapping
            tmp = torch.zeros(64, dtype=torch.int32, device=self.device)
            _lib.check(self.lib.srg_copy_async(fl, C.c_void_p(tmp.data_ptr()), 256, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
            torch.cuda.synchronize()
            h = (C.c_ubyte * 64)()
            _lib.check(self.lib.srg_ipc_get_handle(fl, h))
            handles = [None] * world
            dist.all_gather_object(handles, bytes(h), group=group)
            fptrs = []
            for r in self.peers:
                if r == rank:
                    fptrs.append(fl.value)
                else:
                    q = C.c_void_p()
                    hb = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                    _lib.check(self.lib.srg_ipc_open(hb, C.byref(q)))
                    self._opened.append(q)
                    fptrs.append(q.value)
            self._flag_ptrs = fptrs
            for b in range(2):
                h = (C.c_ubyte * 64)()
                _lib.check(self.lib.srg_ipc_get_handle(self._raw[b], h))
                handles = [None] * world
                dist.all_gather_object(handles, bytes(h), group=group)
                ptrs = []
                for r in self.peers:
                    if r == rank:
                        ptrs.append(self._raw[b].value)
                    else:
                        q = C.c_void_p()
                        hb = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                        _lib.check(self.lib.srg_ipc_open(hb, C.byref(q)))
                        self._opened.append(q)
                        ptrs.append(q.value)
                self.peer_ptrs.append(ptrs)

    def _wrap(self, ptr, nbytes):
        """torch view over a raw device allocation (no ownership)."""
        torch = self.torch

        class _Holder:
            pass
        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (self.n_pad, self.ld), "typestr": "<f4", "data": (int(ptr), False),
                                      "version": 2, "strides": None}
        return torch.as_tensor(h, device=self.device)

    def peer_fence(self):
        """Stream-ordered barrier among the ranks that exchange rows with this one (same feature slice)."""
        from .device import _p, _stream_ptr
        if self.world == 1:
            return
        if self.fence == "flags" and self._flag_ptrs is not None:
            self._epoch += 1
            ptrs = (C.c_void_p * len(self._flag_ptrs))(*self._flag_ptrs)
            self._lib.check(self.lib.srg_peer_barrier(self._flags_local, ptrs, len(self._flag_ptrs), self.ri, self._epoch,
                                                      _p(self._timeout), _stream_ptr(self.device)))
        else:
            self.dist.all_reduce(self._tick, group=self.group)

    def check_fence(self):
        """Host-side check (synchronises): a peer that never showed up at a flag barrier raises instead of hanging."""
        if int(self._timeout.item()):
            raise RuntimeError("multi-GPU hop: a peer did not reach the flag barrier within 2 s")

    def close(self):
        self.torch.cuda.synchronize()
        if self.world > 1 and self.dist.is_initialized():
            self.dist.barrier(group=self.group)
        for q in self._opened:
            self.lib.srg_ipc_close(q)
        self.full = []
        for p in self._raw:
            self.lib.srg_ipc_free(p)
        self._opened, self._raw = [], []


class DeviceOps:
    """``ops`` for propagate_sharded on the GPU."""

    def __init__(self, st: DistState):
        self.st = st
        self._next = 0

    def new_full(self):
        i = self._next
        self._next += 1
        assert i < 2, "two full buffers ping-pong"
        return i

    def local_view(self, i):
        st = self.st
        return st.full[i][st.row0:st.row0 + st.rows_per] if st.n_local == st.rows_per else \
            st.full[i][st.rank * st.rows_per:(st.rank + 1) * st.rows_per]

    def load_local(self, i, x_local_padded):
        st = self.st
        st.full[i][st.row0:st.row0 + st.n_local].copy_(x_local_padded)

    def exchange(self, i, pushed=False):
        st = self.st
        if st.world == 1:
            return
        if pushed:
            # rows are already in every peer's buffer: order the hops across ranks on the stream
            st.peer_fence()
        else:
            view = st.full[i][st.rank * st.rows_per:(st.rank + 1) * st.rows_per]
            st.dist.all_gather_into_tensor(st.full[i], view, group=st.group)

    def hop(self, local_norm, i_in, i_out, keep=None, last=False):
        """One hop.  ``keep`` (push mode): an n_local x ld tensor that receives this rank's rows as one more
        destination of the epilogue - the hop's element of the K+1 list, without a clone afterwards.
        ``last``: nobody gathers from the result of the final hop, so with a ``keep`` destination it is written there
        only (no exchange, no fence)."""
        from . import _lib
        from .device import _p, _stream_ptr
        st = self.st
        xin = st.full[i_in]
        if last and keep is not None and st.mode == "push" and st.world > 1:
            _lib.check(st.lib.srg_spmm_csr_f32(_p(local_norm.indptr), _p(local_norm.indices), _p(local_norm.data),
                                               st.n_local, local_norm.nnz_bound, _p(xin), st.ld, _p(keep), st.ld, st.f_loc,
                                               _stream_ptr(st.device)))
            self._pushed = False
            return
        if st.mode == "copy" and st.world > 1:
            self._hop_copy(local_norm, xin, i_out)
            self._pushed = True
        elif st.mode == "push" and st.world > 1:
            ptrs = list(st.peer_ptrs[i_out])
            row0s = [st.row0] * len(ptrs)
            if keep is not None:
                assert keep.shape == (st.n_local, st.ld) and keep.is_contiguous()
                ptrs.append(keep.data_ptr())
                row0s.append(0)
            dests = (C.c_void_p * len(ptrs))(*ptrs)
            offs = (C.c_int64 * len(ptrs))(*row0s)
            _lib.check(st.lib.srg_spmm_csr_f32_push2(_p(local_norm.indptr), _p(local_norm.indices), _p(local_norm.data),
                                                     st.n_local, local_norm.nnz_bound, _p(xin), st.ld, dests, offs,
                                                     len(ptrs), st.ld, st.f_loc, _stream_ptr(st.device)))
            self._pushed = True
        else:
            out = st.full[i_out][st.row0:st.row0 + st.n_local]
            _lib.check(st.lib.srg_spmm_csr_f32(_p(local_norm.indptr), _p(local_norm.indices), _p(local_norm.data),
                                               st.n_local, local_norm.nnz_bound, _p(xin), st.ld, _p(out), st.ld, st.f_loc,
                                               _stream_ptr(st.device)))
            self._pushed = False

    def _hop_copy(self, local_norm, xin, i_out):
        """Chunked local hop + copy-engine exchange: chunk c is copied to the peers while chunk c+1 computes."""
        import torch

        from . import _lib
        from .device import _p, _stream_ptr
        st = self.st
        cur = torch.cuda.current_stream(st.device)
        out_base = st.full[i_out].data_ptr()
        row_bytes = st.ld * 4
        n_chunks = max(1, min(st.n_chunks, st.n_local))
        per = -(-st.n_local // n_chunks) if st.n_local else 0
        per = (per + 3) // 4 * 4                                   # whole row groups of the stream kernel
        ip = local_norm.indptr.data_ptr()
        r0 = 0
        while r0 < st.n_local:
            r1 = min(st.n_local, r0 + per)
            y = out_base + (st.row0 + r0) * row_bytes
            _lib.check(st.lib.srg_spmm_csr_f32(C.c_void_p(ip + 4 * r0), _p(local_norm.indices), _p(local_norm.data),
                                               r1 - r0, local_norm.nnz_bound, _p(xin), st.ld, C.c_void_p(y), st.ld,
                                               st.f_loc, _stream_ptr(st.device)))
            ev = torch.cuda.Event()
            ev.record(cur)
            st.comm.wait_event(ev)
            cs = C.c_void_p(st.comm.cuda_stream)
            for blk, peer in enumerate(st.peers):
                if peer == st.rank:
                    continue
                dst = st.peer_ptrs[i_out][blk] + (st.row0 + r0) * row_bytes
                _lib.check(st.lib.srg_copy_async(C.c_void_p(dst), C.c_void_p(y), (r1 - r0) * row_bytes, cs))
            r0 = r1
        done = torch.cuda.Event()
        done.record(st.comm)
        cur.wait_event(done)

    def snapshot_local(self, i):
        st = self.st
        return st.full[i][st.row0:st.row0 + st.n_local].clone()


def start_input_exchange(st: DistState, x_local_padded):
    """Start moving this rank's input rows into full buffer 0 of EVERY rank on a side stream; call it
    BEFORE dist_sym_norm so the exchange overlaps the normalisation.  Push mode: one kernel storing to
    the peer mappings (no collective); all-gather mode: copy + NCCL all-gather."""
    import torch

    from . import _lib
    from .device import _p
    side = st.side
    side.wait_stream(torch.cuda.current_stream(st.device))
    input_copy = st.world > 1 and (st.mode == "copy" or (st.mode == "push" and os.environ.get("SRG_INPUT_XCHG", "push") == "copy"))
    with torch.cuda.stream(side):
        if st.p2p:
            # the previous call's last hop (which ends without a fence) may still be reading the buffer this
            # exchange overwrites on a slower peer: order against it here, on the side stream, so that the
            # normalisation on the main stream does not wait for anybody
            st.peer_fence()
        if input_copy:
            # copy engines, one stream per peer so the copies run side by side (SRG_INPUT_XCHG=copy; measured slower
            # than the lean one-block-per-SM push kernel: 0.51 GB in 1.24 ms against 0.76 ms at 2 GPUs)
            st.full[0][st.row0:st.row0 + st.n_local].copy_(x_local_padded)
            src = st.full[0].data_ptr() + st.row0 * st.ld * 4
            ready = torch.cuda.Event()
            ready.record(side)
            for blk, peer in enumerate(st.peers):
                if peer != st.rank:
                    ps = st.peer_streams[blk]
                    ps.wait_event(ready)
                    _lib.check(st.lib.srg_copy_async(C.c_void_p(st.peer_ptrs[0][blk] + st.row0 * st.ld * 4), C.c_void_p(src),
                                                     st.n_local * st.ld * 4, C.c_void_p(ps.cuda_stream)))
                    side.wait_stream(ps)
        elif st.mode == "push" and st.world > 1:
            dests = (C.c_void_p * len(st.peers))(*st.peer_ptrs[0])
            _lib.check(st.lib.srg_push_rows_f32(_p(x_local_padded), st.n_local, st.ld, dests, len(st.peers), st.row0,
                                                C.c_void_p(side.cuda_stream)))
        else:
            st.full[0][st.row0:st.row0 + st.n_local].copy_(x_local_padded)
        ev = torch.cuda.Event()
        ev.record(side)
    x_local_padded.record_stream(side)
    st._x_event = ev


def propagate_device(st: DistState, local_norm, x_local_padded, k, keep_hops=True, marks=None):
    """K hops on the device; returns the list of local hop slices (or only the last when not keep_hops).
    If start_input_exchange was called for this input, only its completion is awaited here.
    ``marks``: optional list that receives (label, recorded CUDA event) pairs after each stage (bench breakdown)."""
    import torch
    ops = DeviceOps(st)
    cur, nxt = 0, 1

    def mark(label):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((label, ev))
    if st._x_event is None:
        start_input_exchange(st, x_local_padded)
    torch.cuda.current_stream(st.device).wait_event(st._x_event)
    st._x_event = None
    mark("input exchange landed")
    if st.world > 1:
        if st.mode in ("push", "copy"):
            st.peer_fence()                                    # every rank's rows have landed everywhere
        else:
            ops.exchange(cur)
    mark("fence")
    fused_keep = keep_hops and st.mode == "push" and st.world > 1
    out = [x_local_padded if fused_keep else ops.snapshot_local(cur)] if keep_hops else []
    for j in range(k):
        keep = torch.empty((st.n_local, st.ld), dtype=torch.float32, device=st.device) if fused_keep else None
        last = fused_keep and j == k - 1
        ops.hop(local_norm, cur, nxt, keep=keep, last=last)
        mark(f"hop {j + 1} kernel")
        if last:
            # a later call reuses the full buffers: its first fence orders them against this hop's reads
            out.append(keep)
            break
        ops.exchange(nxt, pushed=(st.mode in ("push", "copy") and st.world > 1))
        if keep_hops:
            out.append(keep if fused_keep else ops.snapshot_local(nxt))
        mark(f"hop {j + 1} fence")
        cur, nxt = nxt, cur
    if not keep_hops:
        out = [ops.snapshot_local(cur)]
    return out


def dist_sym_norm(st: DistState, a_local, r, ppr_alpha=None, marks=None):
    """Normalise this rank's rows (raw DeviceCSR with global column ids).  Returns (DeviceCSR with
    float32 values, flags tensor).  ``marks``: optional list receiving (label, CUDA event) after each stage."""
    import torch

    from . import _lib
    from .device import DeviceCSR, _p, _stream_ptr
    lib, dev = st.lib, st.device
    s = _stream_ptr(dev)

    def mark(label):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((label, ev))
    n_loc, nnz = st.n_local, a_local.nnz
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    at_indptr = torch.empty(n_loc + 1, dtype=torch.int32, device=dev)
    _lib.check(lib.srg_selfloop_rows_csr(_p(a_local.indptr), _p(a_local.indices), _p(a_local.data), a_local.val_dtype,
                                         n_loc, nnz, st.row0, st.n, _p(at_indptr), None, _p(flags), s))
    mark("norm: count + scan")
    cap = max(nnz + n_loc, 1)
    at_indices = torch.empty(cap, dtype=torch.int32, device=dev)
    at_val = torch.empty(cap, dtype=torch.float64, device=dev) if a_local.data is not None else None
    # degrees of every row block; with feature groups the gather runs over the whole world and the
    # P_f duplicates of each block are dropped afterwards
    gath = torch.zeros(st.world * st.rows_per, dtype=torch.float64, device=dev)
    deg_loc = gath[st.rank * st.rows_per: st.rank * st.rows_per + max(n_loc, 0)]
    _lib.check(lib.srg_selfloop_fill_rows_csr(_p(a_local.indptr), _p(a_local.indices), _p(a_local.data),
                                              a_local.val_dtype, n_loc, nnz, st.row0, st.n, _p(at_indptr), _p(at_indices),
                                              _p(at_val), _p(deg_loc), _p(flags), s))
    mark("norm: fill + degrees")
    if st.world > 1:
        view = gath[st.rank * st.rows_per:(st.rank + 1) * st.rows_per]
        st.dist.all_gather_into_tensor(gath, view, group=st.group)
    mark("norm: degree all-gather (first cross-rank wait of the step)")
    if st.feat_groups > 1:
        deg_all = gath.view(st.n_row_blocks, st.feat_groups, st.rows_per)[:, st.ci, :].contiguous().view(-1)
    else:
        deg_all = gath
    dl = torch.empty(st.n_pad, dtype=torch.float64, device=dev)
    dr = torch.empty(st.n_pad, dtype=torch.float64, device=dev)
    _lib.check(lib.srg_pow_tables_f64(_p(deg_all), st.n_pad, float(r), _p(dl), _p(dr), s))
    val32 = torch.empty(cap, dtype=torch.float32, device=dev)
    alpha = -1.0 if ppr_alpha is None else float(ppr_alpha)
    _lib.check(lib.srg_norm_values_rows_csr(_p(at_indptr), _p(at_indices), _p(at_val), _p(deg_loc), n_loc, nnz + n_loc,
                                            st.row0, st.n_pad, _p(dl), _p(dr), alpha, 0, None, _p(val32), _p(flags), s))
    mark("norm: power tables + values")
    return DeviceCSR(at_indptr, at_indices, val32, n_loc, nnz + n_loc), flags


# ------------------------------------------------------------------------------------------------
# native handle (csrc/dist.cu): the all-gather scheme without torch.distributed
# ------------------------------------------------------------------------------------------------
def native_unique_id() -> bytes:
    """128-byte NCCL unique id (rank 0 creates it, the caller hands it to every rank)."""
    from . import _lib
    buf = (C.c_ubyte * 128)()
    _lib.check(_lib.load().srg_dist_unique_id(buf))
    return bytes(buf)


class NativeDist:
    """``srg_dist_init`` / ``srg_dist_propagate``: NCCL communicator, partition and the two full
    feature buffers live in libsrgnn_b200.so; one call = sharded normalisation (one degree all-gather) + K hops with
    one ``ncclAllGather`` each.  torch only provides the device tensors and the stream."""

    def __init__(self, unique_id: bytes, world: int, rank: int, n: int, f: int):
        from . import _lib
        self._lib = _lib
        self.lib = _lib.load()
        self.world, self.rank, self.n, self.f = world, rank, n, f
        self._h = C.c_void_p()
        idbuf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        _lib.check(self.lib.srg_dist_init(idbuf, world, rank, n, f, C.byref(self._h)))
        row0, n_local, rows_per, ld = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self.lib.srg_dist_partition(self._h, C.byref(row0), C.byref(n_local), C.byref(rows_per), C.byref(ld)))
        self.row0, self.n_local, self.rows_per, self.ld = row0.value, n_local.value, rows_per.value, ld.value

    def propagate(self, a_local, x_local, k, r=0.5, ppr_alpha=None):
        """``a_local``: DeviceCSR of this rank's rows of the raw adjacency (global column ids); ``x_local``: cuda
        float32 n_local x F.  Returns (list of K+1 cuda tensors n_local x F, flags tensor)."""
        import torch

        from .device import _p, _stream_ptr
        dev = x_local.device
        x_local = x_local.contiguous()
        outs = [torch.empty((self.n_local, self.f), dtype=torch.float32, device=dev) for _ in range(k + 1)]
        ptrs = (C.c_void_p * (k + 1))(*[o.data_ptr() for o in outs])
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        alpha = -1.0 if ppr_alpha is None else float(ppr_alpha)
        self._lib.check(self.lib.srg_dist_propagate(self._h, _p(a_local.indptr), _p(a_local.indices), _p(a_local.data),
                                                    a_local.val_dtype, a_local.nnz, _p(x_local), x_local.stride(0), int(k),
                                                    float(r), alpha, ptrs, self.f, _p(flags), _stream_ptr(dev)))
        return outs, flags

    def close(self):
        if self._h:
            self.lib.srg_dist_destroy(self._h)
            self._h = C.c_void_p()
