"""Data-parallel gradient exchange (absent from the reference -- SURVEY.md 8e).

One process per GPU (torchrun); clips/frames are sharded over ranks with no data-path
collective; the only exchange is one sum all-reduce of each optimiser group's flat fp32
gradient range per update, after which the fused Adam kernel applies grad_scale =
1/world_size.  Two transports:

* peer memory (default on one NVLink / NVSwitch box, 2..8 ranks): every rank's flat gradient
  buffer is mapped into every process (CUDA IPC) and ONE kernel of the library per update
  (gg_dp_allreduce: two-shot reduce-scatter + all-gather by peer loads, summed in rank order)
  reduces the group's range in place -- csrc/dp_allreduce.cu has the measurements behind it;
* NCCL (torch.distributed.all_reduce; GG_DP_P2P=0, gloo on CPU, or when the buffers cannot be
  exported): a few large buckets whose boundaries follow the reverse layer order, so that a
  bucket can start as soon as the backward pass has produced it (grad_ready).

Batch-norm statistics stay per replica (the usual DDP semantics; SURVEY.md section 7,
hard part 9): an N-GPU run equals N replicas of the per-GPU batch with averaged gradients.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class DataParallel:
    def __init__(self, world_size=None, rank=None, backend=None, bucket_bytes=32 << 20, init=True, grad_dtype=None, nvls=None, overlap_update=None):
        """grad_dtype: 'fp32' | 'bf16' | None -- dtype the gradient buckets cross NVLink in.  None = GG_DP_GRAD_DTYPE, else 'bf16'
        when the operator layer computes in bf16 (activation gradients are bf16 already; the buckets are cast on the
        communication stream, summed by NCCL in bf16 and cast back: half the bytes) and 'fp32' in the fp32 parity mode.
        nvls: use NCCL's in-switch reduction (NVLS); None = GG_DP_NVLS, else off (measured slower for these 3-20 MB buckets
        launched from a graph: profiles/r01z_*).  Only takes effect if the process group is created here.
        overlap_update: run an update's last bucket AND its Adam launch on the communication stream, so that the next update's
        generator forward (which does not read the discriminator's weights) overlaps them (GG_DP_OVERLAP_UPDATE, default on)."""
        self.world_size = int(os.environ.get("WORLD_SIZE", "1")) if world_size is None else world_size
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else rank
        self.local_rank = int(os.environ.get("LOCAL_RANK", str(self.rank)))
        self.bucket_bytes = bucket_bytes
        self.comm_stream = None
        self.grad_dtype = grad_dtype or os.environ.get("GG_DP_GRAD_DTYPE") or None      # resolved at first use (needs ops' precision)
        self.overlap_update = (os.environ.get("GG_DP_OVERLAP_UPDATE", "1") != "0") if overlap_update is None else bool(overlap_update)
        self._pending = False
        self._bf16 = None
        self._p2p = {}          # id(grads storage) -> peer tables, or False when the exchange cannot use peer memory
        self.p2p = os.environ.get("GG_DP_P2P", "1") != "0"
        # the peer exchange moves fp32 by default: measured at 8 GPUs (profiles/r4h_scale8_p2p.log) its bf16 all-gather phase is
        # request-rate bound on 8-byte peer loads and SLOWER (109 vs 80 us for 17.3 MB); GG_DP_P2P_WIRE=bf16 selects it anyway
        self._wire16 = 1 if os.environ.get("GG_DP_P2P_WIRE", "fp32") == "bf16" else 0
        if init and self.world_size > 1 and not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                # Measured on 8 x B200 (tools/gpu_scale_ab.sh, profiles/r01z_*): for this step's 3-20 MB gradient buckets,
                # issued from inside a CUDA graph, in-switch reduction (NVLS) is slower than NCCL's ring/tree over
                # NVLink (2.097 vs 2.035 ms/step).  The choice is the constructor's `nvls` argument / GG_DP_NVLS (default
                # off); a NCCL_NVLS_ENABLE already set by the user wins.
                want_nvls = (os.environ.get("GG_DP_NVLS", "0") == "1") if nvls is None else bool(nvls)
                os.environ.setdefault("NCCL_NVLS_ENABLE", "1" if want_nvls else "0")
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world_size)

    # ------------------------------------------------------------------------------
    def buckets(self, begin, end):
        """Split the flat element range [begin, end) into buckets of <= bucket_bytes (fp32)."""
        per = max(1, self.bucket_bytes // 4)
        if any(self._p2p.values()):
            per = 1 << 62          # the peer exchange takes a range whole
        out = []
        b = begin
        while b < end:
            e = min(end, b + per)
            out.append((b, e))
            b = e
        return out

    # -- overlapped exchange ------------------------------------------------------------
    # Backward produces gradients in reverse creation order, and a group's gradients are one contiguous range laid out
    # in creation order: when the filter gradient of layer L has been enqueued, everything at or above L's offset is
    # final.  begin_update() arms a hook that ops._run_wgrad calls after each filter-gradient launch; once the finished
    # tail [offset(L), done_hi) reaches `early_bytes` it is all-reduced on the communication stream while the rest of
    # the backward pass (the remaining dgrad / batch-norm / wgrad kernels) keeps the SMs busy.  allreduce() then only
    # has the head of the range left.  All collectives go to the one communication stream, in the same order on every
    # rank; under CUDA-graph capture the stream fork/join becomes parallel graph branches.
    early_bytes = int(float(os.environ.get("GG_DP_EARLY_MB", "2")) * (1 << 20)) or (1 << 62)   # 0 = no early buckets

    def _comm(self):
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream() if torch.cuda.is_available() and dist.get_backend() == "nccl" else False
        return self.comm_stream

    def _segments(self, optim):
        """The flat range of `optim` as a list of [lo, hi) segments, one per optimiser group.  Backward produces the
        gradients of ONE group in reverse layout order, but the groups of a tuple var_list (e.g. ("dvideo", "d_img") with
        --train_img_disc: the video discriminator's filters come first in the layout AND first in backward, the image
        discriminator's last) are not ordered against each other -- so "everything above this offset is final" holds
        inside a segment only, and each segment keeps its own finished tail."""
        if isinstance(optim.group, (tuple, list)):
            return [list(optim.store.ranges[g]) for g in optim.group]
        return [list(optim.range())]

    def begin_update(self, optim):
        """Call before the backward pass of an update whose gradients `optim` owns."""
        if self.world_size <= 1:
            return
        from . import ops
        # seg = [lo, done_hi]: [done_hi, hi) of the segment has been all-reduced already
        self._cur = dict(optim=optim, segs=self._segments(optim), issued=[])
        if self._peer_tables(optim.store.flat["grads"]):
            return            # peer-memory exchange: one launch over the whole range once backward is done (no early buckets)
        ops.GRAD_READY_HOOK = self.grad_ready

    # -- peer-memory exchange ------------------------------------------------------------
    def _peer_tables(self, grads):
        """Map the flat gradient buffer (and a signal buffer) of every rank into this process; returns the ctypes pointer tables
        of gg_dp_allreduce, or False when this run cannot use them (every rank takes the same branch).  Collective: the first
        call for a given buffer must happen on all ranks, outside CUDA-graph capture (prepare())."""
        key = grads.data_ptr()
        if key in self._p2p:
            return self._p2p[key]
        eligible = (self.p2p and grads.is_cuda and dist.is_initialized() and dist.get_backend() == "nccl"
                    and 2 <= self.world_size <= 8 and self.world_size <= torch.cuda.device_count())      # the same on every rank
        if not eligible:
            self._p2p[key] = False
            return False
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("DataParallel: call prepare(store) (or broadcast_parameters) before capturing a step: the peer-memory "
                               "exchange is set up by a collective handshake")
        import ctypes
        from . import _cabi as c
        L = c.lib()
        mine = sig = stage = None
        try:
            sig = torch.zeros(int(L.gg_dp_signal_bytes()), dtype=torch.uint8, device=grads.device)
            stage = torch.empty(-(-grads.numel() // self.world_size) + 8, dtype=torch.float32, device=grads.device)
            mine = []
            for t in (grads, stage, sig):
                h = ctypes.create_string_buffer(64)
                off = ctypes.c_uint64(0)
                c.check(L.gg_ipc_export(c.ptr(t), h, ctypes.byref(off)), "gg_ipc_export")
                mine.append((h.raw, int(off.value)))
            torch.cuda.synchronize()              # the zero fill of the signals has landed before a peer can write a flag
        except RuntimeError:
            mine = None                           # e.g. an expandable-segments pool: cannot be exported
        everyone = [None] * self.world_size
        dist.all_gather_object(everyone, mine)
        tabs = False
        if all(e is not None for e in everyone):
            try:
                gp, tp, sp = (ctypes.c_void_p * 8)(), (ctypes.c_void_p * 8)(), (ctypes.c_void_p * 8)()
                for r, rec in enumerate(everyone):
                    if r == self.rank:
                        gp[r], tp[r], sp[r] = grads.data_ptr(), stage.data_ptr(), sig.data_ptr()
                        continue
                    for arr, (h, off) in zip((gp, tp, sp), rec):
                        out = ctypes.c_void_p()
                        c.check(L.gg_ipc_import(h, off, ctypes.byref(out)), "gg_ipc_import")
                        arr[r] = out.value
                tabs = dict(grads=gp, stage=tp, sig=sp, keep=(sig, stage))
            except RuntimeError:
                tabs = False
        flag = torch.tensor([1 if tabs else 0], device=grads.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)              # everyone, or no one
        if int(flag.item()) == 0:
            tabs = False
        self._p2p[key] = tabs
        return tabs

    def prepare(self, store):
        """Set up the peer-memory exchange for `store`'s gradient buffer (collective; call once after finalize(), before the
        first captured step).  Without it the first update does it lazily -- which must not happen inside a graph capture."""
        if self.world_size > 1 and store.flat is not None:
            self._peer_tables(store.flat["grads"])

    def _use_bf16(self):
        if self.grad_dtype is None:
            from . import ops
            self.grad_dtype = "bf16" if ops.get_precision() == "bf16" else "fp32"
        return self.grad_dtype == "bf16"

    def _reduce(self, grads, x, y):
        """Sum-all-reduce grads[x:y] in place (on the current stream)."""
        tabs = self._peer_tables(grads)
        if tabs:
            from . import _cabi as c
            c.check(c.lib().gg_dp_allreduce(tabs["grads"], tabs["stage"], tabs["sig"], self.rank, self.world_size, x, y - x, self._wire16, c.stream()), "gg_dp_allreduce")
            return
        if self._use_bf16() and grads.is_cuda:
            from . import ops
            if self._bf16 is None or self._bf16.numel() != grads.numel():
                self._bf16 = torch.empty(grads.numel(), dtype=torch.bfloat16, device=grads.device)
            c, L = ops.cabi, ops.cabi.lib()
            half = self._bf16[x:y]
            ops.check(L.gg_cast(c.ptr(grads[x:y]), c.GG_F32, c.ptr(half), c.GG_BF16, y - x, c.stream()), "gg_cast")
            dist.all_reduce(half, op=dist.ReduceOp.SUM)
            ops.check(L.gg_cast(c.ptr(half), c.GG_BF16, c.ptr(grads[x:y]), c.GG_F32, y - x, c.stream()), "gg_cast")
        else:
            dist.all_reduce(grads[x:y], op=dist.ReduceOp.SUM)

    def _issue(self, lo, hi, extra_streams=()):
        grads = self._cur["optim"].store.flat["grads"]
        comm = self._comm()
        if comm is False:                          # gloo / CPU: no streams, reduce in place
            for (x, y) in self.buckets(lo, hi):
                self._reduce(grads, x, y)
        else:
            comm.wait_stream(torch.cuda.current_stream())
            for s in extra_streams:
                comm.wait_stream(s)
            with torch.cuda.stream(comm):
                for (x, y) in self.buckets(lo, hi):
                    self._reduce(grads, x, y)
        self._cur["issued"].append((lo, hi))

    def grad_ready(self, var, producer_stream=None):
        """Hook: `var`'s filter gradient has just been enqueued (on producer_stream, if not the current stream)."""
        cur = getattr(self, "_cur", None)
        if cur is None:
            return
        seg = next((s for s in cur["segs"] if s[0] <= var.offset < s[1]), None)
        if seg is None:
            # not inside the unreduced part of any segment: either another optimiser's variable, or one whose bucket is gone
            b, e = cur["optim"].range()
            if b <= var.offset < e:
                raise RuntimeError(f"{var.name}: a gradient was produced after its bucket had been all-reduced "
                                   "(a filter used at two call sites of one update needs early_bytes = inf)")
            return
        if (seg[1] - var.offset) * 4 < self.early_bytes:
            return
        self._issue(var.offset, seg[1], (producer_stream,) if producer_stream is not None else ())
        seg[1] = var.offset

    def _finish_buckets(self, optim):
        """Issue what begin_update's early buckets have not covered yet (no wait)."""
        from . import ops
        ops.GRAD_READY_HOOK = None
        cur = getattr(self, "_cur", None)
        if cur is None or cur["optim"] is not optim:
            cur = self._cur = dict(optim=optim, segs=self._segments(optim), issued=[])
        # the unreduced heads; adjacent ones (a tuple group with no early bucket in between) go out as one range
        merged = []
        for lo, hi in [(lo, hi) for lo, hi in cur["segs"] if hi > lo]:
            if merged and merged[-1][1] == lo:
                merged[-1] = (merged[-1][0], hi)
            else:
                merged.append((lo, hi))
        for lo, hi in merged:
            self._issue(lo, hi)
        self.last_buckets = cur["issued"]
        self._cur = None

    def allreduce(self, optim):
        """Sum-all-reduce the gradient range of `optim`'s group (what begin_update's early buckets have not covered yet)
        and make the current stream wait for the exchange.  Adam divides by world_size."""
        if self.world_size <= 1:
            return
        self._finish_buckets(optim)
        comm = self._comm()
        if comm is not False:
            torch.cuda.current_stream().wait_stream(comm)

    def finish_update(self, optim, apply_fn):
        """The tail of an update: remaining buckets, then `apply_fn()` (the optimiser's fused Adam launch).  With
        overlap_update both go to the communication stream and the current stream does NOT wait: what follows on it -- the next
        update's generator forward, which reads none of these weights -- overlaps the exchange and the Adam pass; the caller
        must call wait_pending() before the first kernel that reads the updated variables (or their gradients' buffer)."""
        if self.world_size <= 1:
            apply_fn()
            return
        comm = self._comm()
        if not self.overlap_update or comm is False:
            self.allreduce(optim)
            apply_fn()
            return
        self._finish_buckets(optim)
        comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(comm):
            apply_fn()
        self._pending = True

    def wait_pending(self):
        """Make the current stream wait for a finish_update() still in flight on the communication stream."""
        if self._pending:
            torch.cuda.current_stream().wait_stream(self._comm())
            self._pending = False

    def broadcast_parameters(self, store):
        """Make every rank start from rank 0's variables (weights and EMAs); also sets up the peer-memory exchange."""
        self.prepare(store)
        if self.world_size <= 1:
            return
        dist.broadcast(store.flat["params"], src=0)
        for v in store.vars.values():
            v.version += 1                          # bf16 copies of the filters are re-cast at their next use

    def shard(self, n_items):
        """Contiguous shard [lo, hi) of a global batch of n_items clips/frames for this rank."""
        per = n_items // self.world_size
        return self.rank * per, (self.rank + 1) * per

    def barrier(self):
        if self.world_size > 1:
            dist.barrier()

    def max_over_ranks(self, value: float) -> float:
        if self.world_size <= 1:
            return value
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([value], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
