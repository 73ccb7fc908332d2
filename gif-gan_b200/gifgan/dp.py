"""Data-parallel gradient exchange (absent from the reference -- SURVEY.md 8e).

One process per GPU (torchrun); clips/frames are sharded over ranks with no data-path
collective; the only exchange is one sum all-reduce of each optimiser group's flat fp32
gradient range per update (NCCL over NVLink 5 / NVSwitch), after which the fused Adam
kernel applies grad_scale = 1/world_size.  Because a group's gradients are one
contiguous range of the VariableStore's flat buffer, the all-reduce is issued in a few
large buckets (launch-latency-bound at these sizes: D 17.3 MB, G 20.5 MB) -- bucket
boundaries follow the reverse layer order so a bucket can start as soon as the backward
pass has produced it (see DataParallel.allreduce_ready).

Batch-norm statistics stay per replica (the usual DDP semantics; SURVEY.md section 7,
hard part 9): an N-GPU run equals N replicas of the per-GPU batch with averaged gradients.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class DataParallel:
    def __init__(self, world_size=None, rank=None, backend=None, bucket_bytes=32 << 20, init=True):
        self.world_size = int(os.environ.get("WORLD_SIZE", "1")) if world_size is None else world_size
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else rank
        self.local_rank = int(os.environ.get("LOCAL_RANK", str(self.rank)))
        self.bucket_bytes = bucket_bytes
        self.comm_stream = None
        if init and self.world_size > 1 and not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world_size)

    # ------------------------------------------------------------------------------
    def buckets(self, begin, end):
        """Split the flat element range [begin, end) into buckets of <= bucket_bytes (fp32)."""
        per = max(1, self.bucket_bytes // 4)
        out = []
        b = begin
        while b < end:
            e = min(end, b + per)
            out.append((b, e))
            b = e
        return out

    def allreduce(self, optim):
        """Sum-all-reduce the gradient range of `optim`'s group.  Adam divides by world_size."""
        if self.world_size <= 1:
            return
        grads = optim.store.flat["grads"]
        b, e = optim.range()
        for (x, y) in self.buckets(b, e):
            dist.all_reduce(grads[x:y], op=dist.ReduceOp.SUM)

    def broadcast_parameters(self, store):
        """Make every rank start from rank 0's variables (weights and EMAs)."""
        if self.world_size <= 1:
            return
        dist.broadcast(store.flat["params"], src=0)

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
