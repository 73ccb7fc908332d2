"""Data-parallel host logic on CPU: world_size-2 `gloo` processes exercising gifgan.dp.DataParallel -- sharding,
bucketed gradient all-reduce over an optimiser group's flat range, parameter broadcast, max-over-ranks timing
reduction.  (The kernels need a GPU; the exchange logic does not.)"""
import os
import socket
import sys
from collections import OrderedDict

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "gif-gan_b200")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from gifgan import ops
    from gifgan.dp import DataParallel
    dp = DataParallel(backend="gloo")
    assert dp.world_size == world and dp.rank == rank
    # shard a global batch of 8 clips: contiguous, disjoint, complete
    lo, hi = dp.shard(8)
    assert (lo, hi) == (rank * 4, rank * 4 + 4)
    # a store with two optimiser groups; gradients differ per rank
    st = ops.VariableStore(device="cpu", seed=rank)
    a = st.get_variable("d_a", [1000], lambda r, s: np.zeros(s, dtype=np.float32))
    b = st.get_variable("d_b", [37], lambda r, s: np.zeros(s, dtype=np.float32))
    c = st.get_variable("g_c", [501], lambda r, s: np.full(s, float(rank), dtype=np.float32))
    st.finalize(OrderedDict(d=[a, b], g=[c]))
    opt_d = ops.AdamOptimizer(st, "d")
    opt_g = ops.AdamOptimizer(st, "g")
    a.grad.fill_(1.0 + rank); b.grad.fill_(10.0 * (rank + 1)); c.grad.fill_(100.0)
    dp.bucket_bytes = 1024            # force several buckets inside the 'd' range (1040 floats = 5 buckets)
    assert len(dp.buckets(*opt_d.range())) == 5
    dp.allreduce(opt_d)
    assert torch.all(a.grad == 3.0) and torch.all(b.grad == 30.0)       # sum over ranks (Adam applies 1/world)
    assert torch.all(c.grad == 100.0)                                   # the other group is untouched
    dp.allreduce(opt_g)
    assert torch.all(c.grad == 200.0)
    # overlapped exchange: the hook reduces the finished tail of the range early, allreduce() the remaining head
    a.grad.fill_(1.0 + rank); b.grad.fill_(10.0 * (rank + 1))
    dp.early_bytes = 100
    dp.begin_update(opt_d)
    assert ops.GRAD_READY_HOOK is not None
    dp.grad_ready(b)                                                    # [b.offset, end): 40 floats >= 100 B -> reduced now
    assert torch.all(b.grad == 30.0) and torch.all(a.grad == 1.0 + rank)
    dp.allreduce(opt_d)                                                 # the head [a.offset, b.offset)
    assert torch.all(a.grad == 3.0) and torch.all(b.grad == 30.0)
    assert dp.last_buckets == [(b.offset, opt_d.range()[1]), (a.offset, b.offset)] and ops.GRAD_READY_HOOK is None
    dp.begin_update(opt_d)
    dp.grad_ready(a)
    try:
        dp.grad_ready(b)                                                # a second gradient for an already reduced bucket
        raise AssertionError("late gradient not detected")
    except RuntimeError:
        pass
    dp.allreduce(opt_d)
    # a tuple var_list (VID_DCGAN with --train_img_disc: ("dvideo", "d_img")): the first group of the layout is ALSO the first
    # one backward produces, so a finished tail is a per-group notion -- the early bucket of the first group must not
    # take the second group's (still unwritten) gradients with it, and the later gradients must not raise
    st2 = ops.VariableStore(device="cpu", seed=rank)
    v1 = st2.get_variable("dvideo_a", [64], lambda r, s: np.zeros(s, dtype=np.float32))
    v2 = st2.get_variable("dvideo_b", [64], lambda r, s: np.zeros(s, dtype=np.float32))
    i1 = st2.get_variable("d_img_a", [64], lambda r, s: np.zeros(s, dtype=np.float32))
    i2 = st2.get_variable("d_img_b", [64], lambda r, s: np.zeros(s, dtype=np.float32))
    st2.finalize(OrderedDict(dvideo=[v1, v2], d_img=[i1, i2]))
    opt = ops.AdamOptimizer(st2, ("dvideo", "d_img"))
    dp.early_bytes = 100
    dp.begin_update(opt)
    for var, val in ((v2, 1.0), (v1, 2.0), (i2, 3.0), (i1, 4.0)):          # backward order: video D first, image D last
        var.grad.fill_(val * (rank + 1))
        dp.grad_ready(var)
    dp.allreduce(opt)
    for var, val in ((v2, 1.0), (v1, 2.0), (i2, 3.0), (i1, 4.0)):
        assert torch.all(var.grad == val * 3.0), (var.name, var.grad[:3])   # each reduced exactly once: (1 + 2) * val
    assert sorted(dp.last_buckets) == sorted([(v2.offset, v2.offset + 64), (v1.offset, v2.offset), (i2.offset, i2.offset + 64), (i1.offset, i2.offset)])
    # broadcast: every rank ends with rank 0's variables
    dp.broadcast_parameters(st)
    assert torch.all(c.data == 0.0)
    # timing reduction used by bench.py
    assert dp.max_over_ranks(1.0 + rank) == float(world)
    dp.barrier()
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
        f.write("ok")
    torch.distributed.destroy_process_group()


def test_dataparallel_gloo_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_single_process_is_a_noop():
    for p in (ROOT, os.path.join(ROOT, "gif-gan_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from gifgan.dp import DataParallel
    dp = DataParallel(world_size=1, rank=0, init=False)
    assert dp.shard(64) == (0, 64) and dp.max_over_ranks(2.5) == 2.5
    assert dp.buckets(0, 10) == [(0, 10)]
    dp.barrier()
