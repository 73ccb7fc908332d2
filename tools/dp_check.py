#!/usr/bin/env python
"""Data-parallel consistency on real GPUs: N ranks fed the SAME batch must follow the single-GPU trajectory (identical gradients
averaged over ranks are the gradients), for the image DCGAN and the video GAN, through the captured step with the early buckets,
the bf16 / fp32 gradient buckets and the overlapped update tail.
    python tools/dp_check.py single            # writes gpurun_out/dp_single.json
    torchrun --nproc-per-node 2 tools/dp_check.py dp    # compares with it"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gif-gan_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def run(dp):
    from gifgan import ops
    from gifgan.model import DCGAN
    from gifgan.z_model_lib import VID_DCGAN
    out = {}
    ops.set_precision("bf16")
    ops.reset_default_store(device="cuda", seed=7)
    B = 16
    m = DCGAN(None, batch_size=B, output_size=64, c_dim=3, dp=dp)
    if dp:
        dp.broadcast_parameters(m.store)
    img = np.random.RandomState(102).uniform(-1, 1, (B, 64, 64, 3)).astype(np.float32)
    tr = []
    for s in range(4):
        z = np.random.RandomState(1000 + s).uniform(-1, 1, (B, 100)).astype(np.float32)
        o = m.train_step(img, z, use_graph=True)
        tr.append([o["d_loss"], o["g_loss_first"], o["g_loss"]])
    out["dcgan"] = tr
    ops.reset_default_store(device="cuda", seed=7)
    with ops.variable_scope("video_gan"):
        v = VID_DCGAN(None, 4, 120, 100, 16, 64, 64, 3, sample_cols=4, dp=dp)
    if dp:
        dp.broadcast_parameters(v.store)
    img = np.random.RandomState(103).uniform(-1, 1, (64, 64, 64, 3)).astype(np.float32)
    tr = []
    for s in range(3):
        z = np.random.RandomState(1000 + s).uniform(-1, 1, (4, 120)).astype(np.float32)
        o = v.train_step(img, z)
        tr.append([o["d_loss"], o["g_loss"]])
    out["vid"] = tr
    return out


def main():
    mode = sys.argv[1]
    path = os.path.join(ROOT, "gpurun_out", "dp_single.json")
    if mode == "single":
        torch.cuda.set_device(0)
        r = run(None)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        json.dump(r, open(path, "w"))
        print("single", r)
        return
    from gifgan.dp import DataParallel
    dp = DataParallel()
    torch.cuda.set_device(dp.local_rank)
    r = run(dp)
    if dp.rank == 0:
        ref = json.load(open(path))
        worst = first = 0.0
        for k in ref:
            a, b = np.array(ref[k]), np.array(r[k])
            rel = np.abs(a - b) / np.maximum(1.0, np.abs(a))
            worst = max(worst, float(rel.max()))
            first = max(first, float(rel[0].max()))
        print("dp world", dp.world_size, "grad dtype", dp.grad_dtype, "overlap", dp.overlap_update, "p2p", bool(any(dp._p2p.values())),
              "worst rel loss difference vs single GPU: %.3e (first step %.3e)" % (worst, first))
        print("  dcgan", r["dcgan"][-1], "ref", ref["dcgan"][-1], "| vid", r["vid"][-1], "ref", ref["vid"][-1])
        # N ranks fed the same batch average N identical gradients: the FIRST step (one D update and two G updates from identical
        # weights) must agree closely -- fp32 exchange: summation order only; bf16 buckets: 0.4 % rounding of the exchanged
        # gradients.  Later steps only loosely: two single-GPU runs already drift apart by a few % within 3-4 steps (the
        # filter-gradient kernels add their pixel splits in arrival order; profiles/r03b_pack_ab_run_to_run.log).
        assert first < (2e-2 if dp.grad_dtype == "bf16" else 5e-3), first
        assert worst < 0.15, worst
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
