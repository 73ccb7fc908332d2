#!/usr/bin/env python
"""Throughput of the video GAN train step (BASELINE.json config 3: models/recurrent_z, 32 clips x 16 frames of 64x64 RGB,
default flags = image GAN frozen, 1 D update + 2 G updates): one CUDA-graph replay per step, batch resident in HBM,
L2 flushed between steps, CUDA events.  Prints one JSON line.
    python tools/vid_bench.py [--clips 32] [--steps 10] [--train_img]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gif-gan_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--train_img", action="store_true", help="--train_img_gen --train_img_disc variant (SURVEY 8d config 3)")
    a = ap.parse_args()
    from gifgan import ops
    from gifgan.z_model_lib import VID_DCGAN
    ops.set_precision("bf16")
    ops.reset_default_store(device="cuda", seed=7)
    T = 16
    with ops.variable_scope("video_gan"):
        m = VID_DCGAN(None, a.clips, 120, 100, T, 64, 64, 3, train_img_gen=a.train_img, train_img_disc=a.train_img)
    img = torch.from_numpy(np.random.RandomState(103).uniform(-1, 1, (a.clips * T, 64, 64, 3)).astype(np.float32)).pin_memory()
    z = torch.from_numpy(np.random.RandomState(1000).uniform(-1, 1, (a.clips, 120)).astype(np.float32)).pin_memory()
    for _ in range(2):
        m.train_step(img, z, sync=False)
    torch.cuda.synchronize()
    g = m._graphs[(1, 2)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms = 0.0
    for _ in range(a.steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g["graph"].replay()
        e.record()
        e.synchronize()
        ms += s.elapsed_time(e)
    ms /= a.steps
    print(json.dumps({"metric": "GAN train frames/sec", "config": {"workload": "recurrent_z VID_DCGAN, %d clips x 16 frames 64x64x3, 1 D + 2 G updates, image GAN %s"
                                                                   % (a.clips, "trained" if a.train_img else "frozen")},
                      "ms_per_step": round(ms, 4), "value": round(a.clips * T / ms * 1e3, 1), "unit": "frames/s", "clips_per_s": round(a.clips / ms * 1e3, 1),
                      "steps": a.steps, "dtype": "bf16", "gpu_launches_per_step": g["launches"], "l2": "flushed"}), flush=True)


if __name__ == "__main__":
    main()
