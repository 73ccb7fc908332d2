#!/usr/bin/env python
"""Latent-search throughput: search steps per second of gifgan.latent_search.LatentSearch at the
batch sizes of the two reference programs (z_space_finder.py: 8 clips at once; discriminator_activation_optimizer.py:
8 x 8 grid), bf16, default loss (discriminator-activation L2), eager launches vs one CUDA-graph replay per step, with
the oracle's CPU step timed beside it.  Timed with CUDA events after warm-up; prints one JSON line per configuration.
    python tools/latent_bench.py [--steps 200] [--no-cpu]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gif-gan_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    from gifgan import ops
    from gifgan.latent_search import LatentSearch
    from gifgan.model import DCGAN
    for B in (8, 64):
        for mode in ("inference", "train"):
            row = dict(metric="latent search steps/sec", batch=B, discriminator_mode=mode, dtype="bf16", loss="activations_L2", steps=a.steps)
            for use_graph in (False, True):
                ops.set_precision("bf16")
                ops.reset_default_store(device="cuda", seed=7)
                m = DCGAN(None, batch_size=B, output_size=64, c_dim=3)
                s = LatentSearch(m, mode, random_seed=1, use_graph=use_graph)
                tgt = torch.tensor(np.random.RandomState(2).uniform(-1, 1, (B, 64, 64, 3)).astype(np.float32)).cuda()
                acts = s.target_activations(tgt)
                n0 = ops.cabi.launch_count()
                for _ in range(5):
                    s.step(tgt, acts, 0.05, fetch_loss=False)
                launches = (ops.cabi.launch_count() - n0) / 5 if not use_graph else None
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.steps):
                    s.step(tgt, acts, 0.05, fetch_loss=False)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.steps
                row["graph" if use_graph else "eager"] = dict(ms_per_step=round(ms, 4), steps_per_s=round(1e3 / ms, 1),
                                                              images_per_s=round(B * 1e3 / ms, 1))
                if launches is not None:
                    row["launches_per_step_eager"] = launches
            if not a.no_cpu and B == 8 and mode == "inference":
                from oracle.latent import LatentSearch as OS
                from oracle.models import DCGAN as OD
                torch.set_num_threads(os.cpu_count())
                o = OS(OD(batch_size=B, output_size=64, seed=7), mode, random_seed=1)
                t = tgt.cpu().numpy()
                oa = o.target_activations(t)
                o.step(t, oa, 0.05)
                t0 = time.time()
                for _ in range(3):
                    o.step(t, oa, 0.05)
                row["cpu_baseline"] = dict(kind="port", cores=os.cpu_count(), steps_per_s=round(3 / (time.time() - t0), 2), sample="3 steps, fp32 oracle")
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
