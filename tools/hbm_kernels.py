#!/usr/bin/env python
"""Runs the bandwidth-bound kernels of the DCGAN-64 step alone at the step's shapes (bench.hbm_kernel_table): the command
profiled by `ncu --set full -k regex:'bn_|colsum|adam|c3m'` for profiles/.
    python tools/hbm_kernels.py [--batch 64]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gif-gan_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    from gifgan import ops
    from gifgan.model import DCGAN
    ops.set_precision("bf16")
    ops.reset_default_store(device="cuda", seed=7)
    model = DCGAN(None, batch_size=a.batch, output_size=64, c_dim=3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = bench.layer_rooflines(model, a.batch, "bf16", flush, reps=1, launches=1)
    for r in bench.hbm_kernel_table(model, a.batch, flush, bench.load_peaks(), [r for r in rows if r.get("bytes")]):
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
