#!/usr/bin/env python
"""Runs every conv layer's three kernels (down / up / wgrad) of the DCGAN-64 step once or a few times at the
bench shapes and prints one JSON line per kernel: the command profiled by `ncu --set full` for profiles/.
    python tools/layer_kernels.py [--batch 64] [--reps 1] [--launches 1]"""
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
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--launches", type=int, default=1)
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    from gifgan import ops
    from gifgan.model import DCGAN
    ops.set_precision(a.precision)
    ops.reset_default_store(device="cuda", seed=7)
    model = DCGAN(None, batch_size=a.batch, output_size=64, c_dim=3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = bench.layer_rooflines(model, a.batch, a.precision, flush, reps=a.reps, launches=a.launches)
    for r in sorted(rows, key=lambda r: r["order"]):
        print("ORDER", r["kernel"], flush=True)                  # execution order: maps the ncu launch list back to layers
    for r in rows:
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}), flush=True)
    print("TOTAL share_ms", round(sum(r["share_ms"] for r in rows), 4))


if __name__ == "__main__":
    main()
