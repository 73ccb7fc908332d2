#!/usr/bin/env python
"""Run-to-run spread of the captured DCGAN-64 train step against the spread between the per-filter and the batched
filter re-pack (ops.PACK_BATCH): the filter-gradient kernels add their partial sums with fp32 reduce-adds in whatever
order CTAs finish, so two runs of the same configuration already differ in the last bits and the GAN amplifies it."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gif-gan_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

from gifgan import ops  # noqa: E402
from gifgan.model import DCGAN  # noqa: E402

B = 8
img = np.random.RandomState(102).uniform(-1, 1, (B, 64, 64, 3)).astype(np.float32)
for batch in (False, False, True, True):
    ops.PACK_BATCH = batch
    ops.set_precision("bf16")
    ops.reset_default_store(device="cuda", seed=7)
    m = DCGAN(None, batch_size=B, output_size=64, c_dim=3)
    rows = []
    for step in range(3):
        z = np.random.RandomState(1000 + step).uniform(-1, 1, (B, 100)).astype(np.float32)
        o = m.train_step(img, z, use_graph=True)
        rows.append("%.7g %.7g %.7g" % (o["d_loss"], o["g_loss_first"], o["g_loss"]))
    print("PACK_BATCH=%d launches=%d | " % (batch, m._graph["launches"]) + " | ".join(rows), flush=True)
