#!/usr/bin/env python
"""Runs the fused thin linear + batch norm + ReLU kernels (csrc/linbn.cu) alone at the generator's shape: the command profiled by
ncu for profiles/.    python tools/linbn_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gif-gan_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from collections import OrderedDict  # noqa: E402
from gifgan import ops  # noqa: E402

ops.FUSE_LINEAR_BN = True
ops.set_precision("bf16")
st = ops.reset_default_store(device="cuda", seed=3)
bn = ops.batch_norm(name="bn")
ops.linear(torch.empty((64, 100), device="meta"), 8192, "l", bn=bn, bn_channels=512, act="relu")
tv = [v for v in st.vars.values() if v.trainable]
st.finalize(OrderedDict(all=tv))
x = torch.rand(64, 100, device="cuda")
dy = torch.randn(64, 8192, device="cuda").to(torch.bfloat16)
for _ in range(3):
    with ops.trainable(tv):
        y = ops.linear(x, 8192, "l", bn=bn, bn_channels=512, act="relu")
        y.backward(dy)
torch.cuda.synchronize()
print("ok", float(y.float().mean()))
