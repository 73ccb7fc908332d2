#!/usr/bin/env python
"""Per-layer parity diagnostic: one D update of the full-size DCGAN (batch 64, 64x64x3) on the GPU vs the float64
oracle; prints, for every fused layer, the error of the pre-norm tensor, the activation, the incoming gradient and
the pre-norm gradient (max-norm and L2), then the filter gradients.  Usage: python tools/diag_parity.py fp32|bf16 [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gif-gan_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from gifgan import ops  # noqa: E402
from gifgan.model import DCGAN  # noqa: E402
from oracle.models import DCGAN as OracleDCGAN  # noqa: E402


def errs(got, want):
    got, want = got.detach().float().cpu().double(), want.detach().double()
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).item(), ((got - want).norm() / want.norm().clamp_min(1e-30)).item()


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    size, gf = (64, 64) if B >= 32 else (32, 16)
    ora = OracleDCGAN(batch_size=B, output_size=size, gf_dim=gf, df_dim=gf, seed=7, dtype=torch.float64)
    if len(sys.argv) > 3 and sys.argv[3] == "quant":
        ora.quant = "bf16"       # oracle with the product's bf16 quantisation points
    ops.set_precision(precision)
    ops.reset_default_store(device="cuda")
    m = DCGAN(None, batch_size=B, output_size=size, gf_dim=gf, df_dim=gf)
    m.store.load_state_dict(ora.state_dict())
    img = np.random.RandomState(102).uniform(-1, 1, (B, size, size, 3)).astype(np.float32)
    z = np.random.RandomState(1000).uniform(-1, 1, (B, 100)).astype(np.float32)
    ops.DEBUG_TAP = {}
    losses = m.d_update(torch.tensor(img).cuda(), torch.tensor(z).cuda(), apply=False)
    ora.trace = {}
    want = ora.d_update(torch.tensor(img).double(), torch.tensor(z).double(), apply=False)
    print(f"== {precision} B={B}: d_loss got {losses[0].item():.7f} want {want['d_loss']:.7f}")
    T = ora.trace
    fwd = [f for f in ops.DEBUG_TAP["fwd"] if "/d_h" in "/" + f[0]]
    bwd = {b[0]: b for b in ops.DEBUG_TAP["bwd"]}
    for name, pre, y in fwd:
        layer = name.split("/")[0].replace("_conv", "")          # d_h1
        idx = layer[-1]
        pre_w = torch.cat([T[f"d_real_h{idx}_conv"], T[f"d_fake_h{idx}_conv"]], 0)
        y_w = torch.cat([T[f"d_real_h{idx}"], T[f"d_fake_h{idx}"]], 0)
        line = f"{layer}: pre {errs(pre, pre_w)[0]:.2e}/{errs(pre, pre_w)[1]:.2e}  act {errs(y, y_w)[0]:.2e}/{errs(y, y_w)[1]:.2e}"
        if name in bwd:
            _, dy, dpre = bwd[name]
            dy_w = torch.cat([T[f"d_real_h{idx}"].grad, T[f"d_fake_h{idx}"].grad], 0)
            dpre_w = torch.cat([T[f"d_real_h{idx}_conv"].grad, T[f"d_fake_h{idx}_conv"].grad], 0)
            line += f"  dy {errs(dy, dy_w)[0]:.2e}/{errs(dy, dy_w)[1]:.2e}  dpre {errs(dpre, dpre_w)[0]:.2e}/{errs(dpre, dpre_w)[1]:.2e}"
            # decompose the pre-norm gradient error into a per-channel constant part and the rest
            e = (dpre.detach().float().cpu().double() - dpre_w).reshape(-1, dpre_w.shape[-1])
            half = e.shape[0] // 2
            cm = torch.cat([e[:half].mean(0, keepdim=True).expand(half, -1), e[half:].mean(0, keepdim=True).expand(half, -1)], 0)
            line += f"  | dpre err: per-channel-mean part {cm.norm().item() / e.norm().clamp_min(1e-30).item():.2f} of total"
        print(line)
    for k in [v.name for v in m.d_vars]:
        e = errs(m.store.vars[k].grad, want["grads"][k])
        print(f"grad {k:24s} max-rel {e[0]:.2e}  l2-rel {e[1]:.2e}  |want|max {want['grads'][k].abs().max().item():.2e}")
    ops.DEBUG_TAP = {}
    gl = m.g_update(torch.tensor(z).cuda(), apply=False)
    wg = ora.g_update(torch.tensor(z).double(), apply=False)
    print(f"== g_loss got {gl[0].item():.7f} want {wg['g_loss']:.7f}")
    for k in [v.name for v in m.g_vars]:
        e = errs(m.store.vars[k].grad, wg["grads"][k])
        print(f"grad {k:24s} max-rel {e[0]:.2e}  l2-rel {e[1]:.2e}  |want|max {wg['grads'][k].abs().max().item():.2e}")


if __name__ == "__main__":
    main()
