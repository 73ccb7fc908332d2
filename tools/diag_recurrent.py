"""Per-variable gradient error of RecurrentDCGAN (fp32 mode) against the float64 oracle, for the D and the G update.
Usage: python tools/diag_recurrent.py [base|multi|shared_dropout]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gif-gan_b200"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from make_golden import RECURRENT_VARIANTS, recurrent_variant_masks  # noqa: E402
from oracle.models import RecurrentDCGAN as OracleRec  # noqa: E402
from gifgan import ops  # noqa: E402
from gifgan.recurrent_dcgan import RecurrentDCGAN  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "multi"
kw = RECURRENT_VARIANTS.get(tag, {})
ora = OracleRec(batch_size=2, video_length=3, seed=7, dtype=torch.float64, **kw)
ora.masks = torch.tensor(recurrent_variant_masks())
ops.set_precision("fp32")
ops.reset_default_store(device="cuda")
m = RecurrentDCGAN(batch_size=2, video_length=3, **kw)
m.masks = recurrent_variant_masks()
m.store.load_state_dict(ora.state_dict())
inp = np.random.RandomState(104).randint(0, 256, (2, 4, 64, 64, 3)).astype(np.int32)
for which in ("d", "g"):
    w = ora.update(torch.tensor(inp), which, apply=False)
    g = m.update(torch.tensor(inp), which, apply=False)
    key = which + "_loss"
    print(tag, which, "loss got %.8f want %.8f" % (float(g[key]), w[key]))
    rows = []
    for k, want in w["grads"].items():
        got = m.store.vars[k].grad.cpu().double()
        den = want.abs().max().item()
        rows.append(((got - want).abs().max().item() / max(den, 1e-30), k, den))
    for e, k, den in sorted(rows, reverse=True):
        print("  %-40s relmax %.3e  (max|want| %.3e)" % (k, e, den))

# ---- where does a discrepancy come from?  (1) generator output vs oracle, (2) run-to-run spread of the gradients
with torch.no_grad():
    X, Y = m._split(torch.tensor(inp))
    fake = m.generator(X).cpu().double()
    Xo, Yo = ora._split(torch.tensor(inp))
    want = torch.cat(ora.generator(Xo), 0)
    print("generator output: max|got-want| %.3e  (max|want| %.3e, std %.3e)" % ((fake - want).abs().max().item(), want.abs().max().item(), want.std().item()))
for which in ("d", "g"):
    m.update(torch.tensor(inp), which, apply=False)
    a = {k: v.grad.clone() for k, v in m.store.vars.items() if v.grad is not None}
    m.update(torch.tensor(inp), which, apply=False)
    worst = max(((v.grad - a[k]).abs().max().item() / max(a[k].abs().max().item(), 1e-30), k) for k, v in m.store.vars.items() if k in a)
    print("run-to-run", which, "worst relmax %.3e at %s" % worst)

# (3) self-consistency of every fused conv+BN node of the D update on its OWN inputs: recompute the batch-norm backward
# in float64 from the tapped (pre, dy) and compare with the kernel's dpre; same for the statistics of the forward
ops.DEBUG_TAP = {}
m.update(torch.tensor(inp), "d", apply=False)
tap = ops.DEBUG_TAP
ops.DEBUG_TAP = None
fw = tap.get("fwd", [])
bw = tap.get("bwd", [])
print("fwd taps", len(fw), "bwd taps", len(bw))
Tn = m.T
for (nb, dy, dpre) in bw:
    # find the forward tap whose output shape matches and whose name matches, latest first (backward order is reverse)
    idx = [i for i, (n, p, y) in enumerate(fw) if n == nb and tuple(p.shape) == tuple(dpre.shape)]
    if not idx:
        continue
    n, pre, y = fw.pop(idx[-1])
    C = pre.shape[-1]
    p64 = pre.double().reshape(Tn, -1, C)
    mu = p64.mean(1, keepdim=True); var = p64.var(1, unbiased=False, keepdim=True)
    rstd = (var + 1e-5).rsqrt()
    xh = (p64 - mu) * rstd
    y64 = y.double().reshape(Tn, -1, C)
    slope = 0.2 if (y64 < 0).any() else 0.0
    g = dy.double().reshape(Tn, -1, C) * torch.where(xh > 0, 1.0, slope)
    ref = rstd * (g - g.mean(1, keepdim=True) - xh * (g * xh).mean(1, keepdim=True))
    got = dpre.double().reshape(Tn, -1, C)
    yref = torch.where(xh > 0, xh, slope * xh)
    print("  %-28s %-22s y relmax %.2e  dpre relmax %.2e  |dpre|/|g| %.2e  min var %.2e" % (
        n, tuple(pre.shape), ((y64 - yref).abs().max() / yref.abs().max()).item(), ((got - ref).abs().max() / ref.abs().max()).item(),
        (ref.abs().max() / g.abs().max()).item(), var.min().item()))
