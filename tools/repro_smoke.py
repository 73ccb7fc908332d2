import sys, os, gc
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/gif-gan_b200")
import numpy as np, torch
from gifgan import ops
from gifgan.model import DCGAN
variant = sys.argv[1]
B = 8
def fp32_part(keep):
    ops.set_precision("fp32"); ops.reset_default_store(device="cuda")
    m0 = DCGAN(None, batch_size=B, output_size=32, gf_dim=16, df_dim=16)
    img = np.random.RandomState(102).uniform(-1, 1, (B, 32, 32, 3)).astype(np.float32)
    z = np.random.RandomState(1000).uniform(-1, 1, (B, 100)).astype(np.float32)
    l = m0.d_update(torch.tensor(img).cuda(), torch.tensor(z).cuda(), apply=False)
    m0.train_step(img, z, use_graph=False)
    return (m0, l) if keep else None
held = None
if "fp32keep" in variant: held = fp32_part(True)
if "fp32drop" in variant: fp32_part(False); gc.collect()
ops.set_precision("bf16", tensor_cores=True); ops.reset_default_store(device="cuda")
m = DCGAN(None, batch_size=B, output_size=64, gf_dim=64, df_dim=64)
img = np.random.RandomState(102).uniform(-1, 1, (B, 64, 64, 3)).astype(np.float32)
z = np.random.RandomState(1000).uniform(-1, 1, (B, 100)).astype(np.float32)
keep = []
if "upd" in variant:
    keep.append(m.d_update(torch.tensor(img).cuda(), torch.tensor(z).cuda(), apply=False))
    if "sync" in variant: float(keep[-1][0])
    keep.append(m.g_update(torch.tensor(z).cuda(), apply=False))
    if "sync" in variant: float(keep[-1][0])
if "drop" in variant.split("_")[-1:]: keep = []
try:
    print(variant, m.train_step(img, z, use_graph=True))
except Exception as e:
    print(variant, "FAILED", repr(e)[:120])
