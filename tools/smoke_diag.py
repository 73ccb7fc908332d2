import sys
sys.path[:0]=['/root/repo','/root/repo/gif-gan_b200']
import numpy as np, torch
from gifgan import ops
from gifgan.model import DCGAN
from oracle.models import DCGAN as OracleDCGAN
B=8
ora = OracleDCGAN(batch_size=B, output_size=32, gf_dim=16, df_dim=16, seed=7)
ops.set_precision("fp32"); ops.reset_default_store(device="cuda")
m = DCGAN(None, batch_size=B, output_size=32, gf_dim=16, df_dim=16)
m.store.load_state_dict(ora.state_dict())
img = np.random.RandomState(102).uniform(-1, 1, (B, 32, 32, 3)).astype(np.float32)
z = np.random.RandomState(1000).uniform(-1, 1, (B, 100)).astype(np.float32)
got = m.train_step(img, z, use_graph=False)
want = ora.train_step(torch.tensor(img), torch.tensor(z))
print(got, {k: want[k] for k in ("d_loss","g_loss_first","g_loss")})
for k, v in m.store.vars.items():
    if "moving_" in k: continue
    d = (v.data.cpu() - ora.vars[k]).abs()
    far = int((d > 0.05 * 2e-4 * 2).sum())
    if far: print("%-20s far %7d / %7d  max %.2e" % (k, far, d.numel(), d.max().item()))
