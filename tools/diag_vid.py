"""Config-3 full-size gradient errors (L2, vs float64 oracle) of the video nets; env switches A/B the fused paths."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gif-gan_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from oracle.models import VID_DCGAN as OracleVID
from gifgan import ops
from gifgan.z_model_lib import VID_DCGAN
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
Bv, T = int(sys.argv[2]) if len(sys.argv) > 2 else 32, 16
ora = OracleVID(batch_size=Bv, vid_length=T, output_image_size=64, seed=7, dtype=torch.float64)
ops.set_precision(prec); ops.reset_default_store(device="cuda")
with ops.variable_scope('video_gan'):
    m = VID_DCGAN(None, batch_size=Bv, z_input_size=120, z_output_size=100, vid_length=T, input_image_size=64, output_image_size=64, c_dim=3, sample_cols=8)
m.store.load_state_dict(ora.state_dict())
img = np.random.RandomState(103).uniform(-1, 1, (Bv * T, 64, 64, 3)); z = np.random.RandomState(1000).uniform(-1, 1, (Bv, 120))
ti, tz = torch.tensor(img, dtype=torch.float32).cuda(), torch.tensor(z, dtype=torch.float32).cuda()
got = m.d_update(ti, tz, apply=False); want = ora.d_update(torch.tensor(img), torch.tensor(z), apply=False)
print("d_loss", float(got["losses"][0]), want["d_loss"])
for k, g in want["grads"].items():
    if g.abs().max() > 1e-12: print("  D %-55s %.4f" % (k, float((m.store.vars[k].grad.cpu().double() - g).norm() / g.norm())))
gg = m.g_update(tz, apply=False); wg = ora.g_update(torch.tensor(z), apply=False)
print("g_loss", float(gg["losses"][0]), wg["g_loss"])
for k, g in wg["grads"].items():
    if g.abs().max() > 1e-12: print("  G %-55s %.4f" % (k, float((m.store.vars[k].grad.cpu().double() - g).norm() / g.norm())))
