"""Generates tests/golden/*.npz from the CPU oracle (float64).

The reference cannot run here (TensorFlow 0.12 / Python 2.7) and ships no
vectors of its own, so these fixtures are ORACLE outputs: they pin the oracle
against regressions and travel to the GPU box as parity vectors for the CUDA path.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tf_ops as T  # noqa: E402
from oracle.latent import LatentSearch, make_trained_like  # noqa: E402
from oracle.models import DCGAN, VID_DCGAN, RecurrentDCGAN  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
f64 = torch.float64


def t(a):
    return torch.tensor(a, dtype=f64)


def ops_fixture():
    rs = np.random.RandomState(11)
    d = {}
    x, w, b = rs.randn(2, 8, 8, 3), rs.randn(5, 5, 3, 8) * 0.1, rs.randn(8) * 0.1
    d["conv_x"], d["conv_w"], d["conv_b"] = x, w, b
    d["conv_y"] = T.conv2d(t(x), t(w), t(b)).numpy()
    x, w, b = rs.randn(2, 4, 4, 8), rs.randn(5, 5, 6, 8) * 0.1, rs.randn(6) * 0.1
    d["deconv_x"], d["deconv_w"], d["deconv_b"] = x, w, b
    d["deconv_y"] = T.conv2d_transpose(t(x), t(w), [2, 8, 8, 6], t(b)).numpy()
    x, w, b = rs.randn(2, 4, 4, 4, 8), rs.randn(3, 3, 3, 8, 8) * 0.1, rs.randn(8) * 0.1
    d["conv3d_x"], d["conv3d_w"], d["conv3d_b"] = x, w, b
    d["conv3d_y"] = T.conv3d(t(x), t(w), t(b)).numpy()
    x, g, be = rs.randn(4, 4, 4, 8), rs.rand(8) + 0.5, rs.randn(8) * 0.1
    y, mm, mv = T.batch_norm_train(t(x), t(g), t(be), torch.zeros(8, dtype=f64), torch.ones(8, dtype=f64))
    d["bn_x"], d["bn_gamma"], d["bn_beta"] = x, g, be
    d["bn_y"], d["bn_mm"], d["bn_mv"] = y.numpy(), mm.numpy(), mv.numpy()
    lg = rs.randn(16, 1) * 3
    d["ce_logits"] = lg
    d["ce_ones"] = T.sigmoid_cross_entropy_with_logits(t(lg), torch.ones(16, 1, dtype=f64)).numpy()
    d["ce_zeros"] = T.sigmoid_cross_entropy_with_logits(t(lg), torch.zeros(16, 1, dtype=f64)).numpy()
    p, gr = rs.randn(32), rs.randn(32) * 1e-2
    pt = t(p).clone()
    opt = T.TFAdam({"p": pt})
    for _ in range(3):
        opt.apply({"p": t(gr)})
    d["adam_p0"], d["adam_g"], d["adam_p3"] = p, gr, pt.numpy()
    xx, c, h = rs.randn(3, 12), rs.randn(3, 5), rs.randn(3, 5)
    M, bi = rs.randn(17, 20) * 0.3, rs.randn(20) * 0.1
    nc, nh = T.basic_lstm_cell(t(xx), t(c), t(h), t(M), t(bi))
    d.update(lstm_x=xx, lstm_c=c, lstm_h=h, lstm_M=M, lstm_b=bi, lstm_nc=nc.numpy(), lstm_nh=nh.numpy())
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **d)


def dcgan_fixture():
    """Tiny DCGAN (batch 4, 16x16, gf=df=8): 3 reference-schedule steps in float64."""
    m = DCGAN(batch_size=4, output_size=16, gf_dim=8, df_dim=8, seed=7, dtype=f64)
    init = {k: v.numpy().copy() for k, v in m.state_dict().items()}
    img = np.random.RandomState(102).uniform(-1, 1, (4, 16, 16, 3))
    d = {"images": img}
    losses = []
    for step in range(3):
        z = np.random.RandomState(1000 + step).uniform(-1, 1, (4, 100))
        o = m.train_step(t(img), t(z))
        losses.append([o["d_loss"], o["g_loss_first"], o["g_loss"]])
    d["losses"] = np.array(losses)
    for k, v in init.items():
        d["init/" + k] = v
    for k, v in m.state_dict().items():
        d["final/" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "dcgan_tiny.npz"), **d)


def vid_fixture():
    m = VID_DCGAN(batch_size=2, vid_length=16, output_image_size=64, seed=7, dtype=f64)
    # shrink nothing: the image GAN must be 64x64 so that h2 is [*,8,8,256]; only record scalars + gvideo weights
    img = np.random.RandomState(103).uniform(-1, 1, (32, 64, 64, 3))
    losses = []
    for step in range(2):
        z = np.random.RandomState(1000 + step).uniform(-1, 1, (2, 120))
        o = m.train_step(t(img), t(z))
        losses.append([o["d_loss"], o["g_loss"]])
    d = {"losses": np.array(losses)}
    for k, v in m.state_dict().items():
        if "gvideo_3" in k or "dvideo_h4" in k or "dvideo_bn3" in k:
            d["final/" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "vid_tiny.npz"), **d)


def recurrent_fixture():
    m = RecurrentDCGAN(batch_size=2, video_length=3, seed=7, dtype=f64)
    inp = np.random.RandomState(104).randint(0, 256, (2, 4, 64, 64, 3)).astype(np.int32)
    losses = []
    for _ in range(2):
        o = m.train_step(torch.tensor(inp))
        losses.append([o["d_loss"], o["g_loss"]])
    d = {"losses": np.array(losses), "final/generator/lstm/Bias": m.vars["generator/lstm/Bias"].numpy(),
         "final/discriminator/d_final_fc_w": m.vars["discriminator/d_final_fc_w"].numpy()}
    np.savez_compressed(os.path.join(OUT, "recurrent_tiny.npz"), **d)


LATENT_WEIGHTS = dict(pixel_L2_weight=0.3, pixel_L1_weight=0.1, activations_L2_weight=0.3, activations_L1_weight=0.2,
                      generator_loss_weight=0.1)


def latent_fixture():
    """Latent search on a tiny trained-like DCGAN (batch 4, 16x16, gf=df=8), all five loss terms, both modes:
    gradient at the initial latents, 4 steps at lr 0.05 (losses, final latents and images)."""
    tgt = np.random.RandomState(105).uniform(-1, 1, (4, 16, 16, 3))
    d = {"targets": tgt}
    for mode in ("train", "inference"):
        m = make_trained_like(DCGAN(batch_size=4, output_size=16, gf_dim=8, df_dim=8, seed=7, dtype=f64))
        s = LatentSearch(m, mode, random_seed=3, **LATENT_WEIGHTS)
        d[f"{mode}/z0"] = s.z.numpy().copy()
        if mode == "train":        # identical initial state in both modes (train mode never reads the moving averages)
            for k, v in m.state_dict().items():
                d["weights/" + k] = v.numpy()
        acts = s.target_activations(tgt)
        d[f"{mode}/target_activations"] = acts.numpy()
        loss0, g0 = s.loss_and_grad(tgt, acts)
        d[f"{mode}/loss0"], d[f"{mode}/grad0"] = np.array(loss0), g0.numpy()
        d[f"{mode}/losses"] = np.array([s.step(tgt, acts, 0.05) for _ in range(4)])
        d[f"{mode}/z4"], d[f"{mode}/images4"] = s.z.numpy().copy(), s.images().numpy()
    np.savez_compressed(os.path.join(OUT, "latent_tiny.npz"), **d)


RECURRENT_VARIANTS = {"multi": dict(num_layers=3), "shared_dropout": dict(num_layers=3, shared_conv=True, output_keep_prob=0.8)}


def recurrent_variant_masks():
    """The DropoutWrapper draws used by the fixture and the parity tests: [layers, T, B, H] of {0, 1/0.8}."""
    return np.floor(0.8 + np.random.RandomState(106).uniform(size=(3, 3, 2, 100))) / 0.8


def recurrent_variants_fixture():
    """multi-layer_recurrent_DCGAN.py and ..._with_shared_conv_and_drop_out.py at batch 2, 3 frames: one train step."""
    inp = np.random.RandomState(104).randint(0, 256, (2, 4, 64, 64, 3)).astype(np.int32)
    d = {}
    for tag, kw in RECURRENT_VARIANTS.items():
        m = RecurrentDCGAN(batch_size=2, video_length=3, seed=7, dtype=f64, **kw)
        m.masks = torch.tensor(recurrent_variant_masks())
        o = m.train_step(torch.tensor(inp))
        d[tag + "/losses"] = np.array([o["d_loss"], o["g_loss"]])
        d[tag + "/final/generator/lstm/Cell1/Bias"] = m.vars["generator/lstm/Cell1/Bias"].numpy()
        d[tag + "/final/discriminator/d_fc_bias"] = m.vars["discriminator/d_fc_bias"].numpy()
    np.savez_compressed(os.path.join(OUT, "recurrent_variants.npz"), **d)


def hundred_step_fixture():
    """100-step loss trace (north_star: "matching loss over 100 steps") of the tiny DCGAN (batch 8, 16x16, gf=df=8) in
    float64, fresh seeded images and z every step: [100, 2] = (d_loss, g_loss of the second G update)."""
    m = DCGAN(batch_size=8, output_size=16, gf_dim=8, df_dim=8, seed=7, dtype=f64)
    out = []
    for s in range(100):
        img = np.random.RandomState(102 + s).uniform(-1, 1, (8, 16, 16, 3))
        z = np.random.RandomState(1000 + s).uniform(-1, 1, (8, 100))
        o = m.train_step(t(img), t(z))
        out.append([o["d_loss"], o["g_loss"]])
    np.savez_compressed(os.path.join(OUT, "dcgan_100steps.npz"), losses=np.array(out))


if __name__ == "__main__":
    ops_fixture()
    hundred_step_fixture()
    recurrent_variants_fixture()
    latent_fixture()
    dcgan_fixture()
    vid_fixture()
    recurrent_fixture()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
