"""Oracle model graphs: parameter inventory, reference-schedule
semantics (App. A.4/A.6) and regression against tests/golden/*.npz."""
import os

import numpy as np
import torch

from oracle.models import DCGAN, VID_DCGAN, RecurrentDCGAN

GOLD = os.path.join(os.path.dirname(__file__), "golden")
f64 = torch.float64


def _t(a):
    return torch.tensor(a, dtype=f64)


def test_parameter_counts_match_survey():
    m = DCGAN(batch_size=2)
    assert sum(m.vars[k].numel() for k in m.d_vars) == 4316545
    assert sum(m.vars[k].numel() for k in m.g_vars) == 5135363
    v = VID_DCGAN(batch_size=1)
    assert sum(v.vars[k].numel() for k in v.d_vid_vars) == 5310721
    assert sum(v.vars[k].numel() for k in v.g_vid_vars) == 642148
    assert not any("gvideo" in k or "dvideo" in k for k in v.img_dcgan.d_vars + v.img_dcgan.g_vars)


def test_dcgan_tiny_matches_golden():
    g = np.load(os.path.join(GOLD, "dcgan_tiny.npz"))
    m = DCGAN(batch_size=4, output_size=16, gf_dim=8, df_dim=8, seed=7, dtype=f64)
    for k in m.vars:
        np.testing.assert_array_equal(m.vars[k].numpy(), g["init/" + k])   # initialisers are deterministic
    losses = []
    for step in range(3):
        z = np.random.RandomState(1000 + step).uniform(-1, 1, (4, 100))
        o = m.train_step(_t(g["images"]), _t(z))
        losses.append([o["d_loss"], o["g_loss_first"], o["g_loss"]])
    np.testing.assert_allclose(np.array(losses), g["losses"], rtol=1e-9)
    for k in m.vars:
        np.testing.assert_allclose(m.vars[k].numpy(), g["final/" + k], rtol=1e-7, atol=1e-10)


def test_reference_schedule_semantics():
    m = DCGAN(batch_size=4, output_size=16, gf_dim=8, df_dim=8, dtype=f64)
    img = _t(np.random.RandomState(1).uniform(-1, 1, (4, 16, 16, 3)))
    z = _t(np.random.RandomState(2).uniform(-1, 1, (4, 100)))
    g0 = m.vars["g_h1/w"].clone()
    d0 = m.vars["d_h1_conv/w"].clone()
    m.d_update(img, z)
    # D update: only d_ variables move; d_bn EMAs advance twice (real, fake), g_bn EMAs once
    assert torch.equal(m.vars["g_h1/w"], g0) and not torch.equal(m.vars["d_h1_conv/w"], d0)
    assert m.d_optim.t == 1 and m.g_optim.t == 0
    d1 = m.vars["d_h1_conv/w"].clone()
    m.g_update(z); m.g_update(z)
    assert torch.equal(m.vars["d_h1_conv/w"], d1) and m.g_optim.t == 2   # G's Adam counter advances twice per batch
    # moving variance started at 1 and was pulled toward batch var 3 (g) / 4 (d: 2 in D-update + 2 G-updates) times
    assert (m.vars["g_bn0/moving_variance"] != 1).all()


def test_inference_bn_does_not_update_ema():
    m = DCGAN(batch_size=2, output_size=16, gf_dim=8, df_dim=8, dtype=f64)
    before = m.vars["g_bn1/moving_mean"].clone()
    m.sampler(_t(np.random.RandomState(0).uniform(-1, 1, (2, 100))))
    assert torch.equal(m.vars["g_bn1/moving_mean"], before)


def test_vid_and_recurrent_match_golden():
    g = np.load(os.path.join(GOLD, "vid_tiny.npz"))
    m = VID_DCGAN(batch_size=2, vid_length=16, output_image_size=64, seed=7, dtype=f64)
    img0 = {k: v.clone() for k, v in m.vars.items() if "image_gan" in k}
    img = np.random.RandomState(103).uniform(-1, 1, (32, 64, 64, 3))
    losses = []
    for step in range(2):
        z = np.random.RandomState(1000 + step).uniform(-1, 1, (2, 120))
        o = m.train_step(_t(img), _t(z))
        losses.append([o["d_loss"], o["g_loss"]])
    np.testing.assert_allclose(np.array(losses), g["losses"], rtol=1e-8)
    # default flags freeze the image GAN entirely (weights AND EMAs: inference-mode BN only)
    for k, v in img0.items():
        assert torch.equal(m.vars[k], v), k
    r = np.load(os.path.join(GOLD, "recurrent_tiny.npz"))
    rm = RecurrentDCGAN(batch_size=2, video_length=3, seed=7, dtype=f64)
    inp = np.random.RandomState(104).randint(0, 256, (2, 4, 64, 64, 3)).astype(np.int32)
    losses = [[(o := rm.train_step(torch.tensor(inp)))["d_loss"], o["g_loss"]] for _ in range(2)]
    np.testing.assert_allclose(np.array(losses), r["losses"], rtol=1e-8)


def test_ops_golden_roundtrip():
    from oracle import tf_ops as T
    g = np.load(os.path.join(GOLD, "ops.npz"))
    np.testing.assert_allclose(T.conv2d(_t(g["conv_x"]), _t(g["conv_w"]), _t(g["conv_b"])).numpy(), g["conv_y"], rtol=1e-12)
    np.testing.assert_allclose(T.conv2d_transpose(_t(g["deconv_x"]), _t(g["deconv_w"]), [2, 8, 8, 6], _t(g["deconv_b"])).numpy(), g["deconv_y"], rtol=1e-12)
    np.testing.assert_allclose(T.conv3d(_t(g["conv3d_x"]), _t(g["conv3d_w"]), _t(g["conv3d_b"])).numpy(), g["conv3d_y"], rtol=1e-12)


def test_hundred_step_divergence_floor():
    """What "matching loss over 100 steps" (north_star) can mean: the same oracle in float32 and float64, and the float64
    oracle with one filter nudged by 1e-7, drift apart by a fraction of a percent of the loss within 100 steps -- the GAN
    amplifies rounding.  The 100-step GPU parity test allows 3 %: five times this floor."""
    def run(dtype, nudge=0.0):
        m = DCGAN(batch_size=8, output_size=16, gf_dim=8, df_dim=8, seed=7, dtype=dtype)
        if nudge:
            with torch.no_grad():
                m.vars["g_h1/w"].add_(nudge)
        out = []
        for s in range(100):
            img = torch.tensor(np.random.RandomState(102 + s).uniform(-1, 1, (8, 16, 16, 3)), dtype=dtype)
            z = torch.tensor(np.random.RandomState(1000 + s).uniform(-1, 1, (8, 100)), dtype=dtype)
            o = m.train_step(img, z)
            out.append((o["d_loss"], o["g_loss"]))
        return np.array(out)
    a, b, c = run(torch.float64), run(torch.float32), run(torch.float64, 1e-7)
    np.testing.assert_allclose(a, np.load(os.path.join(GOLD, "dcgan_100steps.npz"))["losses"], rtol=1e-6)     # committed trace
    rel_f32 = (np.abs(b - a) / np.maximum(1.0, np.abs(a))).max()
    rel_nudge = (np.abs(c - a) / np.maximum(1.0, np.abs(a))).max()
    assert np.abs(b[:3] - a[:3]).max() < 1e-5                       # identical at first ...
    assert 1e-5 < rel_f32 < 2e-2 and 1e-6 < rel_nudge < 2e-2, (rel_f32, rel_nudge)   # ... visibly apart later, yet bounded
