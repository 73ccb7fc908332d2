"""GPU parity tests for the two video models on the hot path:
  * models/recurrent_z  VID_DCGAN (z_model_lib.py): latent-RNN video GAN around a frozen image DCGAN, conv3d video
    discriminator -- BASELINE config 3 at a small clip batch;
  * models/recurrent_image recurrent_DCGAN.py: per-frame conv encoder -> BasicLSTMCell -> deconv decoder + frame/clip
    discriminator.
Both run the reference schedule (1 D update + 2 G updates) through the C ABI and are compared with the CPU oracle on
identical weights / latents / synthetic frames, and with the committed float64 golden traces (tests/golden/*.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.models import VID_DCGAN as OracleVID, RecurrentDCGAN as OracleRec  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _vid_pair(precision, Bv=2, T=16, **kw):
    from gifgan import ops
    from gifgan.z_model_lib import VID_DCGAN
    ora = OracleVID(batch_size=Bv, vid_length=T, output_image_size=64, seed=7, dtype=torch.float64, **kw)
    ops.set_precision(precision)
    ops.reset_default_store(device="cuda")
    with ops.variable_scope('video_gan'):                       # z_model.py:63
        m = VID_DCGAN(None, batch_size=Bv, z_input_size=120, z_output_size=100, vid_length=T, input_image_size=64,
                      output_image_size=64, c_dim=3, sample_cols=Bv, **kw)
    assert set(m.store.vars) == set(ora.vars), (sorted(set(m.store.vars) ^ set(ora.vars))[:10])
    m.store.load_state_dict(ora.state_dict())
    return m, ora


def test_vid_dcgan_reference_schedule_fp32_and_golden():
    g = np.load(os.path.join(GOLD, "vid_tiny.npz"))
    m, ora = _vid_pair("fp32")
    img = np.random.RandomState(103).uniform(-1, 1, (32, 64, 64, 3))
    frozen = {k: v.data.clone() for k, v in m.store.vars.items() if "image_gan" in k}
    assert len(frozen) > 30
    for step in range(2):
        z = np.random.RandomState(1000 + step).uniform(-1, 1, (2, 120))
        got = m.train_step(img.astype(np.float32), z.astype(np.float32), use_graph=False)
        want = ora.train_step(torch.tensor(img), torch.tensor(z))
        # first step: pure kernel parity.  Second step: every weight has moved by ~lr*sign(g) three times; a gradient
        # element that is rounding noise around zero steps the other way in fp32 vs float64 (same criterion as
        # tests/test_gpu_dcgan.py::test_reference_schedule_three_steps_fp32, on a 2-clip batch)
        tol = 2e-3 if step == 0 else 3e-2
        for k in ("d_loss", "g_loss"):
            assert abs(got[k] - want[k]) < tol * max(1.0, abs(want[k])), (step, k, got[k], want[k])
        assert abs(got["d_loss"] - g["losses"][step][0]) < tol * max(1.0, abs(g["losses"][step][0]))
        assert abs(got["g_loss"] - g["losses"][step][1]) < tol * max(1.0, abs(g["losses"][step][1]))
    # default flags (z_model.py:44-47): the image GAN is frozen -- weights AND batch-norm EMAs
    for k, v in frozen.items():
        assert torch.equal(m.store.vars[k].data, v), k
    assert m.d_optim.t == 2 and m.g_optim.t == 4


def test_vid_dcgan_gradients_fp32():
    """Per-variable gradients of the video nets after one D and one G backward (1e-4 fp32 tolerance on the max norm,
    3e-3 L2 for the deepest ones: the ReLU-mask argument of test_gpu_dcgan applies here too)."""
    m, ora = _vid_pair("fp32")
    img = np.random.RandomState(103).uniform(-1, 1, (32, 64, 64, 3))
    z = np.random.RandomState(1000).uniform(-1, 1, (2, 120))
    ti, tz = torch.tensor(img, dtype=torch.float32).cuda(), torch.tensor(z, dtype=torch.float32).cuda()
    m.d_update(ti, tz, apply=False)
    want = ora.d_update(torch.tensor(img), torch.tensor(z), apply=False)
    for k, gref in want["grads"].items():
        got = m.store.vars[k].grad.detach().cpu().double()
        if gref.abs().max() < 1e-12:
            continue
        err = ((got - gref).norm() / gref.norm()).item()
        assert err < 3e-3, (k, err)
    m.g_update(tz, apply=False)
    wg = ora.g_update(torch.tensor(z), apply=False)
    for k, gref in wg["grads"].items():
        got = m.store.vars[k].grad.detach().cpu().double()
        if gref.abs().max() < 1e-12 or k.endswith("/bias") and "gvideo_3" not in k:
            continue                                           # biases in front of a train-mode batch norm: exact zero
        err = ((got - gref).norm() / gref.norm()).item()
        assert err < 3e-3, (k, err)


def test_vid_dcgan_bf16_graph_step_runs_and_tracks_fp32():
    """bf16 mode through the CUDA graph: losses within 2e-2-class tolerance of the oracle on the first step."""
    m, ora = _vid_pair("bf16")
    img = np.random.RandomState(103).uniform(-1, 1, (32, 64, 64, 3))
    z = np.random.RandomState(1000).uniform(-1, 1, (2, 120))
    got = m.train_step(img.astype(np.float32), z.astype(np.float32), use_graph=True)
    want = ora.train_step(torch.tensor(img), torch.tensor(z))
    for k in ("d_loss", "g_loss"):
        assert abs(got[k] - want[k]) < 5e-2 * max(1.0, abs(want[k])), (k, got[k], want[k])
    got2 = m.train_step(img.astype(np.float32), z.astype(np.float32), use_graph=True)     # replay
    assert np.isfinite(got2["d_loss"]) and np.isfinite(got2["g_loss"])


def test_recurrent_dcgan_reference_schedule_fp32_and_golden():
    from gifgan import ops
    from gifgan.recurrent_dcgan import RecurrentDCGAN
    r = np.load(os.path.join(GOLD, "recurrent_tiny.npz"))
    ora = OracleRec(batch_size=2, video_length=3, seed=7, dtype=torch.float64)
    ops.set_precision("fp32")
    ops.reset_default_store(device="cuda")
    m = RecurrentDCGAN(batch_size=2, video_length=3)
    assert set(m.store.vars) == set(ora.vars), (sorted(set(m.store.vars) ^ set(ora.vars))[:10])
    m.store.load_state_dict(ora.state_dict())
    inp = np.random.RandomState(104).randint(0, 256, (2, 4, 64, 64, 3)).astype(np.int32)
    for step in range(2):
        got = m.train_step(torch.tensor(inp))
        # the schedule of recurrent_DCGAN.py:353-375: d_optim, g_optim, g_optim; d_loss is the D update's, g_loss the last G update's
        wd = ora.update(torch.tensor(inp), "d")
        ora.update(torch.tensor(inp), "g")
        wg = ora.update(torch.tensor(inp), "g")
        tol = 2e-3 if step == 0 else 1e-2
        assert abs(got["d_loss"] - wd["d_loss"]) < tol * max(1.0, abs(wd["d_loss"])), (step, got, wd["d_loss"])
        assert abs(got["g_loss"] - wg["g_loss"]) < tol * max(1.0, abs(wg["g_loss"])), (step, got, wg["g_loss"])
        assert abs(got["g_loss"] - r["losses"][step][1]) < tol * max(1.0, abs(r["losses"][step][1]))      # committed golden trace
    k = "generator/lstm/Bias"
    d = (m.store.vars[k].data.cpu().double() - ora.vars[k]).abs()
    assert (d > 0.05 * 2e-4 * 4).double().mean().item() < 0.05, "LSTM bias drifted from the oracle beyond Adam noise"
