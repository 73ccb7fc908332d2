"""GPU parity tests of the two recurrent_DCGAN variants (multi-layer_recurrent_DCGAN.py and
..._with_shared_conv_and_drop_out.py) against the oracle and the committed golden trace, and north_star's 100-step loss
comparison of the image GAN.  First run on a B200 in round 2 (gpurun_out/r2b_*, tools/diag_recurrent.py).

Tolerances.  fp32 kernels against the float64 oracle agree to ~2e-6 (max-norm, every variable) on the base model and on the
shared-encoder variant.  On the 3-layer variant ONE LeakyReLU mask of the discriminator's first layer flips on the generated
frames (a channel with variance 2.7e-4, i.e. |mean|/std ~ 60: an element within 4e-6 of zero after normalisation lands on
the other side in fp32) -- tools/diag_recurrent.py recomputes that batch-norm backward in float64 from the kernel's own
inputs and sees one O(1) element, every other layer at 1e-7 -- which moves d_conv_f1's gradient by 3.5e-3 and, through
d(fake), every generator gradient by ~1e-3 (dense).  The oracle shows the same class of event between its own float32 and
float64 runs on the shared-encoder variant (up to 2.3e-2 on generator/deconv_f1).  Hence: max-norm 1e-2 and L2 3e-3 on
every variable here; the 1e-4 class is asserted where no mask flips (tests/test_gpu_video.py base model, the op tests)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.models import RecurrentDCGAN as OracleRec  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("tag", ["multi", "shared_dropout"])
def test_variant_reference_schedule_fp32_and_golden(tag):
    from make_golden import RECURRENT_VARIANTS, recurrent_variant_masks
    from gifgan import ops
    from gifgan.recurrent_dcgan import RecurrentDCGAN
    kw = RECURRENT_VARIANTS[tag]
    g = np.load(os.path.join(GOLD, "recurrent_variants.npz"))
    ora = OracleRec(batch_size=2, video_length=3, seed=7, dtype=torch.float64, **kw)
    ora.masks = torch.tensor(recurrent_variant_masks())
    ops.set_precision("fp32")
    ops.reset_default_store(device="cuda")
    m = RecurrentDCGAN(batch_size=2, video_length=3, **kw)
    m.masks = recurrent_variant_masks()
    assert set(m.store.vars) == set(ora.vars)
    m.store.load_state_dict(ora.state_dict())
    inp = np.random.RandomState(104).randint(0, 256, (2, 4, 64, 64, 3)).astype(np.int32)
    # gradients of the first D update and the first G update, before any Adam step: EVERY variable of the var_list
    for which in ("d", "g"):
        want = ora.update(torch.tensor(inp), which, apply=False)
        got = m.update(torch.tensor(inp), which, apply=False)
        key = which + "_loss"
        assert abs(float(got[key].detach()) - want[key]) < 1e-5 * max(1.0, abs(want[key])), (which, float(got[key]), want[key])
        for k, w in want["grads"].items():
            g_ = m.store.vars[k].grad.cpu().double()
            assert ((g_ - w).abs().max() / w.abs().max()).item() < 1e-2, (k, "max-norm")
            assert ((g_ - w).norm() / w.norm()).item() < 3e-3, (k, "L2")
    # one full step of the schedule (d_optim, g_optim, g_optim): d_loss is the D update's, g_loss the last G update's;
    # the golden trace holds the oracle's losses of the LAST update, so only its g_loss is the same quantity
    got = m.train_step(torch.tensor(inp))
    wd = ora.update(torch.tensor(inp), "d")
    ora.update(torch.tensor(inp), "g")
    wg = ora.update(torch.tensor(inp), "g")
    assert abs(got["d_loss"] - wd["d_loss"]) < 2e-3 * max(1.0, abs(wd["d_loss"])), (got, wd["d_loss"])
    assert abs(got["g_loss"] - wg["g_loss"]) < 1e-2 * max(1.0, abs(wg["g_loss"])), (got, wg["g_loss"])
    assert abs(got["g_loss"] - g[tag + "/losses"][1]) < 1e-2 * max(1.0, abs(g[tag + "/losses"][1]))
    k = "generator/lstm/Cell1/Bias"
    d = (m.store.vars[k].data.cpu().double() - torch.tensor(g[tag + "/final/" + k])).abs()
    assert d.max().item() <= 2.2 * 2e-4 * 2


def test_shared_encoder_bf16_step_runs():
    from gifgan import ops
    from gifgan.recurrent_dcgan import RecurrentDCGAN
    ops.set_precision("bf16")
    ops.reset_default_store(device="cuda", seed=3)
    m = RecurrentDCGAN(batch_size=4, video_length=4, num_layers=3, shared_conv=True, output_keep_prob=0.8)
    inp = np.random.RandomState(5).randint(0, 256, (4, 5, 64, 64, 3)).astype(np.int32)
    w0 = m.store.vars["discriminator/d_conv_f2"].data.clone()
    out = m.train_step(torch.tensor(inp))
    assert np.isfinite(out["d_loss"]) and np.isfinite(out["g_loss"])
    assert not torch.equal(m.store.vars["discriminator/d_conv_f2"].data, w0)


def test_losses_follow_the_oracle_over_100_steps_fp32():
    """north_star: "matching loss over 100 steps".  A GAN amplifies rounding differences (tests/test_oracle_models.py::
    test_hundred_step_divergence_floor: float32 vs float64 oracle, or a 1e-7 weight perturbation, drift apart by up to
    0.6 % of the loss over 100 steps at this configuration), so the criterion is 3 % on every one of the 100 steps."""
    from gifgan import ops
    from gifgan.model import DCGAN
    from oracle.models import DCGAN as OracleDCGAN
    B, size, w = 8, 16, 8
    ora = OracleDCGAN(batch_size=B, output_size=size, gf_dim=w, df_dim=w, seed=7, dtype=torch.float32)
    ops.set_precision("fp32")
    ops.reset_default_store(device="cuda")
    m = DCGAN(None, batch_size=B, output_size=size, gf_dim=w, df_dim=w)
    m.store.load_state_dict(ora.state_dict())
    trace = np.load(os.path.join(GOLD, "dcgan_100steps.npz"))["losses"]          # float64 oracle, committed
    worst = 0.0
    for s in range(100):
        img = np.random.RandomState(102 + s).uniform(-1, 1, (B, size, size, 3)).astype(np.float32)
        z = np.random.RandomState(1000 + s).uniform(-1, 1, (B, 100)).astype(np.float32)
        got = m.train_step(img, z, use_graph=True)
        want = ora.train_step(torch.tensor(img), torch.tensor(z))
        for k in ("d_loss", "g_loss"):
            worst = max(worst, abs(got[k] - want[k]) / max(1.0, abs(want[k])))
            assert abs(got[k] - want[k]) < 3e-2 * max(1.0, abs(want[k])), (s, k, got[k], want[k])
        assert abs(got["d_loss"] - trace[s, 0]) < 3e-2 * max(1.0, abs(trace[s, 0])) and abs(got["g_loss"] - trace[s, 1]) < 3e-2 * max(1.0, abs(trace[s, 1])), s
    print("worst relative loss difference over 100 steps: %.2e" % worst)
