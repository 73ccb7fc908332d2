"""GPU parity tests of the two recurrent_DCGAN variants (multi-layer_recurrent_DCGAN.py and
..._with_shared_conv_and_drop_out.py) against the oracle and the committed golden trace, and the 100-step loss
comparison of the image GAN.

NOT YET RUN ON HARDWARE: the variants were written after round 1's GPU budget was spent (they compose kernels that are
covered elsewhere: the LSTM step, thin / generic linears, the conv stack, batch norm with per-frame groups).  They are
skipped unless GG_UNVALIDATED=1 so that the default suite only holds tests that have passed on a B200; run
    GG_UNVALIDATED=1 python -m pytest tests/test_gpu_zz_unvalidated.py -m gpu
first thing on the next GPU visit and drop the gate."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("GG_UNVALIDATED", "0") == "0",
                                                  reason="not yet validated on a B200 (set GG_UNVALIDATED=1)")]

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
    # gradients of the first D update and the first G update, before any Adam step
    wd = ora.update(torch.tensor(inp), "d", apply=False)
    gd = m.update(torch.tensor(inp), "d", apply=False)
    assert abs(float(gd["d_loss"]) - wd["d_loss"]) < 2e-3 * max(1.0, abs(wd["d_loss"]))
    for k in ("discriminator/d_conv_f2", "discriminator/d_fc_w", "discriminator/d_final_fc_w"):
        got, want = m.store.vars[k].grad.cpu().double(), wd["grads"][k]
        assert ((got - want).abs().max() / want.abs().max()).item() < 2e-3, k
    wg = ora.update(torch.tensor(inp), "g", apply=False)
    gg = m.update(torch.tensor(inp), "g", apply=False)
    assert abs(float(gg["g_loss"]) - wg["g_loss"]) < 2e-3 * max(1.0, abs(wg["g_loss"]))
    for k in ("generator/lstm/Cell0/Matrix", "generator/lstm/Cell2/Matrix", "generator/lstm/Cell1/Bias", "generator/output_fc_w", "generator/deconv_f1"):
        got, want = m.store.vars[k].grad.cpu().double(), wg["grads"][k]
        assert ((got - want).abs().max() / want.abs().max()).item() < 2e-3, k
    # one full step of the schedule against the golden trace
    got = m.train_step(torch.tensor(inp))
    assert abs(got["d_loss"] - g[tag + "/losses"][0]) < 2e-3 * max(1.0, abs(g[tag + "/losses"][0]))
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
