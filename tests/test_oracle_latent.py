"""CPU tests of the latent-search oracle (oracle/latent.py): the loss terms against their written-out definitions
(z_space_finder.py:258-292), the gradient with respect to z against central differences in float64, and the
process_batch schedule (learning-rate decay, warm starts, one Adam state for the whole run; z_space_finder.py:122-160)."""
import numpy as np
import pytest
import torch

from oracle.latent import LatentSearch, make_trained_like
from oracle.models import DCGAN


def tiny(dtype=torch.float64, B=3, seed=5):
    d = make_trained_like(DCGAN(batch_size=B, output_size=16, gf_dim=4, df_dim=4, seed=seed, dtype=dtype))
    tgt = np.random.RandomState(12).uniform(-1, 1, (B, 16, 16, 3))
    return d, tgt


ALL = dict(pixel_L2_weight=0.3, pixel_L1_weight=0.1, activations_L2_weight=0.3, activations_L1_weight=0.2, generator_loss_weight=0.1)


@pytest.mark.parametrize("mode", ["train", "inference"])
def test_terms_match_their_definitions(mode):
    d, tgt = tiny()
    s = LatentSearch(d, mode, **ALL)
    acts = s.target_activations(tgt)
    assert tuple(acts.shape) == (3, 2, 2, 16)
    terms = s.loss_terms(s.z, tgt, acts)
    train = mode == "train"
    with torch.no_grad():
        G = d.generator(s.z, train=train)
        _, logits, h2 = d.discriminator(G, train=train)
    t = torch.tensor(tgt)
    n_img, n_act = G[0].numel(), h2[0].numel()
    want = dict(
        pixel_L2=0.3 * np.mean([((G[b] - t[b]) ** 2).sum().item() / n_img for b in range(3)]),
        pixel_L1=0.1 * np.mean([(G[b] - t[b]).abs().sum().item() / n_img for b in range(3)]),
        activations_L2=0.3 * np.mean([((h2[b] - acts[b]) ** 2).sum().item() / n_act for b in range(3)]),
        activations_L1=0.2 * np.mean([(h2[b] - acts[b]).abs().sum().item() / n_act for b in range(3)]),
        generator=0.1 * np.mean([max(x, 0) - x + np.log1p(np.exp(-abs(x))) for x in logits.reshape(-1).tolist()]))
    # inference mode is deterministic; in train mode the only state touched between the two evaluations are the moving
    # averages, which the train-mode graph does not read
    for k, v in want.items():
        assert abs(terms[k].item() - v) < 1e-12 * max(1.0, abs(v)), k


@pytest.mark.parametrize("mode", ["train", "inference"])
def test_gradient_wrt_z_against_central_differences(mode):
    d, tgt = tiny()
    s = LatentSearch(d, mode, **ALL)
    acts = s.target_activations(tgt)
    _, g = s.loss_and_grad(tgt, acts)
    rs = np.random.RandomState(3)
    for _ in range(6):
        b, i = rs.randint(3), rs.randint(100)
        e = torch.zeros_like(s.z)
        e[b, i] = 1e-5
        with torch.no_grad():
            lp = sum(s.loss_terms(s.z + e, tgt, acts).values()).item()
            lm = sum(s.loss_terms(s.z - e, tgt, acts).values()).item()
        fd = (lp - lm) / 2e-5
        assert abs(fd - g[b, i].item()) < 1e-6 * max(1.0, abs(fd) * 1e3), (b, i, fd, g[b, i].item())


def test_zero_weight_terms_are_left_out_and_adam_moves_z():
    d, tgt = tiny()
    s = LatentSearch(d, "inference", activations_L2_weight=1.0)
    acts = s.target_activations(tgt)
    assert list(s.loss_terms(s.z, tgt, acts)) == ["activations_L2"]
    z0 = s.z.clone()
    loss0 = s.step(tgt, acts, 0.05)
    # first TF-Adam step: |dz| = lr_t * |m| / (sqrt(v) + eps) = lr * |g| / (|g| + eps * sqrt(1 - b2)) ~ lr wherever g != 0
    step = (s.z - z0).abs()
    assert step.max().item() <= 0.05 * (1 + 1e-9) and step.median().item() > 0.049
    for _ in range(30):
        loss = s.step(tgt, acts, 0.05)
    assert loss < 0.7 * loss0 and s.optim.t == 31


def test_fit_video_schedule():
    """2 initial steps at lr, then 1 step per frame at lr * decay; frame 0 is revisited first; latents warm-start."""
    d, _ = tiny(B=2)
    rs = np.random.RandomState(4)
    vids = rs.uniform(-1, 1, (2, 3, 16, 16, 3))
    s = LatentSearch(d, "inference", activations_L2_weight=0.5, pixel_L2_weight=0.5, random_seed=9)
    res, zs, losses = s.fit_video(vids, num_initial_steps=2, num_steps_per_frame=1, learning_rate=0.05, lr_decay_amount=0.5)
    assert res.shape == vids.shape and zs.shape == (2, 3, 100) and len(losses) == 5 and s.optim.t == 5
    # replay by hand
    d2, _ = tiny(B=2)
    m = LatentSearch(d2, "inference", activations_L2_weight=0.5, pixel_L2_weight=0.5, random_seed=9)
    acts = [m.target_activations(vids[:, f]) for f in range(3)]
    want = [m.step(vids[:, 0], acts[0], 0.05), m.step(vids[:, 0], acts[0], 0.05)]
    want.append(m.step(vids[:, 0], acts[0], 0.025))
    z_f0 = m.z.clone()
    want.append(m.step(vids[:, 1], acts[1], 0.025))
    want.append(m.step(vids[:, 2], acts[2], 0.025))
    assert np.allclose(losses, want, rtol=1e-12)
    assert np.allclose(zs[:, 0], z_f0.numpy()) and np.allclose(zs[:, 2], m.z.numpy())
    assert np.allclose(res[:, 2], m.images().numpy())


@pytest.mark.parametrize("mode", ["train", "inference"])
def test_latent_tiny_matches_golden(mode):
    """Regression pin of the oracle against tests/golden/latent_tiny.npz (made by tests/golden/make_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "latent_tiny.npz"))
    d = make_trained_like(DCGAN(batch_size=4, output_size=16, gf_dim=8, df_dim=8, seed=7, dtype=torch.float64))
    for k in d.vars:
        np.testing.assert_array_equal(d.vars[k].numpy(), g["weights/" + k])
    s = LatentSearch(d, mode, random_seed=3, **ALL)
    np.testing.assert_array_equal(s.z.numpy(), g[f"{mode}/z0"])
    acts = s.target_activations(g["targets"])
    np.testing.assert_allclose(acts.numpy(), g[f"{mode}/target_activations"], rtol=1e-9, atol=1e-12)
    loss0, g0 = s.loss_and_grad(g["targets"], acts)
    np.testing.assert_allclose(loss0, g[f"{mode}/loss0"], rtol=1e-10)
    np.testing.assert_allclose(g0.numpy(), g[f"{mode}/grad0"], rtol=1e-7, atol=1e-12)
    losses = [s.step(g["targets"], acts, 0.05) for _ in range(4)]
    np.testing.assert_allclose(losses, g[f"{mode}/losses"], rtol=1e-9)
    np.testing.assert_allclose(s.z.numpy(), g[f"{mode}/z4"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(s.images().numpy(), g[f"{mode}/images4"], rtol=1e-6, atol=1e-9)


# ---- nested search: through the video generator (discriminator_activation_optimizer_nested.py) --------------------------------
def tiny_nested(dtype=torch.float64, Bv=2, Tn=3):
    from oracle.models import VID_DCGAN
    v = VID_DCGAN(batch_size=Bv, z_input_size=12, z_output_size=10, vid_length=Tn, output_image_size=16, seed=9, dtype=dtype)
    make_trained_like(v)
    tgt = np.random.RandomState(13).uniform(-1, 1, (Bv, 16, 16, 3))
    return v, tgt


@pytest.mark.parametrize("mode", ["train", "inference"])
def test_nested_terms_use_first_frames_only(mode):
    from oracle.latent import NestedLatentSearch
    v, tgt = tiny_nested()
    s = NestedLatentSearch(v, mode, **ALL)
    assert tuple(s.z.shape) == (2, 12)
    acts = s.target_activations(tgt)
    train = mode == "train"
    with torch.no_grad():
        full = torch.zeros(6, 16, 16, 3, dtype=torch.float64)
        full[::3] = torch.tensor(tgt)
        want_acts = v.img_dcgan.discriminator(full, train=train)[2][::3]
        frames = v.img_dcgan.generator(v.generator(s.z, train=train), train=train)
        _, logits, h2 = v.img_dcgan.discriminator(frames, train=train)
    assert torch.equal(acts, want_acts) and tuple(frames.shape) == (6, 16, 16, 3)
    terms = s.loss_terms(s.z, tgt, acts)
    t = torch.tensor(tgt)
    want = dict(
        pixel_L2=0.3 * np.mean([((frames[3 * b] - t[b]) ** 2).mean().item() for b in range(2)]),
        pixel_L1=0.1 * np.mean([(frames[3 * b] - t[b]).abs().mean().item() for b in range(2)]),
        activations_L2=0.3 * np.mean([((h2[3 * b] - acts[b]) ** 2).mean().item() for b in range(2)]),
        activations_L1=0.2 * np.mean([(h2[3 * b] - acts[b]).abs().mean().item() for b in range(2)]),
        generator=0.1 * np.mean([max(x, 0) - x + np.log1p(np.exp(-abs(x))) for x in logits.reshape(-1).tolist()]))     # ALL frames
    for k, val in want.items():
        assert abs(terms[k].item() - val) < 1e-12 * max(1.0, abs(val)), k


@pytest.mark.parametrize("mode", ["train", "inference"])
def test_nested_gradient_against_central_differences(mode):
    from oracle.latent import NestedLatentSearch
    v, tgt = tiny_nested()
    s = NestedLatentSearch(v, mode, **ALL)
    acts = s.target_activations(tgt)
    _, g = s.loss_and_grad(tgt, acts)
    assert tuple(g.shape) == (2, 12) and g.abs().max() > 0
    rs = np.random.RandomState(4)
    for _ in range(6):
        b, i = rs.randint(2), rs.randint(12)
        e = torch.zeros_like(s.z)
        e[b, i] = 1e-7       # 6 rows through three ReLU layers and two L1 terms: kinks within 1e-5 of z; the slope is exact at 1e-7
        with torch.no_grad():
            lp = sum(s.loss_terms(s.z + e, tgt, acts).values()).item()
            lm = sum(s.loss_terms(s.z - e, tgt, acts).values()).item()
        fd = (lp - lm) / 2e-7
        assert abs(fd - g[b, i].item()) < 1e-5 * max(1.0, abs(fd)) + 1e-8, (b, i, fd, g[b, i].item())
    # one optimiser step moves z; the loss goes down over a few steps at a small step size
    l0 = s.step(tgt, acts, 0.01)
    for _ in range(5):
        l1 = s.step(tgt, acts, 0.01)
    assert l1 < l0
