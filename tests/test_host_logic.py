"""CPU-side tests of the host layer that mirrors the reference's Python (no kernels run): variable names and shapes of
the three models against the oracle's (= the reference's checkpoint keys), scope handling, optimiser
group layout, checkpoint round trip, and the host-only queries of the C ABI."""
import ctypes
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch


@pytest.fixture()
def cpu_store():
    from gifgan import ops
    ops.set_precision("fp32")
    return ops.reset_default_store(device="cpu", seed=1)


def test_dcgan_variable_names_and_shapes_match_the_reference(cpu_store):
    from gifgan.model import DCGAN
    from oracle.models import DCGAN as OracleDCGAN
    m = DCGAN(None, batch_size=4, output_size=64, c_dim=3)
    ora = OracleDCGAN(batch_size=4, output_size=64)
    assert set(m.store.vars) == set(ora.vars)
    for k, v in ora.vars.items():
        assert tuple(m.store.vars[k].shape) == tuple(v.shape), k
    # model.py:136-139: the var_lists split on the 'd_' / 'g_' substrings; parameter counts of SURVEY 8a
    assert sum(v.numel() for v in m.d_vars) == 4316545 and sum(v.numel() for v in m.g_vars) == 5135363
    # optimiser groups are contiguous, 16-byte aligned ranges of one flat buffer
    (db, de), (gb, ge) = m.store.ranges["d"], m.store.ranges["g"]
    assert db % 4 == 0 and gb % 4 == 0 and (de <= gb or ge <= db)
    for v in m.d_vars:
        assert db <= v.offset < de and v.grad.data_ptr() == m.store.flat["grads"][v.offset:].data_ptr()


def test_mnist_conditional_branch_names(cpu_store):
    from gifgan.model import DCGAN
    from oracle.models import DCGAN as OracleDCGAN
    m = DCGAN(None, batch_size=4, output_size=28, y_dim=10, c_dim=1, dataset_name="mnist")
    ora = OracleDCGAN(batch_size=4, output_size=28, y_dim=10, c_dim=1)
    assert set(m.store.vars) == set(ora.vars)


def test_video_models_resolve_their_scopes_from_anywhere(cpu_store):
    """z_model.py:63 builds VID_DCGAN under tf.variable_scope('video_gan'); train/sample code then runs outside it."""
    from gifgan import ops
    from gifgan.z_model_lib import VID_DCGAN
    from oracle.models import VID_DCGAN as OracleVID
    with ops.variable_scope("video_gan"):
        m = VID_DCGAN(None, batch_size=2, z_input_size=120, z_output_size=100, vid_length=16, input_image_size=64,
                      output_image_size=64, c_dim=3, sample_cols=2)
    ora = OracleVID(batch_size=2, vid_length=16, output_image_size=64)
    assert set(m.store.vars) == set(ora.vars)
    n_before = len(m.store.vars)
    z = torch.empty((2, 120), device="meta")                  # trace the whole chain outside every scope
    g, _ = m.generator(z)
    frames = m.img_dcgan.generator(g, train=False)
    act = m.img_dcgan.discriminator(frames, reuse=True, train=False, stop_at_h2=True)[2]
    logits = m.discriminator(act, reuse=True)[1]
    assert tuple(g.shape) == (32, 100) and tuple(frames.shape) == (32, 64, 64, 3)
    assert tuple(act.shape) == (32, 8, 8, 256) and tuple(logits.shape) == (2, 1)
    assert len(m.store.vars) == n_before                      # nothing was re-created under a wrong name
    # z_model_lib.py:166-179: default flags train the video nets only
    assert all("dvideo_" in v.name for v in m.d_var_list) and all("gvideo_" in v.name for v in m.g_var_list)
    m.set_trainable(train_img_gen=True, train_img_disc=True)
    assert any("image_gan/g_" in v.name for v in m.g_var_list) and any("image_gan/d_" in v.name for v in m.d_var_list)


def test_recurrent_dcgan_names(cpu_store):
    from gifgan.recurrent_dcgan import RecurrentDCGAN
    from oracle.models import RecurrentDCGAN as OracleRec
    m = RecurrentDCGAN(batch_size=2, video_length=3)
    ora = OracleRec(batch_size=2, video_length=3)
    assert set(m.store.vars) == set(ora.vars)
    for k, v in ora.vars.items():
        assert tuple(m.store.vars[k].shape) == tuple(v.shape), k


def test_checkpoint_round_trip(cpu_store, tmp_path):
    """model.py:428-452: <dir>/<dataset>_<batch>_<size>/DCGAN.model-<step> + the 'checkpoint' index file."""
    from gifgan import ops
    from gifgan.model import DCGAN
    m = DCGAN(None, batch_size=4, output_size=16, gf_dim=8, df_dim=8, c_dim=3, dataset_name="faces")
    w0 = {k: v.data.clone() for k, v in m.store.vars.items()}
    m.store.flat["m"].uniform_(-1, 1)
    m.d_optim.t, m.g_optim.t = 7, 14
    path = m.save(str(tmp_path), 123)
    assert path.endswith(os.path.join("faces_4_16", "DCGAN.model-123")) and os.path.exists(path)
    assert open(os.path.join(os.path.dirname(path), "checkpoint")).read().startswith('model_checkpoint_path: "DCGAN.model-123"')
    ops.reset_default_store(device="cpu", seed=99)            # different initial values
    m2 = DCGAN(None, batch_size=4, output_size=16, gf_dim=8, df_dim=8, c_dim=3, dataset_name="faces")
    assert not torch.equal(m2.store.vars["g_h1/w"].data, w0["g_h1/w"])
    assert m2.load(str(tmp_path))
    for k, v in w0.items():
        assert torch.equal(m2.store.vars[k].data, v), k
    assert torch.equal(m2.store.flat["m"], m.store.flat["m"]) and (m2.d_optim.t, m2.g_optim.t) == (7, 14)
    assert not m2.load(str(tmp_path / "nothing_here"))


def test_host_only_abi_queries():
    """Entry points that touch no device memory can be called without a GPU."""
    from gifgan import _cabi
    L = _cabi.lib()
    assert L.gg_bn_workspace_bytes(64, 2) >= 2 * 64 * 2 * 8
    d = _cabi.ConvDesc()
    for name, val in dict(N=64, D=1, H=32, W=32, C=64, Do=1, Ho=16, Wo=16, K=128, kd=1, kh=5, kw=5, sd=1, sh=2, sw=2, pd=0, ph=1, pw=1,
                          large_dtype=1, small_dtype=1, act=0, flags=_cabi.CONV_TENSOR_CORE).items():
        setattr(d, name, val)
    assert L.gg_conv_down(ctypes.byref(d), None, None, None, None, None) != 0      # null pointers are an error, not a crash
    assert b"null" in L.gg_last_error()


def test_stats_arena_bump_allocation():
    from gifgan import ops
    a = ops._StatsArena(torch.device("cpu"), nbytes=1024)
    x, y = a.take(3), a.take(4)
    assert x.numel() == 3 and y.numel() == 4 and y.data_ptr() - x.data_ptr() == 4 * 8      # 16-byte granules
    assert a.take(1000) is None                                # exhausted -> callers fall back to a fresh zero tensor
    t, pre = ops._zeroed_f64(5, torch.device("cpu"))
    assert not pre and float(t.abs().sum()) == 0.0


def test_cli_option_contract():
    """Option names / kinds / numeric defaults the reference's programs declare (main.py:10-29, z_model.py:22-56,
    model_sampler.py:9-24): a user's command lines keep working."""
    from gifgan import flags
    img = vars(flags.parse("image_gan", []))
    for name, default in dict(epoch=25, learning_rate=0.0002, beta1=0.5, batch_size=64, image_size=108, output_size=64, c_dim=3,
                              is_train=False, is_crop=False, visualize=False, shuffle=False, dataset="celebA", image_glob="*.jpg").items():
        assert img[name] == default, name
    assert img["train_size"] > 1e18                              # np.inf in the reference
    for name in ("checkpoint_dir", "sample_dir", "data_dir", "log_dir"):
        assert isinstance(img[name], str)
    vid = vars(flags.parse("video_gan", []))
    for name, default in dict(epoch=25, learning_rate=0.0002, beta1=0.5, image_batch_size=64, vid_batch_size=64, vid_length=16, image_size=64,
                              output_size=64, c_dim=3, is_train=False, video_shuffle=True, train_img_gen=False, train_img_disc=False,
                              disc_updates=1, gen_updates=2, image_noise=0.0, activation_noise=0.0, first_frame_loss_scalar=0.0,
                              sample_frequency=10, max_checkpoints_to_keep=5, video_list=[]).items():
        assert vid[name] == default, name
    for name in ("image_model_dir", "video_checkpoint_dir", "video_sample_dir", "video_data_dir", "video_dataset", "log_dir"):
        assert isinstance(vid[name], str)
    smp = vars(flags.parse("sampler", ["--continuous", "--num_samples", "7"]))
    assert smp["continuous"] is True and smp["num_samples"] == 7 and smp["vid_length"] == 16 and smp["random_seed"] == 0
    got = flags.parse("video_gan", ["--video_list", "a.txt", "b.txt", "--train_img_gen", "true", "--gen_updates", "1"])
    assert got.video_list == ["a.txt", "b.txt"] and got.train_img_gen is True and got.gen_updates == 1


def test_latent_search_host_side(cpu_store):
    """gifgan.latent_search without a device: weight normalisation (z_space_finder.py:229-238), the uniform(-1, 1)
    latents, argument checks -- and the loud failure of the first kernel call (no CPU fallback)."""
    import argparse
    from gifgan.latent_search import LatentSearch, WEIGHT_NAMES, normalised_weights
    from gifgan.model import DCGAN
    opts = argparse.Namespace(pixel_L2_weight=1.0, pixel_L1_weight=0.0, activations_L2_weight=2.0, activations_L1_weight=1.0,
                              generator_loss_weight=0.0)
    w = normalised_weights(opts)
    assert tuple(w) == WEIGHT_NAMES and abs(sum(w.values()) - 1.0) < 1e-12 and w["activations_L2_weight"] == 0.5
    with pytest.raises(ValueError):
        normalised_weights(argparse.Namespace(**{k: 0.0 for k in WEIGHT_NAMES}))
    m = DCGAN(None, batch_size=3, output_size=16, gf_dim=8, df_dim=8)
    s = LatentSearch(m, "inference", random_seed=5, **w)
    want = np.random.RandomState(5).uniform(-1.0, 1.0, size=(3, 100)).astype(np.float32)
    assert np.array_equal(s.z.detach().numpy(), want) and s.z.requires_grad and s.t == 0
    assert abs(s.lr_t.__func__(type("T", (), dict(beta1=0.5, beta2=0.999, t=1))(), 0.05) - 0.05 * np.sqrt(1 - 0.999) / 0.5) < 1e-12
    with pytest.raises(ValueError):
        LatentSearch(m, "sometimes")
    with pytest.raises(ValueError):
        LatentSearch(m, "train", z=np.zeros((2, 100)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.step(np.zeros((3, 16, 16, 3), np.float32), np.zeros((3, 2, 2, 32), np.float32), 0.05)


def test_latent_cli_contract_and_clip_reader(tmp_path):
    """Option names / defaults of z_space_finder.py:11-41 and discriminator_activation_optimizer.py:16-45; the clip
    reader's frame skipping, resize, BGR->RGB and x/127.5-1 (z_space_finder.py:70-88); batching of synthetic clips."""
    import cv2
    from gifgan import flags
    from gifgan.latent_search import read_video_frames
    from gifgan.z_space_finder import clip_batches, out_path
    zf = vars(flags.parse("z_space_finder", ["--discriminator_mode", "inference", "--video_list", "a.txt", "b.txt"]))
    for name, default in dict(video_batch_size=8, stop_after=0, random_seed=0, num_initial_steps=500, num_steps_per_frame=100,
                              learning_rate=0.05, lr_decay_amount=0.5, beta1=0.5, vid_length=16, frame_skip=2, pixel_L2_weight=0.0,
                              pixel_L1_weight=0.0, activations_L2_weight=1.0, activations_L1_weight=0.0, generator_loss_weight=0.0,
                              image_size=64, output_size=64, c_dim=3, output_comparison_folder="", output_image_folder="",
                              output_frame_folder="").items():
        assert zf[name] == default, name
    assert zf["video_list"] == ["a.txt", "b.txt"] and zf["discriminator_mode"] == "inference"
    ao = vars(flags.parse("activation_optimizer", ["--input_images", "x.png"]))
    for name, default in dict(num_rows=8, num_cols=8, num_steps=1000, learning_rate=0.0002, beta1=0.5, sample_frequency=100,
                              lr_decay_frequency=0, lr_decay_amount=0.9, activations_L2_weight=1.0, input_videos=[]).items():
        assert ao[name] == default, name
    assert out_path("some/dir/clip7.mp4", "/out") == "/out/clip7.npy" and out_path("clip7.mp4", "/o", ".png") == "/o/clip7.png"

    path = str(tmp_path / "clip.avi")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (48, 32))
    if not wr.isOpened():
        pytest.skip("no MJPG writer in this OpenCV build")
    for i in range(9):                                           # frame i: blue channel (BGR order) = 20 * i, red = 200
        frame = np.zeros((32, 48, 3), np.uint8)
        frame[..., 0], frame[..., 2] = 20 * i, 200
        wr.write(frame)
    wr.release()
    frames = read_video_frames(path, 16, 4, 2)                   # frames 1, 3, 5, 7 of the file
    assert len(frames) == 4 and frames[0].shape == (16, 16, 3)
    for k, f in enumerate(frames):
        assert abs(f[..., 0].mean() - (200 / 127.5 - 1)) < 0.06                       # RGB: red first
        assert abs(f[..., 2].mean() - (20 * (2 * k + 1) / 127.5 - 1)) < 0.06, k
    assert read_video_frames(path, 16, 5, 2) is None             # 10 frames needed, 9 in the file

    opts = flags.parse("z_space_finder", ["--synthetic", "5", "--video_batch_size", "2", "--vid_length", "3", "--image_size", "8"])
    batches = list(clip_batches(opts))
    assert [len(n) for n, _ in batches] == [2, 2, 1] and np.shape(batches[0][1]) == (2, 3, 8, 8, 3)


@pytest.mark.parametrize("kw", [dict(), dict(num_layers=3), dict(num_layers=3, shared_conv=True, output_keep_prob=0.8)])
def test_recurrent_variants_names_and_shapes(cpu_store, kw):
    """recurrent_DCGAN.py and its two variants (multi-layer_recurrent_DCGAN.py, ..._with_shared_conv_and_drop_out.py):
    variable inventory against the oracle's, and the graph traced on meta tensors (shapes only, no kernels)."""
    from gifgan.recurrent_dcgan import RecurrentDCGAN
    from oracle.models import RecurrentDCGAN as OracleRec
    m = RecurrentDCGAN(batch_size=2, video_length=3, **kw)
    ora = OracleRec(batch_size=2, video_length=3, **kw)
    assert set(m.store.vars) == set(ora.vars)
    for k, v in ora.vars.items():
        assert tuple(m.store.vars[k].shape) == tuple(v.shape), k
    assert [v.name for v in m.g_vars] == ora.g_vars and [v.name for v in m.d_vars] == ora.d_vars
    if kw.get("shared_conv"):
        assert not any("/conv_f" in v.name for v in m.g_vars)                    # the encoder's filters are the discriminator's
        assert tuple(m.store.vars["generator/lstm/Cell0/Matrix"].shape) == (200, 400)
    elif kw.get("num_layers"):
        assert tuple(m.store.vars["generator/lstm/Cell0/Matrix"].shape) == (8292, 400)
        assert tuple(m.store.vars["generator/lstm/Cell2/Matrix"].shape) == (200, 400)
    X = torch.empty((3 * 2, 64, 64, 3), dtype=torch.float32, device="meta")
    fake = m.generator(X)
    assert tuple(fake.shape) == (6, 64, 64, 3) and tuple(m.discriminator(fake).shape) == (2, 1)
    if kw.get("output_keep_prob"):
        m.store.device = torch.device("cpu")
        mk = m._mask(1)
        assert tuple(mk.shape) == (3, 2, 100) and set(np.unique(mk.numpy()).round(4)) <= {0.0, 1.25}


def test_activation_optimizer_clip_mode_targets(tmp_path):
    """--vid_length N (discriminator_activation_optimizer_video.py:64-103): every frame_skip-th frame of every clip,
    clip-major, one grid row per clip."""
    import cv2
    from gifgan import flags
    from gifgan.discriminator_activation_optimizer import load_clip_targets
    paths = []
    for c in range(2):
        path = str(tmp_path / ("c%d.avi" % c))
        wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (32, 32))
        if not wr.isOpened():
            pytest.skip("no MJPG writer in this OpenCV build")
        for i in range(7):
            wr.write(np.full((32, 32, 3), 20 * i + 100 * c, np.uint8))
        wr.release()
        paths.append(path)
    opts = flags.parse("activation_optimizer", ["--vid_length", "3", "--image_size", "16", "--input_videos"] + paths)
    assert opts.frame_skip == 2
    t = load_clip_targets(opts)
    assert t.shape == (6, 16, 16, 3) and t.dtype == np.float32
    want = [(20 * i + 100 * c) / 127.5 - 1 for c in range(2) for i in (1, 3, 5)]
    assert np.allclose(t.mean(axis=(1, 2, 3)), want, atol=0.05)
    opts = flags.parse("activation_optimizer", ["--vid_length", "4", "--image_size", "16", "--input_videos"] + paths)
    with pytest.raises(SystemExit):
        load_clip_targets(opts)                                   # 8 frames needed, 7 in the file
    opts = flags.parse("activation_optimizer", ["--vid_length", "4", "--synthetic", "3", "--image_size", "8"])
    assert load_clip_targets(opts).shape == (12, 8, 8, 3)


def test_activation_optimizer_iterative_contract_and_tweening():
    """discriminator_activation_optimizer_video_iterative.py: option names / defaults (lines 15-23) and the latent tweening of
    lines 243-262 (tween_frames latents between consecutive tracked frames: end * delta + start * (1 - delta), delta = j / (n + 1))."""
    from gifgan import flags
    from gifgan.discriminator_activation_optimizer import tween_latents
    o = flags.parse("activation_optimizer", ["--vid_length", "4", "--iterative", "--synthetic", "2"])
    assert o.iterative and (o.num_initial_steps, o.num_steps_per_frame, o.frame_skip, o.tween_frames) == (500, 100, 2, 2)
    assert not flags.parse("activation_optimizer", ["--vid_length", "4"]).iterative
    zs = np.arange(2 * 3 * 5, dtype=np.float32).reshape(2, 3, 5)
    seq = tween_latents(zs, 2)
    assert [i for i, _, _ in seq] == list(range(7)) and [t for _, _, t in seq] == [True, False, False, True, False, False, True]
    assert np.array_equal(seq[0][1], zs[:, 0]) and np.array_equal(seq[3][1], zs[:, 1]) and np.array_equal(seq[6][1], zs[:, 2])
    np.testing.assert_allclose(seq[1][1], zs[:, 1] * (1 / 3.0) + zs[:, 0] * (2 / 3.0), rtol=1e-6)
    np.testing.assert_allclose(seq[5][1], zs[:, 2] * (2 / 3.0) + zs[:, 1] * (1 / 3.0), rtol=1e-6)
    assert [i for i, _, _ in tween_latents(zs, 0)] == [0, 1, 2]


def test_clip_grid_writer_produces_a_readable_mp4(tmp_path):
    """write_clip_grid (discriminator_activation_optimizer_nested.py:305-324) / utils.open_video_writer: the reference's fourcc 0x20 is
    rejected by OpenCV 4 for .mp4 (nothing would be written); the fallback tag must give T frames of the rows x cols grid."""
    cv2 = pytest.importorskip("cv2")
    from gifgan.discriminator_activation_optimizer import write_clip_grid
    v = np.random.RandomState(0).uniform(-1, 1, (2, 3, 5, 16, 16, 3)).astype(np.float32)
    path = str(tmp_path / "grid.mp4")
    write_clip_grid(v, path)
    cap = cv2.VideoCapture(path)
    n, shape = 0, None
    while True:
        ok, im = cap.read()
        if not ok:
            break
        n, shape = n + 1, im.shape
    assert n == 5 and shape == (32, 48, 3)


def test_new_entry_points_have_no_cpu_fallback_and_trace_on_meta(cpu_store):
    """frames_to_input / the fused loss head: shape inference on meta tensors (what the model constructors trace), a loud error on
    host tensors (the product never computes on the CPU), argument validation before any launch."""
    from gifgan import ops
    fr = torch.zeros(3, 12, 16, 3, dtype=torch.uint8)
    assert tuple(ops.frames_to_input(fr.to("meta"), 8).shape) == (3, 8, 8, 3)
    with pytest.raises(RuntimeError):
        ops.frames_to_input(fr, 8)
    with pytest.raises(ValueError):
        ops.frames_to_input(fr.float(), 8)
    with pytest.raises(ValueError):
        ops.frames_to_input(torch.zeros(3, 12, 16, 4, dtype=torch.uint8), 8)
    # linear(..., ce_segments=...) on meta tensors creates the variables and returns logits of the right shape without a launch
    h = torch.empty(6, 128, device="meta")
    y = ops.linear(h, 1, "d_h3_lin", ce_segments=[(0, 3, 1.0, 1.0), (3, 6, 0.0, 1.0)])
    assert tuple(y.shape) == (6, 1) and "d_h3_lin/Matrix" in cpu_store.vars
    L = ops.cabi.lib()
    assert L.gg_loss_head_ok(128, 8192, 2) == 1 and L.gg_loss_head_ok(128, 100, 2) == 0 and L.gg_loss_head_ok(128, 8192, 5) == 0


def test_image_side_index_maps_emulated_on_the_cpu():
    """tools/emulate_c3_mma.py mirrors the thread / fragment index maps of csrc/conv_c3_mma.cu (mma.sync m16n8k16, ldmatrix) line by
    line, including the bias rider of the filter-gradient kernel (ones channel in the staged patch -> the discarded row (1, 1, c = 3)
    sums dy over the pixels) on partial tiles: the emulated kernels must reproduce the direct sums."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "emulate_c3_mma.py")
    spec = importlib.util.spec_from_file_location("emulate_c3_mma", path)
    E = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(E)
    rs = np.random.RandomState(0)
    x, w = rs.randn(1, 20, 36, 3), rs.randn(5, 5, 3, 64)
    assert np.abs(E.emu_down(x, w) - E.ref_down(x, w)).max() < 1e-9
    xs = rs.randn(1, 10, 18, 64)
    assert np.abs(E.emu_up(xs, w) - E.ref_up(xs, w)).max() < 1e-9
    ys = rs.randn(1, 10, 18, 64)
    dw, db = E.emu_wgrad(x, ys, with_bias=True)
    assert np.abs(dw - E.ref_wgrad(x, ys)).max() < 1e-9 and np.abs(db - ys.sum((0, 1, 2))).max() < 1e-9
