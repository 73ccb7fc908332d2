"""GPU parity tests for the tcgen05 / TMEM / TMA kernels (bf16 mode, tensor-core path) against the CPU oracle:
every DCGAN-64 tensor-core layer shape (d_h1..3, g_h1..3), the video discriminator's conv3d shapes, ragged
sizes that exercise the TMA zero-fill and partial tiles, and the forward/dgrad adjointness property at
BASELINE.json's full config-2 batch.  Tolerance: 2e-2 relative (north_star, bf16)."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tf_ops as T  # noqa: E402

TOL = 2e-2


def relerr(got, want):
    got = got.detach().float().cpu().double()
    want = want.detach().double()
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-12)).item()


def bf16_round(a):
    return torch.tensor(a, dtype=torch.float32).to(torch.bfloat16).float()


def _store(builder, shape):
    from gifgan import ops
    ops.set_precision("bf16", tensor_cores=True)
    st = ops.reset_default_store(device="cuda", seed=5)
    builder(torch.empty(shape, device="meta", dtype=torch.bfloat16))
    tv = [v for v in st.vars.values() if v.trainable]
    st.finalize(OrderedDict(all=tv))
    return ops, st, tv


# (name, B, H, Cin, Cout): discriminator convs of DCGAN-64 (model.py:273-276) that run on tensor cores
@pytest.mark.parametrize("name,B,H,Ci,Co", [("d_h1", 16, 32, 64, 128), ("d_h2", 16, 16, 128, 256), ("d_h3", 16, 8, 256, 512),
                                           ("ragged", 3, 14, 64, 192), ("wide", 2, 40, 128, 64)])
def test_tc_conv2d(name, B, H, Ci, Co):
    rs = np.random.RandomState(Ci + H)
    W = H if name != "ragged" else 10
    x, w, b = bf16_round(rs.randn(B, H, W, Ci)), bf16_round(rs.randn(5, 5, Ci, Co) * 0.05), torch.tensor(rs.randn(Co) * 0.1, dtype=torch.float32)
    from gifgan import ops as _o
    ops, st, tv = _store(lambda t: _o.conv2d(t, Co, name="c"), (B, H, W, Ci))
    st.load_state_dict({"c/w": w.numpy(), "c/biases": b.numpy()})
    xt = x.cuda().to(torch.bfloat16).requires_grad_(True)
    assert ops._tc_ok(Ci, Co, xt)
    Ho, Wo = -(-H // 2), -(-W // 2)
    dy = bf16_round(rs.randn(B, Ho, Wo, Co))
    with ops.trainable(tv):
        y = ops.conv2d(xt, Co, name="c")
        y.backward(dy.cuda().to(torch.bfloat16))
    xr = x.double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = T.conv2d(xr, wr, br)
    gx, gw, gb = torch.autograd.grad(yr, [xr, wr, br], dy.double())
    assert y.dtype == torch.bfloat16 and tuple(y.shape) == (B, Ho, Wo, Co)
    assert relerr(y, yr) < TOL
    assert relerr(xt.grad, gx) < TOL
    assert relerr(st.vars["c/w"].grad, gw) < TOL
    assert relerr(st.vars["c/biases"].grad, gb) < TOL


@pytest.mark.parametrize("name,B,h,Ci,Co", [("g_h1", 16, 4, 512, 256), ("g_h2", 16, 8, 256, 128), ("g_h3", 16, 16, 128, 64),
                                           ("ragged", 3, 7, 128, 64)])
def test_tc_deconv2d(name, B, h, Ci, Co):
    rs = np.random.RandomState(Ci + h)
    w_ = h if name != "ragged" else 5
    x, w, b = bf16_round(rs.randn(B, h, w_, Ci)), bf16_round(rs.randn(5, 5, Co, Ci) * 0.05), torch.tensor(rs.randn(Co) * 0.1, dtype=torch.float32)
    out_shape = [B, 2 * h, 2 * w_, Co]
    from gifgan import ops as _o
    ops, st, tv = _store(lambda t: _o.deconv2d(t, out_shape, name="g"), (B, h, w_, Ci))
    st.load_state_dict({"g/w": w.numpy(), "g/biases": b.numpy()})
    xt = x.cuda().to(torch.bfloat16).requires_grad_(True)
    dy = bf16_round(rs.randn(*out_shape))
    with ops.trainable(tv):
        y = ops.deconv2d(xt, out_shape, name="g")
        y.backward(dy.cuda().to(torch.bfloat16))
    xr = x.double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = T.conv2d_transpose(xr, wr, out_shape, br)
    gx, gw, gb = torch.autograd.grad(yr, [xr, wr, br], dy.double())
    assert relerr(y, yr) < TOL
    assert relerr(xt.grad, gx) < TOL
    assert relerr(st.vars["g/w"].grad, gw) < TOL
    assert relerr(st.vars["g/biases"].grad, gb) < TOL


# video discriminator (z_model_lib.py:409-413): [Bv,16,8,8,256] -> [Bv,8,4,4,256] -> [Bv,4,2,2,256] -> [Bv,2,1,1,256]
@pytest.mark.parametrize("B,D,H", [(2, 16, 8), (4, 8, 4), (8, 4, 2), (3, 6, 5)])
def test_tc_conv3d(B, D, H):
    rs = np.random.RandomState(D)
    Ci = Co = 256
    x, w, b = bf16_round(rs.randn(B, D, H, H, Ci)), bf16_round(rs.randn(3, 3, 3, Ci, Co) * 0.03), torch.tensor(rs.randn(Co) * 0.1, dtype=torch.float32)
    from gifgan import ops as _o
    ops, st, tv = _store(lambda t: _o.conv3d(t, Co, name="v"), (B, D, H, H, Ci))
    st.load_state_dict({"v/w": w.numpy(), "v/biases": b.numpy()})
    xt = x.cuda().to(torch.bfloat16).requires_grad_(True)
    xr = x.double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = T.conv3d(xr, wr, br)
    dy = bf16_round(rs.randn(*yr.shape))
    with ops.trainable(tv):
        y = ops.conv3d(xt, Co, name="v")
        y.backward(dy.cuda().to(torch.bfloat16))
    gx, gw, gb = torch.autograd.grad(yr, [xr, wr, br], dy.double())
    assert relerr(y, yr) < TOL
    assert relerr(xt.grad, gx) < TOL
    assert relerr(st.vars["v/w"].grad, gw) < TOL


def test_tc_matches_simt_path_and_accumulates():
    """Same library, two kernels: tensor-core vs SIMT on identical bf16 operands; and wgrad accumulates (+=)."""
    from gifgan import ops as _o
    B, H, Ci, Co = 64, 16, 128, 256
    ops, st, tv = _store(lambda t: _o.conv2d(t, Co, name="c"), (B, H, H, Ci))
    x = torch.randn(B, H, H, Ci, device="cuda").to(torch.bfloat16)
    dy = torch.randn(B, H // 2, H // 2, Co, device="cuda").to(torch.bfloat16)
    res = {}
    for tc in (True, False):
        ops.set_precision("bf16", tensor_cores=tc)
        for v in tv:
            v.grad.zero_()
        xt = x.clone().requires_grad_(True)
        with ops.trainable(tv):
            y = ops.conv2d(xt, Co, name="c")
            y.backward(dy)
            if tc:                      # a second use of the same filter adds its gradient
                y2 = ops.conv2d(xt, Co, name="c")
                y2.backward(dy)
        res[tc] = (y.float(), xt.grad.float(), st.vars["c/w"].grad.clone())
    assert relerr(res[True][0].cpu(), res[False][0].cpu()) < 1e-2
    assert relerr(res[True][1].cpu(), 2 * res[False][1].cpu()) < 1e-2
    assert relerr(res[True][2].cpu(), 2 * res[False][2].cpu()) < 1e-3


def test_tc_adjointness_full_batch():
    """<conv(x), dy> == <x, dgrad(dy)> at config-2 batch (2B = 128 images through d_h2): no oracle needed."""
    from gifgan import ops as _o
    B, H, Ci, Co = 128, 16, 128, 256
    ops, st, tv = _store(lambda t: _o.conv2d(t, Co, name="c", bias=False), (B, H, H, Ci))
    gen = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(B, H, H, Ci, device="cuda", generator=gen).to(torch.bfloat16).requires_grad_(True)
    dy = torch.randn(B, H // 2, H // 2, Co, device="cuda", generator=gen).to(torch.bfloat16)
    y = ops.conv2d(x, Co, name="c", bias=False)
    y.backward(dy)
    lhs = (y.float().double() * dy.float().double()).sum().item()
    rhs = (x.detach().float().double() * x.grad.float().double()).sum().item()
    # both inner products are zero-mean random sums (|lhs| ~ ||y|| ||dy|| / sqrt(n) can come out small by chance): scale the
    # tolerance by the Cauchy-Schwarz bound instead of |lhs|; bf16 rounding of y / dx gives ~4e-3 / sqrt(n) of it
    scale = y.float().double().norm().item() * dy.float().double().norm().item()
    assert abs(lhs - rhs) < 1e-4 * scale, (lhs, rhs, scale)


# ---- image-side layers (3 channels <-> 64k channels): warp-MMA kernels of conv_c3_mma.cu in bf16 mode --------------
# d_h0_conv (model.py:273): fp32 image in, bf16 activation out.  Sizes cover: one partial tile (20x20 -> 10x10),
# several tiles per image (64x64), two 64-channel blocks (Co=128), many images per persistent CTA.
@pytest.mark.parametrize("B,H,W,Co", [(2, 16, 16, 64), (3, 20, 20, 64), (2, 64, 64, 64), (5, 32, 32, 128), (40, 64, 64, 64)])
def test_c3_conv2d_image_side(B, H, W, Co):
    rs = np.random.RandomState(H + Co + B)
    x = torch.tensor(rs.uniform(-1, 1, (B, H, W, 3)), dtype=torch.float32)
    w, b = bf16_round(rs.randn(5, 5, 3, Co) * 0.05), torch.tensor(rs.randn(Co) * 0.1, dtype=torch.float32)
    from gifgan import ops as _o
    ops, st, tv = _store(lambda t: _o.conv2d(t, Co, name="c"), (B, H, W, 3))
    st.load_state_dict({"c/w": w.numpy(), "c/biases": b.numpy()})
    xt = x.cuda().requires_grad_(True)                      # images stay fp32 in HBM
    Ho, Wo = H // 2, W // 2
    dy = bf16_round(rs.randn(B, Ho, Wo, Co))
    with ops.trainable(tv):
        y = ops.conv2d(xt, Co, name="c")
        y.backward(dy.cuda().to(torch.bfloat16))
    xr = x.double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = T.conv2d(xr, wr, br)
    gx, gw, gb = torch.autograd.grad(yr, [xr, wr, br], dy.double())
    assert y.dtype == torch.bfloat16 and tuple(y.shape) == (B, Ho, Wo, Co)
    assert relerr(y, yr) < TOL
    assert relerr(xt.grad, gx) < TOL
    assert relerr(st.vars["c/w"].grad, gw) < TOL
    assert relerr(st.vars["c/biases"].grad, gb) < TOL


# g_h4 (model.py:321): bf16 activation in, fp32 image out (tanh applied by the caller in the model; here plain).
# act = "tanh" is the model's form (model.py:324): its backward runs tanh' and the 3-channel bias gradient as one launch (gg_act_bwd_bias)
@pytest.mark.parametrize("act", [None, "tanh"])
@pytest.mark.parametrize("B,h,w_,Ci", [(2, 8, 8, 64), (3, 10, 6, 64), (2, 32, 32, 64), (4, 16, 16, 128), (40, 32, 32, 64)])
def test_c3_deconv2d_image_side(B, h, w_, Ci, act):
    rs = np.random.RandomState(h + Ci + B)
    x, w, b = bf16_round(rs.randn(B, h, w_, Ci)), bf16_round(rs.randn(5, 5, 3, Ci) * 0.05), torch.tensor(rs.randn(3) * 0.1, dtype=torch.float32)
    out_shape = [B, 2 * h, 2 * w_, 3]
    from gifgan import ops as _o
    ops, st, tv = _store(lambda t: _o.deconv2d(t, out_shape, name="g", act=act, out_dtype=torch.float32), (B, h, w_, Ci))
    st.load_state_dict({"g/w": w.numpy(), "g/biases": b.numpy()})
    xt = x.cuda().to(torch.bfloat16).requires_grad_(True)
    dy = torch.tensor(rs.randn(*out_shape), dtype=torch.float32)
    with ops.trainable(tv):
        y = ops.deconv2d(xt, out_shape, name="g", act=act, out_dtype=torch.float32)
        y.backward(dy.cuda().to(y.dtype))
    xr = x.double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = T.conv2d_transpose(xr, wr, out_shape, br)
    if act == "tanh":
        yr = torch.tanh(yr)
    gx, gw, gb = torch.autograd.grad(yr, [xr, wr, br], dy.double())
    assert tuple(y.shape) == tuple(out_shape)
    assert relerr(y, yr) < TOL
    assert relerr(xt.grad, gx) < TOL
    assert relerr(st.vars["g/w"].grad, gw) < TOL
    assert relerr(st.vars["g/biases"].grad, gb) < TOL


@pytest.mark.parametrize("mode,splitk", [("l2", "auto"), ("l2", 2), ("l2", 4), ("l2", 8), ("l2", 1), ("cluster", "auto"), ("cluster", 2), ("cluster", 4)])
def test_tc_split_k(mode, splitk, monkeypatch):
    """Split-K of the tcgen05 pixel GEMM (GG_TC_SPLITK: auto = the library's cycle model, 2 / 4 / 8 forced, 1 = off): the K loop of
    a wide tile split over S CTAs whose fp32 partials meet in a reduce-scatter -- through L2 (default: TMA stores / loads of
    16 KB boxes in the workspace + a per-tile counter; deterministic) or, opt-in, through a thread-block cluster's distributed
    shared memory -- must give the same conv / deconv results (fp32-accumulated partials, one final rounding): forward, input
    gradient, fused statistics, fused batch-norm backward reductions.  The L2 variant is also run twice for bit equality."""
    monkeypatch.setenv("GG_TC_SPLITK", str(splitk))
    monkeypatch.setenv("GG_TC_SPLITK_MODE", mode)
    try:
        test_tc_conv2d("d_h3", 16, 8, 256, 512)
        test_tc_deconv2d("g_h1", 16, 4, 512, 256)
        test_tc_conv2d("d_h2", 16, 16, 128, 256)
        test_bn_backward_reductions_fused_into_dgrad("conv", 16, 16, 128, 256, 512, 2, monkeypatch)      # d_h2 -> bn2 -> d_h3: split dgrad + fused reductions
        test_bn_backward_reductions_fused_into_dgrad("deconv", 16, 4, 512, 256, 128, 1, monkeypatch)
    finally:
        monkeypatch.delenv("GG_TC_SPLITK", raising=False)
        monkeypatch.delenv("GG_TC_SPLITK_MODE", raising=False)


def test_tc_split_k_full_batch_deterministic_and_equal_to_unsplit(monkeypatch):
    """At the bench shapes (batch 64: g_h1 / d_h3, where the split is chosen by default) the split launch must (a) reproduce
    itself bit for bit -- the partials are summed in rank order -- and (b) agree with the unsplit launch to fp32 summation-order
    noise on the fp32 pre-norm output and its fused statistics."""
    from gifgan import ops as _o
    B, H, Ci, Co = 64, 8, 256, 512
    outs = {}
    for sk in ("1", "auto", "auto", "8"):
        monkeypatch.setenv("GG_TC_SPLITK", sk)
        ops, st, tv = _store(lambda t: _o.conv2d(t, Co, name="c"), (B, H, H, Ci))
        gen = torch.Generator(device="cuda").manual_seed(3)
        x = torch.randn(B, H, H, Ci, device="cuda", generator=gen).to(torch.bfloat16).requires_grad_(True)
        dy = torch.randn(B, H // 2, H // 2, Co, device="cuda", generator=gen).to(torch.bfloat16)
        y = ops.conv2d(x, Co, name="c")
        y.backward(dy)
        torch.cuda.synchronize()
        outs.setdefault(sk, []).append((y.float().clone(), x.grad.float().clone()))
    monkeypatch.delenv("GG_TC_SPLITK", raising=False)
    a, b = outs["auto"]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    for k in ("auto", "8"):
        for i in range(2):
            ref, got = outs["1"][0][i], outs[k][0][i]
            assert ((got - ref).abs().max() / ref.abs().max()).item() < 1e-2, (k, i)      # one bf16 ulp of the largest element at most


def test_tc_wgrad_deterministic_mode(monkeypatch):
    """GG_DETERMINISTIC=1: the tensor-core filter-gradient kernel runs without its pixel split (one reduce-add per element), so two
    runs give bit-identical gradients -- and they agree with the split (default) launch to fp32 summation-order noise."""
    from gifgan import ops as _o
    B, H, Ci, Co = 64, 16, 128, 256
    ops, st, tv = _store(lambda t: _o.conv2d(t, Co, name="c", bias=False), (B, H, H, Ci))
    gen = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(B, H, H, Ci, device="cuda", generator=gen).to(torch.bfloat16)
    dy = torch.randn(B, H // 2, H // 2, Co, device="cuda", generator=gen).to(torch.bfloat16)
    g = ops._Geom(B, (1, H, H), Ci, (1, H // 2, H // 2), Co, (1, 5, 5), (1, 2, 2), (0, 1, 1))
    wv = st.vars["c/w"]
    outs = []
    for mode in ("1", "1", "0"):
        monkeypatch.setenv("GG_DETERMINISTIC", mode)
        wv.grad.zero_()
        ops._run_wgrad(g, x, dy, wv)
        ops.join_side()
        torch.cuda.synchronize()
        outs.append(wv.grad.clone())
    monkeypatch.delenv("GG_DETERMINISTIC", raising=False)
    assert torch.equal(outs[0], outs[1])
    assert ((outs[2] - outs[0]).abs().max() / outs[0].abs().max()).item() < 1e-5


def test_adam_keeps_the_bf16_shadow_current():
    """gg_adam_graph(p, p_bf16, ...): the bf16 shadow of the flat parameter buffer -- the filter operand of every tensor-core
    kernel -- is rewritten by the Adam launch itself and equals the rounded fp32 masters bit for bit."""
    from gifgan import _cabi, ops
    L = _cabi.lib()
    n = 100003                                                      # odd tail: the scalar path writes the shadow too
    rs = np.random.RandomState(5)
    p = torch.tensor(rs.randn(n + 5).astype(np.float32)).cuda()[:n]
    g = torch.tensor(rs.randn(n + 5).astype(np.float32)).cuda()[:n]
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    pb = torch.full((n,), 7.0, dtype=torch.bfloat16, device="cuda")
    state = torch.zeros(4, dtype=torch.int32, device="cuda")
    for _ in range(3):
        ops.check(L.gg_adam_graph(ops.ptr(p), ops.ptr(pb), ops.ptr(g), ops.ptr(m), ops.ptr(v), n, ops.ptr(state), 2e-4, 0.5, 0.999, 1e-8, 1.0,
                                  ops.stream()), "gg_adam_graph")
        assert torch.equal(pb, p.to(torch.bfloat16))
    assert int(state[0]) == 3


def test_train_step_enters_the_graph_with_current_filter_copies():
    """No re-pack launches: the captured DCGAN-64 step must follow the eager step (which re-casts stale copies at first
    use), and a weight change behind the optimisers' back (checkpoint load) must be picked up before the next replay."""
    from gifgan import ops
    from gifgan.model import DCGAN
    B = 8
    img = np.random.RandomState(102).uniform(-1, 1, (B, 64, 64, 3)).astype(np.float32)
    runs = []
    for graph in (False, True):
        ops.set_precision("bf16")
        ops.reset_default_store(device="cuda", seed=7)
        m = DCGAN(None, batch_size=B, output_size=64, c_dim=3)
        losses = []
        for step in range(3):
            z = np.random.RandomState(1000 + step).uniform(-1, 1, (B, 100)).astype(np.float32)
            o = m.train_step(img, z, use_graph=graph)
            losses.append([o["d_loss"], o["g_loss_first"], o["g_loss"]])
            w = m.store.vars["d_h2_conv/w"]
            assert not w.packs_stale() and torch.equal(w._bf16, w.data.to(torch.bfloat16))       # shadow == rounded masters
        runs.append(np.array(losses))
        if graph:
            sd = m.store.state_dict()
            sd["d_h1_conv/w"] = sd["d_h1_conv/w"] * 0.5
            m.store.load_state_dict(sd)
            assert m.store.vars["d_h1_conv/w"].packs_stale()
            m.train_step(img, z, use_graph=True)
            w = m.store.vars["d_h1_conv/w"]
            assert not w.packs_stale() and torch.equal(w._bf16, w.data.to(torch.bfloat16))
    l0, l1 = runs
    # Same kernels on the same bits, except that the filter-gradient kernels add their partial sums with fp32 reduce-adds
    # in CTA-finish order (last-bit differences that the GAN amplifies from step to step).  A stale filter copy would show
    # as an O(1) difference in step 0: g_loss_first is ~13.8 with the updated D, ~1 with the old one.
    assert abs(l1[0, 0] - l0[0, 0]) <= 1e-6 * abs(l0[0, 0]), (l0, l1)
    assert abs(l1[0, 1] - l0[0, 1]) <= 1e-3 * abs(l0[0, 1]), (l0, l1)
    np.testing.assert_allclose(l1, l0, rtol=0.15, atol=0.02)


@pytest.mark.parametrize("kind,B,H,C0,C1,C2,groups", [
    ("conv", 16, 32, 64, 128, 256, 2),      # d_h1 -> bn1 -> lrelu -> d_h2: the batch norm's dy comes from a conv_up launch
    ("conv", 8, 16, 128, 64, 128, 1),       # 64-channel batch norm: its dy comes from the class-concatenated conv_up (N = 256)
    ("conv", 6, 12, 64, 128, 64, 3),        # partial tiles (12 x 12 -> 6 x 6 -> 3 x 3), three row groups of two images
    ("deconv", 16, 4, 512, 256, 128, 1),    # g_h1 -> bn1 -> relu -> g_h2: dy comes from a conv_down launch
    ("deconv", 8, 16, 128, 64, 3, 1),       # g_h3 -> bn3 -> relu -> g_h4 (image side): dy comes from the warp-MMA c3m_down launch
    ("deconv", 3, 10, 64, 64, 3, 1),        # the same with partial 8 x 16 tiles (20 x 20 small grid)
])
def test_bn_backward_reductions_fused_into_dgrad(kind, B, H, C0, C1, C2, groups, monkeypatch):
    """gg_conv_dgrad_bnbwd: (sum g, sum g*xhat) of a train-mode batch norm accumulated in the epilogue of the dgrad launch
    that produces its dy (ops.FUSE_BN_BWD) must give the gradients of the two-pass path (colsum + apply) -- and both must
    match the float64 oracle of the same two-layer stack."""
    from gifgan import ops as _o
    rs = np.random.RandomState(B + H + C1)
    act = "lrelu" if kind == "conv" else "relu"

    def build(t, bn):
        if kind == "conv":
            h = _o.conv2d(t, C1, name="a", bn=bn, act=act, groups=groups)
            return _o.conv2d(h, C2, name="b", bias=False)
        h = _o.deconv2d(t, [B, 2 * H, 2 * H, C1], name="a", bn=bn, act=act, groups=groups)
        # (3 output channels: the image side -- fp32 image out, as g_h4 in the model)
        return _o.deconv2d(h, [B, 4 * H, 4 * H, C2], name="b", bias=False, out_dtype=torch.float32 if C2 == 3 else None)

    x = bf16_round(rs.randn(B, H, H, C0))
    results = []
    for fuse in (False, True):
        monkeypatch.setattr(_o, "FUSE_BN_BWD", fuse)
        bn = _o.batch_norm(name="bn")
        ops, st, tv = _store(lambda t: build(t, bn), (B, H, H, C0))
        if not results:
            sd = {k: v.clone() for k, v in st.state_dict().items()}
            sd["bn/gamma"] = torch.tensor(rs.uniform(0.5, 1.5, C1), dtype=torch.float32)
            sd["bn/beta"] = torch.tensor(rs.uniform(-0.3, 0.3, C1), dtype=torch.float32)
            for k in ("a/w", "b/w"):
                sd[k] = bf16_round(sd[k] * 2.5)
        st.load_state_dict(sd)
        xt = x.cuda().to(torch.bfloat16).requires_grad_(True)
        n0 = ops.cabi.launch_count()
        with ops.trainable(tv), ops.stats_arena():
            y = build(xt, bn)
            if not results:
                dy = bf16_round(rs.randn(*y.shape))
            y.backward(dy.cuda().to(y.dtype))
        results.append(dict(dx=xt.grad.float().cpu(), launches=ops.cabi.launch_count() - n0,
                            grads={k: st.vars[k].grad.clone().cpu() for k in ("a/w", "b/w", "bn/gamma", "bn/beta")}))
    two, one = results
    # the colsum launch is gone -- unless a tile would straddle two row groups (groups = 3: two images per group, four per
    # tile), where the library must fall back to the two-pass path on its own
    assert one["launches"] == two["launches"] - (0 if groups == 3 else 1), (one["launches"], two["launches"])
    assert relerr(one["dx"], two["dx"]) < 5e-3
    for k in two["grads"]:
        assert relerr(one["grads"][k], two["grads"][k]) < 5e-3, k
    # float64 oracle of the same stack
    xr = x.double().requires_grad_(True)
    wa, wb = sd["a/w"].double().requires_grad_(True), sd["b/w"].double().requires_grad_(True)
    ga, be = sd["bn/gamma"].double().requires_grad_(True), sd["bn/beta"].double().requires_grad_(True)
    if kind == "conv":
        pre = T.conv2d(xr, wa, sd["a/biases"].double())
    else:
        pre = T.conv2d_transpose(xr, wa, [B, 2 * H, 2 * H, C1]) + sd["a/biases"].double()
    parts = []
    for gidx in range(groups):
        p_ = pre[gidx * (B // groups):(gidx + 1) * (B // groups)]
        mu, var = p_.mean((0, 1, 2)), p_.var((0, 1, 2), unbiased=False)
        parts.append((p_ - mu) / torch.sqrt(var + 1e-5) * ga + be)
    h = torch.cat(parts, 0)
    h = T.lrelu(h) if act == "lrelu" else torch.relu(h)
    h = bf16_round(h.detach()).double() + (h - h.detach())          # the activation crosses the node boundary in bf16
    yr = T.conv2d(h, wb) if kind == "conv" else T.conv2d_transpose(h, wb, [B, 4 * H, 4 * H, C2])
    gx, gwa, gwb, gga, gbe = torch.autograd.grad(yr, [xr, wa, wb, ga, be], dy.double())
    # L2 metric: the activation gradient is re-rounded to bf16 between the layers (as in the model-level bf16 tests)
    l2 = lambda a, b: float((a.double() - b).norm() / b.norm())
    assert l2(one["dx"], gx) < 2e-2
    for k, want in (("a/w", gwa), ("b/w", gwb), ("bn/gamma", gga), ("bn/beta", gbe)):
        assert l2(one["grads"][k], want) < 2e-2, k


# wide linears (z_model_lib.py:160-161 gvideo_1 / gvideo_2: [clips*frames, 512] x [512, 512]) run as ONE-TAP tcgen05 GEMMs
@pytest.mark.parametrize("rows,ind,outd,act", [(512, 512, 512, "relu"), (200, 128, 192, None), (64, 64, 64, "lrelu")])
def test_tc_linear(rows, ind, outd, act):
    rs = np.random.RandomState(rows + ind)
    x, w, b = bf16_round(rs.randn(rows, ind)), bf16_round(rs.randn(ind, outd) * 0.05), torch.tensor(rs.randn(outd) * 0.1, dtype=torch.float32)
    from gifgan import ops as _o
    ops, st, tv = _store(lambda t: _o.linear(t, outd, "l", act=act), (rows, ind))
    st.load_state_dict({"l/Matrix": w.numpy(), "l/bias": b.numpy()})
    xt = x.cuda().to(torch.bfloat16).requires_grad_(True)
    assert ops._lin_tc(rows, ind, outd, xt)
    dy = bf16_round(rs.randn(rows, outd))
    L = __import__("gifgan._cabi", fromlist=["lib"]).lib()
    n0 = L.gg_launch_count()
    with ops.trainable(tv):
        y = ops.linear(xt, outd, "l", act=act)
        y.backward(dy.cuda().to(torch.bfloat16))
    ops.join_side()
    torch.cuda.synchronize()
    assert L.gg_launch_count() - n0 <= 6        # (bf16 cast of the Matrix,) fwd, (act'), bias grad, wgrad, dgrad: one tcgen05 launch each
    xr = x.double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = T.linear(xr, wr, br)
    if act == "relu":
        yr = torch.relu(yr)
    elif act == "lrelu":
        yr = T.lrelu(yr)
    gx, gw, gb = torch.autograd.grad(yr, [xr, wr, br], dy.double())
    assert y.dtype == torch.bfloat16 and tuple(y.shape) == (rows, outd)
    assert relerr(y, yr) < TOL
    assert relerr(xt.grad, gx) < TOL
    assert relerr(st.vars["l/Matrix"].grad, gw) < TOL
    assert relerr(st.vars["l/bias"].grad, gb) < TOL


def test_tc_linear_bn_chain_matches_simt_route():
    """Two 512-wide linear + batch norm + ReLU layers (gvideo_1 -> gvideo_2): the tcgen05 route -- statistics from the GEMM epilogue,
    g_bn1's backward reductions from gvideo_2's input-gradient launch -- against the SIMT route on the same bf16 inputs."""
    import gifgan.ops as _o
    rs = np.random.RandomState(11)
    rows = 512
    x = bf16_round(rs.randn(rows, 512))
    dy = bf16_round(rs.randn(rows, 512))
    res = {}
    for tc in (True, False):
        _o.TC_LINEAR = tc
        try:
            def build(t):
                h = _o.linear(t, 512, "gvideo_1", bn=_o.batch_norm(name="g_bn1"), act="relu")
                return _o.linear(h, 512, "gvideo_2", bn=_o.batch_norm(name="g_bn2"), act="relu")
            ops, st, tv = _store(build, (rows, 512))
            st.load_state_dict({"gvideo_1/Matrix": bf16_round(np.random.RandomState(1).randn(512, 512) * 0.05).numpy(),
                                "gvideo_2/Matrix": bf16_round(np.random.RandomState(2).randn(512, 512) * 0.05).numpy()}, strict=False)
            xt = x.cuda().to(torch.bfloat16).requires_grad_(True)
            bn1, bn2 = _o.batch_norm(name="g_bn1"), _o.batch_norm(name="g_bn2")
            with ops.trainable(tv):
                h = _o.linear(xt, 512, "gvideo_1", bn=bn1, act="relu")
                y = _o.linear(h, 512, "gvideo_2", bn=bn2, act="relu")
                y.backward(dy.cuda().to(torch.bfloat16))
            ops.join_side()
            torch.cuda.synchronize()
            res[tc] = dict(y=y.float().cpu(), gx=xt.grad.float().cpu(), g1=st.vars["gvideo_1/Matrix"].grad.clone().cpu(),
                           g2=st.vars["gvideo_2/Matrix"].grad.clone().cpu(), gg=st.vars["g_bn1/gamma"].grad.clone().cpu(),
                           gb=st.vars["g_bn1/beta"].grad.clone().cpu())
        finally:
            _o.TC_LINEAR = True
    for k in res[True]:
        a, b = res[True][k].double(), res[False][k].double()
        err = ((a - b).norm() / b.norm().clamp_min(1e-12)).item()
        assert err < TOL, (k, err)
