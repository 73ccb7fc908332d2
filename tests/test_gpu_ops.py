"""GPU parity tests, op level: every kernel behind gifgan.ops (through the C ABI) against the CPU
oracle on the same seeded inputs.  Tolerances are the ones BASELINE.json's north_star states:
1e-4 relative in fp32 mode, 2e-2 in bf16 mode (relative to the largest reference magnitude)."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tf_ops as T  # noqa: E402

TOL = {"fp32": 1e-4, "bf16": 2e-2}


def relerr(got, want):
    got = got.detach().float().cpu().double()
    want = want.detach().double()
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-12)).item()


def _setup(precision, builder, in_shapes):
    """Trace `builder` on meta tensors (creates the variables), finalize the store."""
    from gifgan import ops
    ops.set_precision(precision)
    st = ops.reset_default_store(device="cuda", seed=3)
    builder(*[torch.empty(s, device="meta") for s in in_shapes])
    tv = [v for v in st.vars.values() if v.trainable]
    st.finalize(OrderedDict(all=tv))
    return ops, st, tv


def _cuda(a, precision, grad=True, keep_f32=False):
    t = torch.tensor(a, dtype=torch.float32, device="cuda")
    if precision == "bf16" and not keep_f32:
        t = t.to(torch.bfloat16)
    return t.requires_grad_(grad)


def _ref_in(a, precision, keep_f32=False):
    """The oracle sees the same (possibly bf16-rounded) input values."""
    t = torch.tensor(a, dtype=torch.float32)
    if precision == "bf16" and not keep_f32:
        t = t.to(torch.bfloat16).float()
    return t.double().requires_grad_(True)


CONV_CASES = [
    # B, H, W, Cin, Cout
    (2, 8, 8, 3, 8),
    (3, 16, 16, 64, 128),
    (2, 7, 5, 5, 6),       # ragged: odd sizes, channels not multiples of 4
    (1, 4, 4, 256, 64),
    (2, 28, 28, 11, 11),   # MNIST conditional d_h0_conv
]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("B,H,W,Ci,Co", CONV_CASES)
def test_conv2d_fwd_bwd(precision, B, H, W, Ci, Co):
    rs = np.random.RandomState(B * 1000 + H * 10 + Ci)
    x, w, b = rs.randn(B, H, W, Ci), rs.randn(5, 5, Ci, Co) * 0.05, rs.randn(Co) * 0.1
    Ho, Wo = -(-H // 2), -(-W // 2)
    dy = rs.randn(B, Ho, Wo, Co)
    ops, st, tv = _setup(precision, lambda t: ops_conv(t, Co), [(B, H, W, Ci)])
    st.load_state_dict({"c/w": w, "c/biases": b})
    xt = _cuda(x, precision)
    with ops.trainable(tv):
        y = ops.conv2d(xt, Co, name="c")
        y.backward(_cuda(dy, precision, grad=False))
    xr = _ref_in(x, precision)
    wr, br = torch.tensor(w, dtype=torch.float64, requires_grad=True), torch.tensor(b, dtype=torch.float64, requires_grad=True)
    if precision == "bf16":   # the kernel sees fp32 master weights (SIMT) -- identical values
        pass
    yr = T.conv2d(xr, wr.float().double(), br.float().double())
    dyr = _ref_in(dy, precision).detach()
    gx, gw, gb = torch.autograd.grad(yr, [xr, wr, br], dyr)
    tol = TOL[precision]
    assert y.shape == (B, Ho, Wo, Co)
    assert relerr(y, yr) < tol
    assert relerr(xt.grad, gx) < tol
    assert relerr(st.vars["c/w"].grad, gw) < tol
    assert relerr(st.vars["c/biases"].grad, gb) < tol


def ops_conv(t, Co):
    from gifgan import ops
    return ops.conv2d(t, Co, name="c")


DECONV_CASES = [(2, 4, 4, 8, 3), (3, 8, 8, 128, 64), (2, 3, 5, 6, 5), (1, 2, 2, 512, 256), (2, 7, 7, 138, 128)]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("B,h,w_,Ci,Co", DECONV_CASES)
def test_deconv2d_fwd_bwd(precision, B, h, w_, Ci, Co):
    rs = np.random.RandomState(B * 1000 + h * 10 + Ci)
    x, w, b = rs.randn(B, h, w_, Ci), rs.randn(5, 5, Co, Ci) * 0.05, rs.randn(Co) * 0.1
    dy = rs.randn(B, 2 * h, 2 * w_, Co)
    out_shape = [B, 2 * h, 2 * w_, Co]
    from gifgan import ops as _o
    ops, st, tv = _setup(precision, lambda t: _o.deconv2d(t, out_shape, name="g"), [(B, h, w_, Ci)])
    st.load_state_dict({"g/w": w, "g/biases": b})
    xt = _cuda(x, precision)
    with ops.trainable(tv):
        y = ops.deconv2d(xt, out_shape, name="g")
        y.backward(_cuda(dy, precision, grad=False))
    xr = _ref_in(x, precision)
    wr, br = torch.tensor(w, dtype=torch.float64, requires_grad=True), torch.tensor(b, dtype=torch.float64, requires_grad=True)
    yr = T.conv2d_transpose(xr, wr, out_shape, br)
    gx, gw, gb = torch.autograd.grad(yr, [xr, wr, br], _ref_in(dy, precision).detach())
    tol = TOL[precision]
    assert relerr(y, yr) < tol
    assert relerr(xt.grad, gx) < tol
    assert relerr(st.vars["g/w"].grad, gw) < tol
    assert relerr(st.vars["g/biases"].grad, gb) < tol


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("B,D,H,W,Ci,Co", [(2, 4, 4, 4, 8, 8), (2, 16, 8, 8, 64, 64), (1, 2, 1, 1, 64, 32), (2, 3, 5, 2, 3, 5)])
def test_conv3d_fwd_bwd(precision, B, D, H, W, Ci, Co):
    rs = np.random.RandomState(D * 100 + Ci)
    x, w, b = rs.randn(B, D, H, W, Ci), rs.randn(3, 3, 3, Ci, Co) * 0.05, rs.randn(Co) * 0.1
    from gifgan import ops as _o
    ops, st, tv = _setup(precision, lambda t: _o.conv3d(t, Co, name="v"), [(B, D, H, W, Ci)])
    st.load_state_dict({"v/w": w, "v/biases": b})
    xt = _cuda(x, precision)
    xr = _ref_in(x, precision)
    wr, br = torch.tensor(w, dtype=torch.float64, requires_grad=True), torch.tensor(b, dtype=torch.float64, requires_grad=True)
    yr = T.conv3d(xr, wr, br)
    dy = rs.randn(*yr.shape)
    with ops.trainable(tv):
        y = ops.conv3d(xt, Co, name="v")
        y.backward(_cuda(dy, precision, grad=False))
    gx, gw, gb = torch.autograd.grad(yr, [xr, wr, br], _ref_in(dy, precision).detach())
    tol = TOL[precision]
    assert tuple(y.shape) == tuple(yr.shape)
    assert relerr(y, yr) < tol
    assert relerr(xt.grad, gx) < tol
    assert relerr(st.vars["v/w"].grad, gw) < tol
    assert relerr(st.vars["v/biases"].grad, gb) < tol


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("rows,i,o", [(64, 100, 8192), (64, 8192, 1), (5, 121, 512), (7, 3636, 1024), (9, 1034, 1), (512, 512, 100)])
def test_linear_fwd_bwd(precision, rows, i, o):
    rs = np.random.RandomState(rows + i)
    x, M, b = rs.randn(rows, i), rs.randn(i, o) * 0.05, rs.randn(o) * 0.1
    dy = rs.randn(rows, o)
    from gifgan import ops as _o
    ops, st, tv = _setup(precision, lambda t: _o.linear(t, o, "l"), [(rows, i)])
    st.load_state_dict({"l/Matrix": M, "l/bias": b})
    xt = _cuda(x, precision)
    with ops.trainable(tv):
        y = ops.linear(xt, o, "l")
        y.backward(torch.tensor(dy, dtype=y.dtype, device="cuda"))
    xr = _ref_in(x, precision)
    Mr, br = torch.tensor(M, dtype=torch.float64, requires_grad=True), torch.tensor(b, dtype=torch.float64, requires_grad=True)
    yr = T.linear(xr, Mr, br)
    dyr = torch.tensor(dy, dtype=y.dtype).double()
    gx, gM, gb = torch.autograd.grad(yr, [xr, Mr, br], dyr)
    tol = TOL[precision]
    assert relerr(y, yr) < tol
    assert relerr(xt.grad, gx) < tol
    assert relerr(st.vars["l/Matrix"].grad, gM) < tol
    assert relerr(st.vars["l/bias"].grad, gb) < tol


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("shape,act,groups", [((4, 8, 8, 64), "lrelu", 1), ((6, 4, 4, 128), "relu", 2), ((64, 512), "relu", 1),
                                             ((3, 5, 5, 74), "lrelu", 1), ((2, 2, 2, 2, 256), "lrelu", 1), ((8, 16, 16, 6), None, 2)])
def test_batch_norm_train(precision, shape, act, groups):
    rs = np.random.RandomState(len(shape) + shape[-1])
    C = shape[-1]
    x = rs.randn(*shape) * 1.7 + 0.3
    gam, bet = rs.rand(C) + 0.5, rs.randn(C) * 0.2
    dy = rs.randn(*shape)
    from gifgan import ops as _o
    bn = _o.batch_norm(name="bn")
    ops, st, tv = _setup(precision, lambda t: bn(t, train=True, act=act, groups=groups), [shape])
    st.load_state_dict({"bn/gamma": gam, "bn/beta": bet, "bn/moving_mean": np.zeros(C), "bn/moving_variance": np.ones(C)})
    xt = _cuda(x, precision)
    with ops.trainable(tv):
        y = bn(xt, train=True, act=act, groups=groups)
        y.backward(_cuda(dy, precision, grad=False))
    xr = _ref_in(x, precision)
    g, b = torch.tensor(gam, dtype=torch.float64, requires_grad=True), torch.tensor(bet, dtype=torch.float64, requires_grad=True)
    mm, mv = torch.zeros(C, dtype=torch.float64), torch.ones(C, dtype=torch.float64)
    outs = []
    for xs in torch.chunk(xr, groups, 0):
        yy, mm, mv = T.batch_norm_train(xs, g, b, mm, mv)
        outs.append(yy)
    yr = torch.cat(outs, 0)
    yr = {"lrelu": T.lrelu, "relu": torch.relu, None: lambda t: t}[act](yr)
    gx, gg, gb = torch.autograd.grad(yr, [xr, g, b], _ref_in(dy, precision).detach())
    tol = TOL[precision]
    assert relerr(y, yr) < tol
    assert relerr(xt.grad, gx) < tol * (3 if precision == "bf16" else 1)
    assert relerr(st.vars["bn/gamma"].grad, gg) < tol
    assert relerr(st.vars["bn/beta"].grad, gb) < tol
    assert relerr(st.vars["bn/moving_mean"].data, mm) < 1e-4
    assert relerr(st.vars["bn/moving_variance"].data, mv) < 1e-4
    # inference mode uses the EMAs and leaves them untouched
    before = st.vars["bn/moving_mean"].data.clone()
    xi = _cuda(x, precision)
    with ops.trainable(tv):
        yi = bn(xi, train=False, act=act)
        yi.backward(_cuda(dy, precision, grad=False))
    xr2 = _ref_in(x, precision)
    yir = {"lrelu": T.lrelu, "relu": torch.relu, None: lambda t: t}[act](T.batch_norm_infer(xr2, g, b, mm, mv))
    (gxi,) = torch.autograd.grad(yir, [xr2], _ref_in(dy, precision).detach())
    assert relerr(yi, yir) < tol and relerr(xi.grad, gxi) < tol
    assert torch.equal(before, st.vars["bn/moving_mean"].data)


def test_batch_norm_plain_affine_free():
    """rnn_test variant: tf.nn.moments + tf.nn.batch_normalization(x, mean, var, None, None, 1e-5)."""
    rs = np.random.RandomState(5)
    x, dy = rs.randn(4, 8, 8, 64), rs.randn(4, 8, 8, 64)
    from gifgan import ops as _o
    bn = _o.batch_norm(name="p", affine=False, ema=False)
    ops, st, tv = _setup("fp32", lambda t: bn(t, act="relu"), [x.shape])
    assert not st.vars
    xt = _cuda(x, "fp32")
    y = bn(xt, act="relu")
    y.backward(_cuda(dy, "fp32", grad=False))
    xr = _ref_in(x, "fp32")
    yr = torch.relu(T.batch_norm_plain(xr))
    (gx,) = torch.autograd.grad(yr, [xr], torch.tensor(dy, dtype=torch.float64))
    assert relerr(y, yr) < 1e-4 and relerr(xt.grad, gx) < 1e-4


def test_activation_conventions():
    from gifgan import ops
    x = torch.tensor([-1.0, 0.0, 2.0, -0.0], device="cuda", requires_grad=True)
    ops.lrelu(x).sum().backward()
    assert x.grad.tolist() == [pytest.approx(0.2), 1.0, 1.0, 1.0]     # d/dx = 1 at x == 0
    x = torch.tensor([-1.0, 0.0, 2.0], device="cuda", requires_grad=True)
    ops.relu(x).sum().backward()
    assert x.grad.tolist() == [0.0, 0.0, 1.0]
    x = torch.linspace(-3, 3, 1001, device="cuda", requires_grad=True)
    y = ops.tanh(x)
    y.backward(torch.ones_like(y))
    assert relerr(y, torch.tanh(x.detach().cpu())) < 1e-6
    assert relerr(x.grad, 1 - torch.tanh(x.detach().cpu()) ** 2) < 1e-5


def test_sigmoid_ce_and_golden():
    import os
    from gifgan import ops
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ops.npz"))
    lg = torch.tensor(g["ce_logits"], dtype=torch.float32, device="cuda").requires_grad_(True)
    out = ops.sigmoid_cross_entropy_loss(lg, target=1.0)
    assert abs(out[0].item() - g["ce_ones"].mean()) < 1e-6 * max(1, abs(g["ce_ones"].mean()))
    out.backward(torch.ones_like(out))
    want = (torch.sigmoid(torch.tensor(g["ce_logits"])) - 1.0) / 16
    assert relerr(lg.grad, want) < 1e-5
    # two segments: d_loss = CE(real,1) + CE(fake,0)
    lg2 = torch.tensor(np.concatenate([g["ce_logits"], -g["ce_logits"]]), dtype=torch.float32, device="cuda")
    o2 = ops.sigmoid_cross_entropy_loss(lg2, [(0, 16, 1.0, 1.0), (16, 32, 0.0, 1.0)])
    assert abs(o2[0].item() - (o2[1].item() + o2[2].item())) < 1e-6
    assert abs(o2[1].item() - g["ce_ones"].mean()) < 1e-5 and abs(o2[2].item() - g["ce_ones"].mean()) < 1e-5
    # extreme logits stay finite
    big = torch.tensor([[80.0], [-80.0]], device="cuda")
    assert torch.isfinite(ops.sigmoid_cross_entropy_loss(big, target=0.0)).all()


def test_adam_matches_tf_semantics():
    from gifgan import ops
    rs = np.random.RandomState(9)
    n = 1003   # not a multiple of 4
    p0, grads = rs.randn(n).astype(np.float32), [rs.randn(n).astype(np.float32) * s for s in (1e-2, 1e-6, 3.0)]
    ops.set_precision("fp32")
    st = ops.reset_default_store(device="cuda")
    v = st.get_variable("p", [n], lambda r, s: p0)
    st.finalize(OrderedDict(grp=[v]))
    opt = ops.AdamOptimizer(st, "grp")
    ref = torch.tensor(p0, dtype=torch.float64)
    tf = T.TFAdam({"p": ref})
    for gnp in grads:
        v.grad.copy_(torch.tensor(gnp))
        opt.apply()
        tf.apply({"p": torch.tensor(gnp, dtype=torch.float64)})
        assert relerr(v.data, ref) < 2e-6
    assert opt.t == 3 and int(opt.state[0].item()) == 3
    # grad_scale = 1/world_size after a sum all-reduce
    v.grad.copy_(torch.tensor(grads[0]) * 4)
    opt.apply(grad_scale=0.25)
    tf.apply({"p": torch.tensor(grads[0], dtype=torch.float64)})
    assert relerr(v.data, ref) < 2e-6


def test_lstm_step_vs_oracle():
    from gifgan import _cabi as cabi
    rs = np.random.RandomState(4)
    B, I, H = 6, 40, 100
    x, c, h = rs.randn(B, I), rs.randn(B, H), rs.randn(B, H)
    M, bias = rs.randn(I + H, 4 * H) * 0.1, rs.randn(4 * H) * 0.1
    dh, dc = rs.randn(B, H), rs.randn(B, H)
    tt = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    xr, cr, hr, Mr, br = tt(x), tt(c), tt(h), tt(M), tt(bias)
    nc, nh = T.basic_lstm_cell(xr, cr, hr, Mr, br)
    gx, gc, gh, gM, gb = torch.autograd.grad([nh, nc], [xr, cr, hr, Mr, br], [torch.tensor(dh), torch.tensor(dc)])
    dev = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda").contiguous()
    gates_x = dev(x @ M[:I] + bias)
    Wh = dev(M[I:])
    c_out, h_out, gates = torch.empty(B, H, device="cuda"), torch.empty(B, H, device="cuda"), torch.empty(B, 4 * H, device="cuda")
    L = cabi.lib()
    cd, hd = dev(c), dev(h)      # keep alive: a freed temporary's block would be reused by the next allocation
    cabi.check(L.gg_lstm_step_fwd(gates_x.data_ptr(), Wh.data_ptr(), cd.data_ptr(), hd.data_ptr(), c_out.data_ptr(),
                                  h_out.data_ptr(), gates.data_ptr(), B, H, 1.0, cabi.stream()))
    assert relerr(c_out, nc) < 1e-5 and relerr(h_out, nh) < 1e-5
    dg, dcp, dhp = torch.empty(B, 4 * H, device="cuda"), torch.empty(B, H, device="cuda"), torch.empty(B, H, device="cuda")
    dhd, dcd = dev(dh), dev(dc)
    cabi.check(L.gg_lstm_step_bwd(gates.data_ptr(), cd.data_ptr(), c_out.data_ptr(), dhd.data_ptr(), dcd.data_ptr(), Wh.data_ptr(),
                                  dg.data_ptr(), dcp.data_ptr(), dhp.data_ptr(), B, H, 1.0, cabi.stream()))
    assert relerr(dcp, gc) < 1e-5 and relerr(dhp, gh) < 1e-5
    assert relerr(dg.cpu().double() @ torch.tensor(M[:I]).T, gx) < 1e-4          # dx = dgates @ Wx^T
    assert relerr(dg.sum(0), gb) < 1e-5


def test_get_std_and_errors():
    from gifgan import ops
    x = torch.randn(64, 8, 8, 256, device="cuda")
    want = T.get_std(x.cpu().double().reshape(64, -1))
    assert abs(ops.get_std(x).item() - want.item()) < 1e-5 * want.item()
    with pytest.raises(RuntimeError):          # no CPU fallback
        ops.lrelu(torch.zeros(4))
    with pytest.raises(TypeError):
        ops.lrelu(torch.zeros(4, device="cuda", dtype=torch.float16))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("rows,i,o,C", [(64, 100, 8192, 512), (128, 100, 2048, 128), (37, 110, 1024, 1024), (8, 100, 512, 128)])
def test_linear_bn_act_one_kernel_per_direction(precision, rows, i, o, C, monkeypatch):
    """gg_linear_bn_fwd / gg_linear_bn_bwd (the generator's input projection: thin linear + train-mode batch norm + ReLU,
    model.py:304-307) against the composed path (gg_linear_fwd + statistics + apply; colsum + bn_bwd_apply + thin_wgrad) and
    the float64 oracle: output, EMAs, and the gradients of Matrix, gamma, beta."""
    from gifgan import ops as _o
    rs = np.random.RandomState(rows + o)
    x, M, b = rs.uniform(-1, 1, (rows, i)), rs.randn(i, o) * 0.05, rs.randn(o) * 0.1
    gam, bet = rs.rand(C) + 0.5, rs.randn(C) * 0.2
    dy = rs.randn(rows, o)
    res = []
    for fuse in (False, True):
        monkeypatch.setattr(_o, "FUSE_LINEAR_BN", fuse)
        bn = _o.batch_norm(name="bn")
        ops, st, tv = _setup(precision, lambda t: _o.linear(t, o, "l", bn=bn, bn_channels=C, act="relu"), [(rows, i)])
        st.load_state_dict({"l/Matrix": M, "l/bias": b, "bn/gamma": gam, "bn/beta": bet, "bn/moving_mean": np.zeros(C), "bn/moving_variance": np.ones(C)})
        assert bool(ops.cabi.lib().gg_linear_bn_ok(rows, i, o, C, 1))
        xt = torch.tensor(x, dtype=torch.float32, device="cuda")                      # z stays fp32 (model.py feeds it as such)
        n0 = ops.cabi.launch_count()
        with ops.trainable(tv):
            y = ops.linear(xt, o, "l", bn=bn, bn_channels=C, act="relu")
            y.backward(_cuda(dy, precision, grad=False))
        res.append(dict(y=y.detach().float().cpu(), n=ops.cabi.launch_count() - n0, mm=st.vars["bn/moving_mean"].data.cpu().clone(),
                        mv=st.vars["bn/moving_variance"].data.cpu().clone(),
                        g={k: st.vars[k].grad.cpu().clone() for k in ("l/Matrix", "bn/gamma", "bn/beta", "l/bias")}))
    two, one = res
    assert one["n"] == 2 and two["n"] >= 5, (one["n"], two["n"])
    tol = TOL[precision]
    # float64 oracle
    xr = torch.tensor(x, dtype=torch.float64)
    Mr, br = torch.tensor(M, dtype=torch.float64, requires_grad=True), torch.tensor(b, dtype=torch.float64, requires_grad=True)
    gr, ber = torch.tensor(gam, dtype=torch.float64, requires_grad=True), torch.tensor(bet, dtype=torch.float64, requires_grad=True)
    pre = (xr @ Mr + br).reshape(-1, C)
    mu, var = pre.mean(0), pre.var(0, unbiased=False)
    yr = torch.relu((pre - mu) / torch.sqrt(var + 1e-5) * gr + ber).reshape(rows, o)
    dyr = torch.tensor(dy, dtype=y.dtype).double()
    gM, gg, gb = torch.autograd.grad(yr, [Mr, gr, ber], dyr)
    for r_ in (one, two):
        assert relerr(r_["y"], yr) < tol
        assert relerr(r_["mm"], 0.1 * mu.detach()) < max(tol, 1e-4) and relerr(r_["mv"], 0.9 + 0.1 * var.detach()) < max(tol, 1e-4)
        l2 = lambda a, w: float((a.double() - w).norm() / w.norm())
        assert (relerr(r_["g"]["l/Matrix"], gM) < tol) if precision == "fp32" else (l2(r_["g"]["l/Matrix"], gM) < tol)
        assert relerr(r_["g"]["bn/gamma"], gg) < tol and relerr(r_["g"]["bn/beta"], gb) < tol
        assert float(r_["g"]["l/bias"].abs().max()) == 0.0          # exactly zero through a train-mode batch norm: left untouched


# discriminator loss head (model.py:277 d_h3_lin + model.py:121-131 the cross-entropy means) as two launches, with the backward
# reductions of the batch norm that feeds it (d_bn3: channel = column % 512, one set per real / fake half)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("B,hw,C,groups", [(64, 4, 512, 2), (64, 4, 512, 1), (6, 2, 64, 2), (5, 1, 64, 1)])
def test_loss_head_fused(precision, B, hw, C, groups):
    from gifgan import ops as _o
    rows = B * groups
    rs = np.random.RandomState(B + C)
    xin = rs.randn(rows, hw, hw, 64).astype(np.float32)
    wc, wl = (rs.randn(1, 1, 64, C) * 0.1).astype(np.float32), (rs.randn(hw * hw * C, 1) * 0.05).astype(np.float32)
    ga, be = rs.uniform(0.5, 1.5, C).astype(np.float32), (rs.randn(C) * 0.1).astype(np.float32)
    segs = [(0, B, 1.0, 1.0), (B, 2 * B, 0.0, 1.0)] if groups == 2 else [(0, rows, 1.0, 0.7)]

    def run(fused):
        _o.FUSE_LOSS_HEAD = fused
        try:
            bnv = _o.batch_norm(name="d_bn3")

            def build(t):
                h = _o.conv2d(t, C, 1, 1, 1, 1, name="c", bn=bnv, act="lrelu", groups=groups)
                return _o.linear(_o.reshape(h, (rows, -1)), 1, "d_h3_lin")
            ops, st, tv = _setup(precision, build, [(rows, hw, hw, 64)])
            st.load_state_dict({"c/w": wc, "d_h3_lin/Matrix": wl, "d_h3_lin/bias": np.array([0.3], np.float32), "d_bn3/gamma": ga,
                                "d_bn3/beta": be}, strict=False)
            x = _cuda(xin, precision)
            L = ops.cabi.lib()
            with ops.trainable(tv), ops.stats_arena():
                h = _o.conv2d(x, C, 1, 1, 1, 1, name="c", bn=bnv, act="lrelu", groups=groups)
                n0 = L.gg_launch_count()
                logits = _o.linear(_o.reshape(h, (rows, -1)), 1, "d_h3_lin", ce_segments=segs)
                losses = _o.sigmoid_cross_entropy_loss(logits, segs)
                n_fwd = L.gg_launch_count() - n0
                torch.autograd.backward(losses, grad_tensors=torch.ones_like(losses))
            ops.join_side()
            torch.cuda.synchronize()
            n_bwd = L.gg_launch_count() - n0 - n_fwd
            g = {k: st.vars[k].grad.clone().cpu() for k in ("c/w", "d_h3_lin/Matrix", "d_h3_lin/bias", "d_bn3/gamma", "d_bn3/beta")}
            return logits.detach().cpu(), losses.detach().cpu(), x.grad.float().cpu(), g, (n_fwd, n_bwd)
        finally:
            _o.FUSE_LOSS_HEAD = True

    lf, pf, gxf, gf, nf = run(True)
    lc, pc, gxc, gc, nc = run(False)
    assert nf[0] == 1 and nc[0] == 1 + 2 * len(segs)    # one launch instead of linear + (cross-entropy, sum) per segment
    assert nf[1] == nc[1] - 3                           # Matrix gradient + bias gradient + dh + d_bn3's reduction pass -> one launch
    assert torch.equal(lf, lc) and torch.equal(pf, pc)  # same formula, same summation order: bit-identical logits and losses
    tol = TOL[precision]
    for k in gf:
        assert relerr(gf[k], gc[k]) < tol, k
    assert relerr(gxf, gxc) < tol
    # and against the float64 oracle (fp32 mode: no quantisation points to mirror)
    if precision == "fp32":
        xr = torch.tensor(xin).double().requires_grad_(True)
        wr, lr_ = torch.tensor(wc).double().requires_grad_(True), torch.tensor(wl).double().requires_grad_(True)
        gr, br = torch.tensor(ga).double().requires_grad_(True), torch.tensor(be).double().requires_grad_(True)
        pre = T.conv2d(xr, wr, None, 1, 1)
        rpg = rows // groups
        hs = []
        for g_ in range(groups):
            p_ = pre[g_ * rpg:(g_ + 1) * rpg]
            m_, v_ = p_.mean((0, 1, 2)), p_.var((0, 1, 2), unbiased=False)
            hs.append(T.lrelu((p_ - m_) / torch.sqrt(v_ + 1e-5) * gr + br))
        lg = T.linear(torch.cat(hs, 0).reshape(rows, -1), lr_, torch.tensor([0.3]).double())
        tot = sum(w_ * T.sigmoid_cross_entropy_with_logits(lg[a:b], torch.full_like(lg[a:b], t_)).mean() for a, b, t_, w_ in segs)
        gx, gw, gl, gg_, gb_ = torch.autograd.grad(tot, [xr, wr, lr_, gr, br])
        assert abs(pf[0].item() - tot.item()) < 1e-5 * max(1.0, abs(tot.item()))
        assert relerr(gxf, gx) < 1e-4 and relerr(gf["c/w"], gw) < 1e-4 and relerr(gf["d_h3_lin/Matrix"], gl) < 1e-4
        assert relerr(gf["d_bn3/gamma"], gg_) < 1e-4 and relerr(gf["d_bn3/beta"], gb_) < 1e-4
