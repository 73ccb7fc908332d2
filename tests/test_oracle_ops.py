"""Pins oracle/tf_ops.py (torch) against the independent numpy-loop restatement
oracle/kat.py, hand-computed known answers, identities and float64 finite
differences.  The reference has no tests of its own."""
import math

import numpy as np
import pytest
import torch

from oracle import kat
from oracle import tf_ops as T

torch.manual_seed(0)
f64 = torch.float64


def test_same_pad_table():
    # SURVEY App. A.1
    for n in (64, 32, 16, 8, 28, 14):
        assert T.same_pad(n, 5, 2) == (n // 2, 1, 2)
    for n in (16, 8, 4, 2):
        assert T.same_pad(n, 3, 2) == (n // 2, 0, 1)
    assert T.same_pad(7, 5, 2) == (4, 2, 2)


@pytest.mark.parametrize("H,W,ci,co", [(8, 8, 3, 4), (6, 10, 2, 5), (7, 5, 1, 2), (4, 4, 1, 1)])
def test_conv2d_vs_loops(H, W, ci, co):
    rs = np.random.RandomState(1)
    x, w, b = rs.randn(2, H, W, ci), rs.randn(5, 5, ci, co), rs.randn(co)
    y = T.conv2d(torch.tensor(x), torch.tensor(w), torch.tensor(b)).numpy()
    np.testing.assert_allclose(y, kat.conv2d(x, w, b), rtol=1e-12, atol=1e-12)


def test_conv2d_known_answer():
    # 1-channel 4x4 input of ones, all-ones 5x5 filter: output = number of valid taps.
    x = torch.ones(1, 4, 4, 1, dtype=f64)
    w = torch.ones(5, 5, 1, 1, dtype=f64)
    y = T.conv2d(x, w)[0, :, :, 0]
    # rows: p=0 covers input rows -1..3 -> 4 valid; p=1 covers 1..5 -> 3 valid
    assert y.tolist() == [[16.0, 12.0], [12.0, 9.0]]


@pytest.mark.parametrize("h,w_,ci,co", [(4, 4, 3, 2), (2, 3, 2, 4), (1, 1, 1, 1), (7, 7, 2, 3)])
def test_conv2d_transpose_vs_loops(h, w_, ci, co):
    rs = np.random.RandomState(2)
    x, w, b = rs.randn(2, h, w_, ci), rs.randn(5, 5, co, ci), rs.randn(co)
    y = T.conv2d_transpose(torch.tensor(x), torch.tensor(w), [2, 2 * h, 2 * w_, co], torch.tensor(b)).numpy()
    np.testing.assert_allclose(y, kat.conv2d_transpose(x, w, (2 * h, 2 * w_), b), rtol=1e-12, atol=1e-12)


def test_deconv_is_input_gradient_of_same_conv():
    # conv2d_transpose(x, w) == d/d(inp) <conv2d(inp, w), x>
    rs = np.random.RandomState(3)
    x = torch.tensor(rs.randn(2, 4, 4, 6))           # [B,h,w,Cin]
    w = torch.tensor(rs.randn(5, 5, 3, 6))           # [kh,kw,Cout,Cin] == HWIO with I=Cout, O=Cin
    inp = torch.zeros(2, 8, 8, 3, dtype=f64, requires_grad=True)
    (T.conv2d(inp, w) * x).sum().backward()
    y = T.conv2d_transpose(x, w, [2, 8, 8, 3])
    np.testing.assert_allclose(y.numpy(), inp.grad.numpy(), rtol=1e-12, atol=1e-12)
    # even output rows use kernel rows {1,3}, odd rows {0,2,4}
    x1 = torch.zeros(1, 2, 2, 1, dtype=f64); x1[0, 0, 0, 0] = 1
    wk = torch.arange(25, dtype=f64).reshape(5, 5, 1, 1)
    y1 = T.conv2d_transpose(x1, wk, [1, 4, 4, 1])[0, :, :, 0]
    assert y1[0, 0].item() == wk[1, 1, 0, 0].item() and y1[1, 1].item() == wk[2, 2, 0, 0].item()
    assert y1[3, 3].item() == wk[4, 4, 0, 0].item() and y1[2, 0].item() == wk[3, 1, 0, 0].item()


def test_conv3d_vs_loops():
    rs = np.random.RandomState(4)
    x, w, b = rs.randn(2, 4, 4, 2, 3), rs.randn(3, 3, 3, 3, 4), rs.randn(4)
    y = T.conv3d(torch.tensor(x), torch.tensor(w), torch.tensor(b)).numpy()
    np.testing.assert_allclose(y, kat.conv3d(x, w, b), rtol=1e-12, atol=1e-12)
    assert y.shape == (2, 2, 2, 1, 4)


def test_batch_norm_known_answer_and_ema():
    # two samples, one channel: x = {1, 3}: mean 2, biased var 1
    x = torch.tensor([[1.0], [3.0]], dtype=f64)
    g, b = torch.tensor([2.0], dtype=f64), torch.tensor([0.5], dtype=f64)
    mm, mv = torch.zeros(1, dtype=f64), torch.ones(1, dtype=f64)
    y, nmm, nmv = T.batch_norm_train(x, g, b, mm, mv)
    r = 1 / math.sqrt(1 + 1e-5)
    np.testing.assert_allclose(y.numpy().ravel(), [0.5 - 2 * r, 0.5 + 2 * r], rtol=1e-14)
    assert abs(nmm.item() - 0.2) < 1e-15 and abs(nmv.item() - 1.0) < 1e-15
    yi = T.batch_norm_infer(x, g, b, torch.tensor([2.0], dtype=f64), torch.tensor([1.0], dtype=f64))
    np.testing.assert_allclose(yi.numpy(), y.numpy(), rtol=1e-14)
    rs = np.random.RandomState(5)
    xx = rs.randn(3, 4, 4, 5)
    yk, mmk, mvk = kat.batch_norm_train(xx, np.ones(5) * 1.5, np.ones(5) * 0.1, np.zeros(5), np.ones(5))
    yt, mmt, mvt = T.batch_norm_train(torch.tensor(xx), torch.full((5,), 1.5, dtype=f64), torch.full((5,), 0.1, dtype=f64),
                                      torch.zeros(5, dtype=f64), torch.ones(5, dtype=f64))
    np.testing.assert_allclose(yt.numpy(), yk, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(mvt.numpy(), mvk, rtol=1e-12)
    np.testing.assert_allclose(T.batch_norm_plain(torch.tensor(xx)).numpy(),
                               kat.batch_norm_train(xx, 1.0, 0.0, 0, 1)[0], rtol=1e-11, atol=1e-12)


def test_lrelu_gradient_convention():
    x = torch.tensor([-1.0, 0.0, 2.0], dtype=f64, requires_grad=True)
    T.lrelu(x).sum().backward()
    assert x.grad.tolist() == [0.2, 1.0, 1.0]      # d/dx = 1 at x == 0 (TF MaximumGrad)
    x = torch.tensor([0.0], dtype=f64, requires_grad=True)
    T.relu(x).sum().backward()
    assert x.grad.item() == 0.0


def test_sigmoid_ce_known_answers():
    x = torch.tensor([0.0, 2.0, -3.0, 50.0, -50.0], dtype=f64)
    for z in (0.0, 1.0):
        got = T.sigmoid_cross_entropy_with_logits(x, torch.full_like(x, z)).numpy()
        np.testing.assert_allclose(got, kat.sigmoid_ce(x.numpy(), z), rtol=1e-14)
    assert abs(T.sigmoid_cross_entropy_with_logits(torch.tensor(0.0), torch.tensor(1.0)).item() - math.log(2)) < 1e-7
    # direct definition -z*log(s) - (1-z)*log(1-s)
    s = torch.sigmoid(x[:3])
    np.testing.assert_allclose(T.sigmoid_cross_entropy_with_logits(x[:3], torch.ones(3, dtype=f64)).numpy(), (-s.log()).numpy(), rtol=1e-12)


def test_tf_adam_first_steps():
    # t=1: lr_t = lr*sqrt(1-b2)/(1-b1); m=(1-b1)g; v=(1-b2)g^2
    p = torch.tensor([1.0, -2.0], dtype=f64)
    g = torch.tensor([1e-3, -4.0], dtype=f64)
    opt = T.TFAdam({"p": p})
    opt.apply({"p": g})
    lr_t = 2e-4 * math.sqrt(1 - 0.999) / 0.5
    want = np.array([1.0, -2.0]) - lr_t * (0.5 * g.numpy()) / (np.sqrt(0.001 * g.numpy() ** 2) + 1e-8)
    np.testing.assert_allclose(p.numpy(), want, rtol=1e-14)
    # epsilon is NOT bias-corrected: a tiny gradient moves much less than lr
    q = torch.tensor([0.0], dtype=f64)
    o2 = T.TFAdam({"q": q}); o2.apply({"q": torch.tensor([1e-10], dtype=f64)})
    assert abs(q.item()) < 0.3 * 2e-4
    # second step against the loop restatement
    pk, mk, vk = np.array([1.0, -2.0]), np.zeros(2), np.zeros(2)
    for t in (1, 2, 3):
        pk, mk, vk = kat.adam_step(pk, g.numpy(), mk, vk, t)
    opt.apply({"p": g}); opt.apply({"p": g})
    np.testing.assert_allclose(p.numpy(), pk, rtol=1e-13)


def test_lstm_gate_order_and_forget_bias():
    rs = np.random.RandomState(6)
    B, I, H = 3, 5, 4
    x, c, h = rs.randn(B, I), rs.randn(B, H), rs.randn(B, H)
    M, bias = rs.randn(I + H, 4 * H), rs.randn(4 * H)
    nc, nh = T.basic_lstm_cell(*(torch.tensor(a) for a in (x, c, h, M, bias)))
    kc, kh = kat.lstm_cell(x, c, h, M, bias)
    np.testing.assert_allclose(nc.numpy(), kc, rtol=1e-12)
    np.testing.assert_allclose(nh.numpy(), kh, rtol=1e-12)
    # zero weights: i=j=f=o=0 -> c' = c*sigmoid(1), h' = tanh(c')*0.5
    nc0, nh0 = T.basic_lstm_cell(torch.tensor(x), torch.tensor(c), torch.tensor(h), torch.zeros(I + H, 4 * H, dtype=f64), torch.zeros(4 * H, dtype=f64))
    np.testing.assert_allclose(nc0.numpy(), c / (1 + math.exp(-1.0)), rtol=1e-12)
    np.testing.assert_allclose(nh0.numpy(), np.tanh(nc0.numpy()) * 0.5, rtol=1e-12)


def test_finite_difference_gradients():
    rs = np.random.RandomState(7)
    x = torch.tensor(rs.randn(2, 4, 4, 2), requires_grad=True)
    w = torch.tensor(rs.randn(5, 5, 2, 3), requires_grad=True)
    wt = torch.tensor(rs.randn(5, 5, 3, 2), requires_grad=True)
    g = torch.tensor(rs.rand(3) + 0.5, requires_grad=True)
    b = torch.tensor(rs.randn(3), requires_grad=True)

    def f(x, w, wt, g, b):
        y = T.conv2d(x, w, b)
        y, _, _ = T.batch_norm_train(y, g, b, torch.zeros(3, dtype=f64), torch.ones(3, dtype=f64))
        y = T.conv2d_transpose(T.lrelu(y), wt.permute(0, 1, 3, 2), [2, 4, 4, 2])
        return torch.tanh(y)

    assert torch.autograd.gradcheck(f, (x, w, wt, g, b), eps=1e-6, atol=1e-6)


def test_get_std_and_cond_concat():
    x = torch.tensor([[1.0, 2.0], [3.0, 6.0]], dtype=f64)
    assert abs(T.get_std(x).item() - math.sqrt((1 + 4) / 2)) < 1e-12
    y = torch.tensor([[1.0, 0.0]], dtype=f64).reshape(1, 1, 1, 2)
    out = T.conv_cond_concat(torch.zeros(1, 2, 2, 1, dtype=f64), y)
    assert out.shape == (1, 2, 2, 3) and out[0, 1, 1].tolist() == [0.0, 1.0, 0.0]


def test_truncated_normal_bounds():
    w = T.truncated_normal(np.random.RandomState(0), (5, 5, 8, 16), 0.02)
    assert np.abs(w).max() <= 0.04 + 1e-9 and 0.015 < w.std() < 0.02
