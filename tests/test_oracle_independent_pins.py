"""The oracle against INDEPENDENT restatements of TensorFlow semantics that happen to be installed here (the reference
itself pins nothing and TensorFlow cannot run: SURVEY 8c):

  * TF 'SAME' padding (App. A.1: pad_lo = total // 2, the extra pixel goes to the bottom / right) -- Hugging Face
    transformers ships its own port of TensorFlow's documented rule for its TF-converted MobileNet checkpoints
    (`apply_tf_padding`); the oracle's conv2d and the product's `same_pad` must agree with it;
  * sigmoid_cross_entropy_with_logits (App. A.5) -- torch's binary_cross_entropy_with_logits is the same function
    written by someone else.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import tf_ops as T


def _tf_pad():
    """Imported inside the test: collecting this file (e.g. for `-m gpu` runs) must not import transformers."""
    return pytest.importorskip("transformers.models.mobilenet_v2.modeling_mobilenet_v2").apply_tf_padding


class _Conv:                      # the attributes apply_tf_padding reads from an nn.Conv2d
    def __init__(self, k, s):
        self.kernel_size, self.stride, self.dilation = (k, k), (s, s), (1, 1)


@pytest.mark.parametrize("k,s", [(5, 2), (3, 2), (4, 2), (3, 1), (5, 3)])
def test_same_padding_against_the_transformers_port(k, s):
    from gifgan.ops import same_pad
    tf_pad = _tf_pad()
    rs = np.random.RandomState(k * 10 + s)
    for n in (64, 32, 28, 16, 14, 8, 7, 5, 3, 10):
        x = torch.tensor(rs.randn(2, n, n + 1, 3))                       # NHWC, non-square
        w = torch.tensor(rs.randn(k, k, 3, 4))
        want = F.conv2d(tf_pad(x.permute(0, 3, 1, 2), _Conv(k, s)), w.permute(3, 2, 0, 1), stride=s).permute(0, 2, 3, 1)
        got = T.conv2d(x, w, None, d_h=s, d_w=s)
        assert got.shape == want.shape and torch.allclose(got, want, rtol=0, atol=1e-12), (n, k, s)
        # the product's geometry helper: output size and the low-side pad
        padded = tf_pad(torch.ones(1, 1, n, n), _Conv(k, s))
        out, lo, hi = same_pad(n, k, s)
        assert out == -(-n // s) and padded.shape[-1] == n + lo + hi
        assert float(padded[0, 0, :lo].abs().sum()) == 0.0 and (lo == 0 or float(padded[0, 0, lo].sum()) == n), (n, k, s, lo, hi)
        assert hi - lo in (0, 1)


def test_sigmoid_cross_entropy_against_torch():
    rs = np.random.RandomState(0)
    x = torch.tensor(np.concatenate([rs.randn(50) * 4, [0.0, 40.0, -40.0, 1e-9]]))
    for z in (0.0, 1.0, 0.3):
        want = F.binary_cross_entropy_with_logits(x, torch.full_like(x, z), reduction="none")
        assert torch.allclose(T.sigmoid_cross_entropy_with_logits(x, torch.full_like(x, z)), want, rtol=1e-12, atol=1e-15)


def test_adam_recurrences_against_torch_adam_at_zero_epsilon():
    """TF Adam and torch.optim.Adam differ only in where epsilon sits (App. A.6: TF adds "epsilon hat" to sqrt(v) before the
    bias correction); with epsilon = 0 the two are the same recurrences -- moments, bias corrections, step."""
    rs = np.random.RandomState(1)
    p0 = rs.randn(64)
    grads = [rs.randn(64) + 0.1 for _ in range(6)]
    p = torch.tensor(p0.copy())
    opt = T.TFAdam({"p": p}, lr=2e-4, beta1=0.5, eps=0.0)
    q = torch.nn.Parameter(torch.tensor(p0.copy()))
    ref = torch.optim.Adam([q], lr=2e-4, betas=(0.5, 0.999), eps=0.0)
    for g in grads:
        opt.apply({"p": torch.tensor(g)})
        q.grad = torch.tensor(g)
        ref.step()
        assert torch.allclose(p, q.detach(), rtol=0, atol=1e-15)
    # and the documented difference: with epsilon > 0, TF's update equals torch's with eps / sqrt(1 - beta2^t)
    p = torch.tensor(p0.copy())
    opt = T.TFAdam({"p": p}, lr=2e-4, beta1=0.5, eps=1e-8)
    opt.apply({"p": torch.tensor(grads[0] * 1e-8)})                         # tiny gradient: epsilon matters
    g = grads[0] * 1e-8
    m, v = 0.5 * g, 0.001 * g * g
    want = p0 - 2e-4 * np.sqrt(1 - 0.999) / (1 - 0.5) * m / (np.sqrt(v) + 1e-8)
    assert np.allclose(p.numpy(), want, rtol=1e-12)
    torch_style = p0 - 2e-4 / (1 - 0.5) * m / (np.sqrt(v / (1 - 0.999)) + 1e-8)
    assert np.abs(want - torch_style).max() > 1e-5 * 2e-4                    # the two conventions really differ there


def test_lstm_cell_against_torch_lstmcell():
    """BasicLSTMCell (App. A.7): gates (i, j, f, o) of [x, h] @ Matrix + Bias with forget_bias 1 -- against
    torch.nn.LSTMCell (gates i, f, g, o; no forget bias) with the columns permuted and 1 added to the forget bias."""
    rs = np.random.RandomState(2)
    I, H, B = 12, 5, 3
    x, c, h = (torch.tensor(rs.randn(B, n)) for n in (I, H, H))
    M, b = torch.tensor(rs.randn(I + H, 4 * H) * 0.3), torch.tensor(rs.randn(4 * H) * 0.1)
    nc, nh = T.basic_lstm_cell(x, c, h, M, b)
    cell = torch.nn.LSTMCell(I, H).double()
    i_, j_, f_, o_ = (slice(k * H, (k + 1) * H) for k in range(4))
    order = [i_, f_, j_, o_]                                                 # torch: input, forget, cell (g), output
    with torch.no_grad():
        cell.weight_ih.copy_(torch.cat([M[:I, s].t() for s in order], 0))
        cell.weight_hh.copy_(torch.cat([M[I:, s].t() for s in order], 0))
        bias = torch.cat([b[i_], b[f_] + 1.0, b[j_], b[o_]])
        cell.bias_ih.copy_(bias)
        cell.bias_hh.zero_()
        th, tc = cell(x, (h, c))
    assert torch.allclose(nh, th, atol=1e-13) and torch.allclose(nc, tc, atol=1e-13)


def test_batch_norm_against_torch_batch_norm():
    """Normalisation with the BIASED batch variance (torch agrees); the moving variance is fed the biased variance in
    TF's contrib batch_norm (App. A.4) where torch feeds the unbiased one: the two running variances differ by n/(n-1)."""
    rs = np.random.RandomState(3)
    x = torch.tensor(rs.randn(4, 6, 6, 8) * 2 + 1)
    g, b = torch.tensor(rs.rand(8) + 0.5), torch.tensor(rs.randn(8))
    y, mm, mv = T.batch_norm_train(x, g, b, torch.zeros(8, dtype=torch.float64), torch.ones(8, dtype=torch.float64))
    rm, rv = torch.zeros(8, dtype=torch.float64), torch.ones(8, dtype=torch.float64)
    want = F.batch_norm(x.permute(0, 3, 1, 2), rm, rv, g, b, training=True, momentum=0.1, eps=1e-5).permute(0, 2, 3, 1)
    assert torch.allclose(y, want, atol=1e-12) and torch.allclose(mm, rm, atol=1e-14)
    n = 4 * 6 * 6
    assert torch.allclose((mv - 0.9) * n / (n - 1), rv - 0.9, atol=1e-14)
    z = T.batch_norm_infer(x, g, b, mm, mv)
    wz = F.batch_norm(x.permute(0, 3, 1, 2), mm, mv, g, b, training=False, eps=1e-5).permute(0, 2, 3, 1)
    assert torch.allclose(z, wz, atol=1e-12)
