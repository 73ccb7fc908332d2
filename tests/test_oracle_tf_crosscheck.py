"""Cross-check of the oracle's TensorFlow semantics against TensorFlow ITSELF -- runs only where `import tensorflow` succeeds
(SURVEY 8c-iv).  It does not run in this image (no TensorFlow wheel, no network: every test below is skipped, which is why the
oracle's header says "parity unpinned"); on any machine that has TensorFlow (1.x or 2.x) it pins, op by op, the conventions of
oracle/tf_ops.py that the reference relies on and that differ from PyTorch's defaults:

  * SAME padding of strided conv2d / conv3d (extra pad at the END) and conv2d_transpose as the input-gradient of that conv
    (/root/reference/models/recurrent_z/ops.py:51-100);
  * batch norm over all-but-last axes with the BIASED variance, and the moving-average update without zero-debias (ops.py:10-24);
  * tf.nn.sigmoid_cross_entropy_with_logits (model.py:121-126);
  * tf.train.AdamOptimizer: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t), epsilon OUTSIDE the square root (model.py:153-156);
  * BasicLSTMCell: gate order i, j, f, o and forget_bias = 1 (rnn_test/recurrent_DCGAN.py:199-200).

Inputs are seeded, float64 on the oracle side, float32 in TensorFlow; tolerance 1e-5 relative to the largest magnitude."""
import numpy as np
import pytest
import torch

tf = pytest.importorskip("tensorflow")

from oracle import tf_ops as T  # noqa: E402

TOL = 1e-5


def _np(t):
    return t.numpy() if hasattr(t, "numpy") else np.asarray(t)


def close(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return np.abs(got - want).max() <= TOL * max(1.0, np.abs(want).max())


def _eager():
    if not tf.executing_eagerly():                      # TensorFlow 1.x: run the comparisons eagerly as well
        tf.compat.v1.enable_eager_execution()


@pytest.mark.parametrize("H,W", [(64, 64), (7, 10), (28, 28)])
def test_conv2d_same_stride2(H, W):
    _eager()
    rs = np.random.RandomState(H)
    x, w, b = rs.randn(2, H, W, 3).astype(np.float32), rs.randn(5, 5, 3, 8).astype(np.float32), rs.randn(8).astype(np.float32)
    want = _np(tf.nn.bias_add(tf.nn.conv2d(x, w, strides=[1, 2, 2, 1], padding='SAME'), b))
    got = T.conv2d(torch.tensor(x).double(), torch.tensor(w).double(), torch.tensor(b).double())
    assert got.shape == want.shape and close(got.numpy(), want)


@pytest.mark.parametrize("h,w_", [(4, 4), (7, 5), (14, 14)])
def test_conv2d_transpose_same_stride2(h, w_):
    _eager()
    rs = np.random.RandomState(h)
    x, w = rs.randn(2, h, w_, 8).astype(np.float32), rs.randn(5, 5, 3, 8).astype(np.float32)      # filter [kh, kw, Cout, Cin]
    out_shape = [2, 2 * h, 2 * w_, 3]
    want = _np(tf.nn.conv2d_transpose(x, w, output_shape=out_shape, strides=[1, 2, 2, 1], padding='SAME'))
    got = T.conv2d_transpose(torch.tensor(x).double(), torch.tensor(w).double(), out_shape)
    assert close(got.numpy(), want)


@pytest.mark.parametrize("D,H", [(16, 8), (3, 5), (2, 1)])
def test_conv3d_same_stride2(D, H):
    _eager()
    rs = np.random.RandomState(D)
    x, w = rs.randn(2, D, H, H, 4).astype(np.float32), rs.randn(3, 3, 3, 4, 6).astype(np.float32)
    want = _np(tf.nn.conv3d(x, w, strides=[1, 2, 2, 2, 1], padding='SAME'))
    got = T.conv3d(torch.tensor(x).double(), torch.tensor(w).double())
    assert got.shape == want.shape and close(got.numpy(), want)


def test_batch_norm_moments_and_moving_average():
    _eager()
    rs = np.random.RandomState(1)
    x = (rs.randn(4, 6, 6, 5) * 2 + 1).astype(np.float32)
    gamma, beta = rs.uniform(0.5, 1.5, 5).astype(np.float32), rs.randn(5).astype(np.float32)
    mean, var = tf.nn.moments(tf.constant(x), axes=[0, 1, 2])
    want = _np(tf.nn.batch_normalization(x, mean, var, beta, gamma, 1e-5))
    mm, mv = np.zeros(5), np.ones(5)
    got, new_mm, new_mv = T.batch_norm_train(torch.tensor(x).double(), torch.tensor(gamma).double(), torch.tensor(beta).double(),
                                             torch.tensor(mm), torch.tensor(mv))
    assert close(got.numpy(), want)
    m, v = T.moments(torch.tensor(x).double())
    assert close(m.numpy(), _np(mean)) and close(v.numpy(), _np(var))           # biased variance
    # tf.contrib.layers.batch_norm(decay=0.9, zero_debias_moving_mean=False): assign_moving_average = moving - (moving - batch)*(1 - decay)
    assert close(new_mm.numpy(), mm - (mm - _np(mean)) * 0.1) and close(new_mv.numpy(), mv - (mv - _np(var)) * 0.1)


def test_sigmoid_cross_entropy_with_logits():
    _eager()
    x = np.array([-30.0, -2.5, -1e-3, 0.0, 1e-3, 2.5, 30.0], dtype=np.float32)
    for z in (0.0, 1.0):
        want = _np(tf.nn.sigmoid_cross_entropy_with_logits(labels=np.full_like(x, z), logits=x))
        got = T.sigmoid_cross_entropy_with_logits(torch.tensor(x).double(), torch.full((7,), z, dtype=torch.float64))
        assert close(got.numpy(), want)


def test_adam_optimizer_steps():
    _eager()
    rs = np.random.RandomState(2)
    p0 = rs.randn(6).astype(np.float32)
    grads = [rs.randn(6).astype(np.float32) * s for s in (1.0, 1e-3, 10.0, 1e-6)]
    var = tf.Variable(p0)
    opt = tf.compat.v1.train.AdamOptimizer(2e-4, beta1=0.5)
    ours = T.TFAdam({"p": torch.tensor(p0).double()}, 2e-4, 0.5)
    for g in grads:
        opt.apply_gradients([(tf.constant(g), var)])
        ours.apply({"p": torch.tensor(g).double()})
        assert np.abs(ours.params["p"].numpy() - _np(var)).max() <= 2e-7 * max(1.0, np.abs(p0).max())


def test_basic_lstm_cell_gate_order_and_forget_bias():
    _eager()
    rs = np.random.RandomState(3)
    B, I, H = 3, 5, 4
    x, c, h = rs.randn(B, I).astype(np.float32), rs.randn(B, H).astype(np.float32), rs.randn(B, H).astype(np.float32)
    cell = tf.compat.v1.nn.rnn_cell.BasicLSTMCell(H, forget_bias=1.0, state_is_tuple=True)
    out, (new_c, new_h) = cell(tf.constant(x), (tf.constant(c), tf.constant(h)))
    kernel, bias = [_np(v) for v in cell.weights]                                # [I + H, 4H], [4H]
    gc, gh = T.basic_lstm_cell(torch.tensor(x).double(), torch.tensor(c).double(), torch.tensor(h).double(),
                               torch.tensor(kernel).double(), torch.tensor(bias).double(), forget_bias=1.0)
    assert close(gc.numpy(), _np(new_c)) and close(gh.numpy(), _np(new_h)) and close(gh.numpy(), _np(out))
