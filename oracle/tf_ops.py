"""CPU oracle -- op layer.  TEST INFRASTRUCTURE ONLY.

A restatement, in PyTorch-CPU (float32 or float64), of the TensorFlow-0.12
arithmetic that gif-gan's hot path reaches through
``/root/reference/models/recurrent_z/ops.py`` and the rnn_test scripts.

PARITY UNPINNED: the reference ships no tests, golden vectors or checkpoints
 and TensorFlow cannot be installed in this image, so
nothing produced by the reference itself pins these functions.  They are
pinned instead by (i) the independent numpy-loop restatement in
``oracle/kat.py`` and hand-computed known-answer cases, (ii) float64
finite-difference gradient checks and (iii) algebraic identities
(deconv == input-gradient of the SAME conv) -- see tests/test_oracle_*.py.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product
(``gif-gan_b200/gifgan``) never does.

Conventions: activations NHWC / NDHWC, conv filters
HWIO ``[kh,kw,Cin,Cout]``, deconv filters ``[kh,kw,Cout,Cin]``, conv3d
``[kd,kh,kw,Cin,Cout]``, linear ``Matrix [in,out]``, LSTM ``Matrix
[in+H, 4H]`` with gate order i, j, f, o.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# SAME padding  (TF: out = ceil(n/s); total = max((out-1)*s + k - n, 0);
# lo = total//2, hi = total-lo).  Used by tf.nn.conv2d at ops.py:57,
# tf.nn.conv3d at ops.py:70 and (mirrored) tf.nn.conv2d_transpose at ops.py:86.
# --------------------------------------------------------------------------
def same_pad(n: int, k: int, s: int):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    lo = total // 2
    return out, lo, total - lo


def conv2d(x, w, b=None, d_h=2, d_w=2):
    """ops.py:51-62 -- tf.nn.conv2d(x, w, [1,d_h,d_w,1], 'SAME') + bias_add.

    x [B,H,W,Cin], w [kh,kw,Cin,Cout] -> [B,ceil(H/d_h),ceil(W/d_w),Cout].
    """
    kh, kw = w.shape[0], w.shape[1]
    _, pt, pb = same_pad(x.shape[1], kh, d_h)
    _, pl, pr = same_pad(x.shape[2], kw, d_w)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xn, w.permute(3, 2, 0, 1), stride=(d_h, d_w))
    y = y.permute(0, 2, 3, 1)
    if b is not None:
        y = y + b
    return y.contiguous()


def conv2d_transpose(x, w, output_shape, b=None, d_h=2, d_w=2):
    """ops.py:77-100 -- tf.nn.conv2d_transpose(x, w[kh,kw,Cout,Cin], output_shape,
    [1,d_h,d_w,1]) (padding SAME) + bias_add.

    Defined as the input-gradient of the SAME conv that maps
    [B,Ho,Wo,Cout] -> [B,h,w,Cin] with filter w viewed as HWIO (I=Cout, O=Cin):
      y[n,i,j,co] = sum_{p,q,ci} x[n,p,q,ci] * w[i - d_h*p + pad_lo_h, j - d_w*q + pad_lo_w, co, ci].
    """
    B, Ho, Wo, Cout = output_shape
    kh, kw = w.shape[0], w.shape[1]
    oh, pt, _ = same_pad(Ho, kh, d_h)
    ow, pl, _ = same_pad(Wo, kw, d_w)
    assert (oh, ow) == (x.shape[1], x.shape[2]), "output_shape inconsistent with input"
    assert w.shape[2] == Cout and w.shape[3] == x.shape[3]
    # conv_transpose2d weight is [Cin, Cout, kh, kw]
    full = F.conv_transpose2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), stride=(d_h, d_w))
    # full has size (h-1)*s + k; the SAME conv's padded input started pad_lo before pixel 0
    y = full[:, :, pt:pt + Ho, pl:pl + Wo]
    # rows past the end of `full` can only be needed when hi-pad rows exist (never for k=5,s=2)
    if y.shape[2] != Ho or y.shape[3] != Wo:
        y = F.pad(y, (0, Wo - y.shape[3], 0, Ho - y.shape[2]))
    y = y.permute(0, 2, 3, 1)
    if b is not None:
        y = y + b
    return y.contiguous()


def conv3d(x, w, b=None, d_d=2, d_h=2, d_w=2):
    """ops.py:64-75 -- tf.nn.conv3d(x, w[kd,kh,kw,Cin,Cout], [1,d_d,d_h,d_w,1], 'SAME') + bias."""
    kd, kh, kw = w.shape[:3]
    _, p0, p1 = same_pad(x.shape[1], kd, d_d)
    _, p2, p3 = same_pad(x.shape[2], kh, d_h)
    _, p4, p5 = same_pad(x.shape[3], kw, d_w)
    xn = F.pad(x.permute(0, 4, 1, 2, 3), (p4, p5, p2, p3, p0, p1))
    y = F.conv3d(xn, w.permute(4, 3, 0, 1, 2), stride=(d_d, d_h, d_w))
    y = y.permute(0, 2, 3, 4, 1)
    if b is not None:
        y = y + b
    return y.contiguous()


def linear(x, matrix, bias=None):
    """ops.py:106-117 -- tf.matmul(input_, Matrix) + bias."""
    y = x @ matrix
    if bias is not None:
        y = y + bias
    return y


# --------------------------------------------------------------------------
# activations
# --------------------------------------------------------------------------
class _LRelu(torch.autograd.Function):
    """ops.py:103-104 -- tf.maximum(x, leak*x).  TF's MaximumGrad routes the
    gradient to the first argument where x >= leak*x, i.e. d/dx = 1 at x == 0."""

    @staticmethod
    def forward(ctx, x, leak):
        ctx.save_for_backward(x)
        ctx.leak = leak
        return torch.maximum(x, leak * x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        first = x >= ctx.leak * x
        return torch.where(first, g, g * ctx.leak), None


def lrelu(x, leak=0.2):
    return _LRelu.apply(x, leak)


def relu(x):
    # tf.nn.relu: gradient 0 at x == 0 (same as torch)
    return torch.relu(x)


# --------------------------------------------------------------------------
# batch norm  
# --------------------------------------------------------------------------
def moments(x):
    """tf.nn.moments(x, axes=all-but-last): mean and *biased* variance."""
    axes = tuple(range(x.dim() - 1))
    mean = x.mean(dim=axes)
    var = ((x - mean) ** 2).mean(dim=axes)
    return mean, var


def batch_norm_train(x, gamma, beta, moving_mean, moving_var, eps=1e-5, decay=0.9):
    """ops.py:10-24 with train=True -- tf.contrib.layers.batch_norm(decay=0.9,
    updates_collections=None, epsilon=1e-5, scale=True, is_training=True).

    Returns y and the updated (moving_mean, moving_var); gradients flow through
    the batch statistics.  The EMA update uses the biased batch variance and
    no zero-debias:  mm -= (mm - mean) * (1 - decay).
    """
    mean, var = moments(x)
    y = (x - mean) * torch.rsqrt(var + eps) * gamma + beta
    with torch.no_grad():
        new_mm = moving_mean - (moving_mean - mean.detach()) * (1.0 - decay)
        new_mv = moving_var - (moving_var - var.detach()) * (1.0 - decay)
    return y, new_mm, new_mv


def batch_norm_infer(x, gamma, beta, moving_mean, moving_var, eps=1e-5):
    """ops.py:10-24 with train=False: y = (x - mm) * rsqrt(mv + eps) * gamma + beta."""
    return (x - moving_mean) * torch.rsqrt(moving_var + eps) * gamma + beta


def batch_norm_plain(x, eps=1e-5):
    """rnn_test/recurrent_DCGAN.py:190-191 -- tf.nn.moments(axes=[0,1,2]) +
    tf.nn.batch_normalization(x, mean, var, None, None, 1e-5): no gamma/beta, no EMA."""
    mean, var = moments(x)
    return (x - mean) * torch.rsqrt(var + eps)


# --------------------------------------------------------------------------
# losses, misc
# --------------------------------------------------------------------------
def sigmoid_cross_entropy_with_logits(logits, targets):
    """tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log(1+exp(-|x|))."""
    return torch.clamp(logits, min=0) - logits * targets + torch.log1p(torch.exp(-logits.abs()))


def get_std(x):
    """ops.py:125-128 -- sqrt(mean_over_features(var_over_batch))."""
    var = ((x - x.mean(dim=0)) ** 2).mean(dim=0)
    return torch.sqrt(var.mean())


def conv_cond_concat(x, y):
    """ops.py:45-49 -- concat y (broadcast over H,W) on the channel axis."""
    B, H, W, _ = x.shape
    return torch.cat([x, y * torch.ones(B, H, W, y.shape[3], dtype=x.dtype)], dim=3)


# --------------------------------------------------------------------------
# TF Adam -- NOT torch.optim.Adam
# --------------------------------------------------------------------------
class TFAdam:
    """tf.train.AdamOptimizer(lr, beta1).minimize: model.py:153-156.

    t starts at 1 on the first apply;  lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g^2;  p -= lr_t * m / (sqrt(v) + eps).
    """

    def __init__(self, params, lr=2e-4, beta1=0.5, beta2=0.999, eps=1e-8):
        self.params = dict(params)  # name -> tensor (updated in place)
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.m = {k: torch.zeros_like(v) for k, v in self.params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.params.items()}

    def lr_t(self, t=None):
        t = self.t if t is None else t
        return self.lr * math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)

    @torch.no_grad()
    def apply(self, grads):
        self.t += 1
        lr_t = self.lr_t()
        for k, p in self.params.items():
            g = grads[k]
            self.m[k].mul_(self.b1).add_(g, alpha=1.0 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
            p.sub_(lr_t * self.m[k] / (self.v[k].sqrt() + self.eps))


# --------------------------------------------------------------------------
# BasicLSTMCell
# --------------------------------------------------------------------------
def basic_lstm_cell(x, c, h, matrix, bias, forget_bias=1.0):
    """tf.nn.rnn_cell.BasicLSTMCell(H, state_is_tuple=True) as used at
    rnn_test/recurrent_DCGAN.py:199: concat=[x,h]@Matrix+Bias; i,j,f,o=split;
    c' = c*sigmoid(f+1) + sigmoid(i)*tanh(j);  h' = tanh(c')*sigmoid(o)."""
    concat = torch.cat([x, h], dim=1) @ matrix + bias
    i, j, f, o = torch.chunk(concat, 4, dim=1)
    new_c = c * torch.sigmoid(f + forget_bias) + torch.sigmoid(i) * torch.tanh(j)
    new_h = torch.tanh(new_c) * torch.sigmoid(o)
    return new_c, new_h


# --------------------------------------------------------------------------
# initialisers; deterministic from a numpy RandomState
# --------------------------------------------------------------------------
def truncated_normal(rs: np.random.RandomState, shape, stddev=0.02):
    """tf.truncated_normal_initializer: resample values beyond 2 sigma."""
    out = rs.normal(0.0, 1.0, size=shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rs.normal(0.0, 1.0, size=int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * stddev).astype(np.float32)


def random_normal(rs: np.random.RandomState, shape, stddev=0.02):
    return rs.normal(0.0, stddev, size=shape).astype(np.float32)
