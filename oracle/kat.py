"""CPU oracle -- independent numpy-loop restatement used to pin oracle/tf_ops.py.
TEST INFRASTRUCTURE ONLY (see oracle/tf_ops.py header; PARITY UNPINNED by the
reference, which has no tests).

Every function here is written straight from the defining sums in SURVEY.md
Appendix A with explicit Python loops over taps -- no library convolution -- so
that an indexing mistake in the torch-based oracle (padding side, kernel flip,
filter axis order) cannot hide.  Small shapes only.
"""
from __future__ import annotations

import numpy as np


def same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2, total - total // 2


def conv2d(x, w, b=None, s=2):
    """y[n,p,q,co] = sum_{r,t,ci} x[n, s*p + r - lo_h, s*q + t - lo_w, ci] * w[r,t,ci,co]
    (models/recurrent_z/ops.py:57, TF SAME)."""
    B, H, W, Ci = x.shape
    kh, kw, _, Co = w.shape
    Ho, lo_h, _ = same_pad(H, kh, s)
    Wo, lo_w, _ = same_pad(W, kw, s)
    y = np.zeros((B, Ho, Wo, Co), dtype=np.float64)
    for p in range(Ho):
        for q in range(Wo):
            for r in range(kh):
                for t in range(kw):
                    i, j = s * p + r - lo_h, s * q + t - lo_w
                    if 0 <= i < H and 0 <= j < W:
                        y[:, p, q, :] += x[:, i, j, :].astype(np.float64) @ w[r, t].astype(np.float64)
    if b is not None:
        y += b
    return y


def conv2d_transpose(x, w, out_hw, b=None, s=2):
    """y[n,i,j,co] = sum_{p,q,ci} x[n,p,q,ci] * w[i - s*p + lo_h, j - s*q + lo_w, co, ci]
    with (lo) the SAME padding of the forward conv on the *output* grid
    (models/recurrent_z/ops.py:86)."""
    B, h, wd, Ci = x.shape
    kh, kw, Co, _ = w.shape
    Ho, Wo = out_hw
    _, lo_h, _ = same_pad(Ho, kh, s)
    _, lo_w, _ = same_pad(Wo, kw, s)
    y = np.zeros((B, Ho, Wo, Co), dtype=np.float64)
    for p in range(h):
        for q in range(wd):
            for r in range(kh):
                for t in range(kw):
                    i, j = s * p + r - lo_h, s * q + t - lo_w
                    if 0 <= i < Ho and 0 <= j < Wo:
                        y[:, i, j, :] += x[:, p, q, :].astype(np.float64) @ w[r, t].astype(np.float64).T
    if b is not None:
        y += b
    return y


def conv3d(x, w, b=None, s=2):
    """models/recurrent_z/ops.py:70 -- NDHWC x DHWIO, stride s, SAME."""
    B, D, H, W, Ci = x.shape
    kd, kh, kw, _, Co = w.shape
    Do, lo_d, _ = same_pad(D, kd, s)
    Ho, lo_h, _ = same_pad(H, kh, s)
    Wo, lo_w, _ = same_pad(W, kw, s)
    y = np.zeros((B, Do, Ho, Wo, Co), dtype=np.float64)
    for o in range(Do):
        for p in range(Ho):
            for q in range(Wo):
                for u in range(kd):
                    for r in range(kh):
                        for t in range(kw):
                            d, i, j = s * o + u - lo_d, s * p + r - lo_h, s * q + t - lo_w
                            if 0 <= d < D and 0 <= i < H and 0 <= j < W:
                                y[:, o, p, q, :] += x[:, d, i, j, :].astype(np.float64) @ w[u, r, t].astype(np.float64)
    if b is not None:
        y += b
    return y


def batch_norm_train(x, gamma, beta, mm, mv, eps=1e-5, decay=0.9):
    xf = x.reshape(-1, x.shape[-1]).astype(np.float64)
    mean = xf.sum(0) / xf.shape[0]
    var = ((xf - mean) ** 2).sum(0) / xf.shape[0]
    y = (x - mean) / np.sqrt(var + eps) * gamma + beta
    return y, mm - (mm - mean) * (1 - decay), mv - (mv - var) * (1 - decay)


def sigmoid_ce(x, z):
    return np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))


def adam_step(p, g, m, v, t, lr=2e-4, b1=0.5, b2=0.999, eps=1e-8):
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    return p - lr_t * m / (np.sqrt(v) + eps), m, v


def lstm_cell(x, c, h, matrix, bias, forget_bias=1.0):
    sig = lambda a: 1.0 / (1.0 + np.exp(-a))
    H = c.shape[1]
    cat = np.concatenate([x, h], 1) @ matrix + bias
    i, j, f, o = cat[:, :H], cat[:, H:2 * H], cat[:, 2 * H:3 * H], cat[:, 3 * H:]
    nc = c * sig(f + forget_bias) + sig(i) * np.tanh(j)
    return nc, np.tanh(nc) * sig(o)
