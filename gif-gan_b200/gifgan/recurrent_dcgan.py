"""Recurrent image GAN -- the graph and train step of
/root/reference/models/recurrent_image/rnn_test/recurrent_DCGAN.py (module-level script there; a class here).

Generator (recurrent_DCGAN.py:170-229): per frame 4 x [conv 5x5 s2 (no bias) -> batch norm without gamma/beta (batch
statistics of THAT frame) -> ReLU] -> [B, 8192]; BasicLSTMCell(100) over the T frames (zero initial state);
per step  h_t @ W[100, 8192] + b -> [B,4,4,512] -> 4 x [BN -> ReLU -> conv2d_transpose] -> (tanh + 1) / 2.
Discriminator (recurrent_DCGAN.py:236-291): per frame 4 x [conv -> BN -> lrelu] -> FC 8192 -> 100; concat T frames ->
FC 100*T -> 1; applied to the generated frames and to the real next frames.  Losses: sigmoid CE; Adam(2e-4, 0.5);
step = d_optim, g_optim, g_optim (recurrent_DCGAN.py:353-375).

All T frames of a stage go through ONE kernel launch: frames are stacked on the batch axis ([T*B, ...]) and the
per-frame batch statistics are the `groups=T` mode of the batch-norm kernels.  The LSTM is one input-projection
GEMM over all T (x_t @ W_x, M = T*B) plus one fused launch per step (h @ W_h + gates + state update).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from . import ops
from .ops import batch_norm, check, dt, ptr, stream

cabi = ops.cabi


class _LSTM(torch.autograd.Function):
    """tf.nn.rnn(BasicLSTMCell(H, state_is_tuple=True), inputs, zero state) (recurrent_DCGAN.py:199-200; App. A.7).
    x: [T, B, I] fp32; Matrix [I+H, 4H] (rows [:I] multiply x_t, rows [I:] multiply h); returns h: [T, B, H]."""

    @staticmethod
    def forward(ctx, x, matrix, bias, mvar, bvar, H):
        T, B, I = x.shape
        L = cabi.lib()
        dev = x.device
        wx, wh = matrix[:I], matrix[I:]
        gx = torch.empty((T * B, 4 * H), dtype=torch.float32, device=dev)
        check(L.gg_linear_fwd(ptr(x.reshape(T * B, I)), dt(x), ptr(wx), ptr(bias), ptr(gx), 0, T * B, I, 4 * H, 0, 0.0, stream()), "gg_linear_fwd")
        gx = gx.view(T, B, 4 * H)
        c = torch.zeros((T + 1, B, H), dtype=torch.float32, device=dev)
        h = torch.zeros((T + 1, B, H), dtype=torch.float32, device=dev)
        gates = torch.empty((T, B, 4 * H), dtype=torch.float32, device=dev)
        for t in range(T):
            check(L.gg_lstm_step_fwd(ptr(gx[t]), ptr(wh), ptr(c[t]), ptr(h[t]), ptr(c[t + 1]), ptr(h[t + 1]), ptr(gates[t]), B, H, 1.0,
                                     stream()), "gg_lstm_step_fwd")
        ctx.mvar, ctx.bvar, ctx.H = mvar, bvar, H
        ctx.save_for_backward(x, matrix, c, h, gates)
        return h[1:]

    @staticmethod
    def backward(ctx, dh_out):
        x, matrix, c, h, gates = ctx.saved_tensors
        T, B, I = x.shape
        H = ctx.H
        L = cabi.lib()
        dev = x.device
        wx, wh = matrix[:I], matrix[I:]
        dh_out = dh_out.contiguous()
        dgates = torch.empty((T, B, 4 * H), dtype=torch.float32, device=dev)
        dc = torch.zeros((B, H), dtype=torch.float32, device=dev)
        dh_rec = torch.zeros((B, H), dtype=torch.float32, device=dev)
        dc_prev = torch.empty_like(dc)
        dh_prev = torch.empty_like(dc)
        dh_tot = torch.empty_like(dc)
        for t in range(T - 1, -1, -1):
            # total gradient wrt h_t = from the output at t + from step t+1
            check(L.gg_axpby(ptr(dh_out[t]), 1.0, ptr(dh_tot), 0.0, B * H, stream()), "gg_axpby")
            check(L.gg_axpby(ptr(dh_rec), 1.0, ptr(dh_tot), 1.0, B * H, stream()), "gg_axpby")
            check(L.gg_lstm_step_bwd(ptr(gates[t]), ptr(c[t]), ptr(c[t + 1]), ptr(dh_tot), ptr(dc), ptr(wh), ptr(dgates[t]), ptr(dc_prev),
                                     ptr(dh_prev), B, H, 1.0, stream()), "gg_lstm_step_bwd")
            dc, dc_prev = dc_prev, dc
            dh_rec, dh_prev = dh_prev, dh_rec
        dg2 = dgates.view(T * B, 4 * H)
        if ctx.needs_input_grad[1]:
            g = ctx.mvar.grad
            check(L.gg_linear_wgrad(ptr(x.reshape(T * B, I)), dt(x), ptr(dg2), 0, ptr(g[:I]), None, T * B, I, 4 * H, stream()), "gg_linear_wgrad")
            check(L.gg_linear_wgrad(ptr(h[:T].reshape(T * B, H)), 0, ptr(dg2), 0, ptr(g[I:]), None, T * B, H, 4 * H, stream()), "gg_linear_wgrad")
        if ctx.needs_input_grad[2]:
            check(L.gg_bias_grad(ptr(dg2), 0, ptr(ctx.bvar.grad), T * B, 4 * H, stream()), "gg_bias_grad")
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((T * B, I), dtype=x.dtype, device=dev)
            check(L.gg_linear_dgrad(ptr(dg2), 0, ptr(wx), ptr(dx), dt(dx), T * B, I, 4 * H, stream()), "gg_linear_dgrad")
            dx = dx.view(T, B, I)
        return dx, None, None, None, None, None


class RecurrentDCGAN(object):
    CH = [3, 64, 128, 256, 512]

    def __init__(self, batch_size=40, video_length=16, image_dimension=64, state_size=100, stddev=0.02,
                 learning_rate=0.0002, beta1=0.5, store=None, num_layers=1, shared_conv=False, output_keep_prob=1.0):
        """The two variants of the reference script are options here:
        num_layers=3: multi-layer_recurrent_DCGAN.py (:22, :206-207) -- MultiRNNCell of separately parameterised
          BasicLSTMCells (`lstm/Cell<k>/{Matrix,Bias}`); every layer is one _LSTM over the whole sequence.
        shared_conv=True, output_keep_prob=0.8: multi-layer_recurrent_DCGAN_with_shared_conv_and_drop_out.py -- the
          encoder runs the DISCRIMINATOR's conv filters and d_fc (:207, :213-215; LSTM input = state_size), so the D
          update also differentiates through the generator's encoder; the discriminator uses ReLU (:280, :300);
          DropoutWrapper(output_keep_prob) scales every cell's OUTPUT by a {0, 1/keep} mask (:219) while the recurrent
          state stays undropped.  `self.masks` ([layers, T, B, H]) fixes the draws (parity tests); None = draw per call."""
        self.B, self.T, self.S, self.H = batch_size, video_length, image_dimension, state_size
        self.L, self.shared_conv, self.keep, self.masks = int(num_layers), bool(shared_conv), float(output_keep_prob), None
        self.store = st = store if store is not None else ops.default_store()
        rn = ops.random_normal_initializer(stddev)
        zeros = ops.constant_initializer(0.0)
        CH = self.CH
        s16 = image_dimension // 16
        self.fc = fc = s16 * s16 * 512
        V = {}
        with ops.variable_scope("generator"):
            if not self.shared_conv:
                for i in range(4):
                    V[f"g_conv{i}"] = st.get_variable(f"conv_f{i + 1}", [5, 5, CH[i], CH[i + 1]], rn, filter_taps=25)
            lstm_in = state_size if self.shared_conv else fc
            with ops.variable_scope("lstm"):
                if self.L == 1:
                    V["lstm_m0"] = st.get_variable("Matrix", [lstm_in + state_size, 4 * state_size], rn)
                    V["lstm_b0"] = st.get_variable("Bias", [4 * state_size], zeros)
                else:
                    for k in range(self.L):
                        with ops.variable_scope("Cell%d" % k):
                            V[f"lstm_m{k}"] = st.get_variable("Matrix", [(lstm_in if k == 0 else state_size) + state_size, 4 * state_size], rn)
                            V[f"lstm_b{k}"] = st.get_variable("Bias", [4 * state_size], zeros)
            V["out_w"] = st.get_variable("output_fc_w", [state_size, fc], rn)
            V["out_b"] = st.get_variable("output_fc_bias", [1, fc], zeros)
            for i in range(4):
                V[f"g_deconv{i}"] = st.get_variable(f"deconv_f{i + 1}", [5, 5, CH[3 - i], CH[4 - i]], rn, filter_taps=25)
        with ops.variable_scope("discriminator"):
            for i in range(4):
                V[f"d_conv{i}"] = st.get_variable(f"d_conv_f{i + 1}", [5, 5, CH[i], CH[i + 1]], rn, filter_taps=25)
            V["d_fc_w"] = st.get_variable("d_fc_w", [fc, state_size], rn)
            V["d_fc_b"] = st.get_variable("d_fc_bias", [1, state_size], zeros)
            V["d_final_w"] = st.get_variable("d_final_fc_w", [state_size * video_length, 1], rn)
            V["d_final_b"] = st.get_variable("d_final_fc_bias", [1, 1], zeros)
        self.V = V
        self.g_vars = [v for v in st.vars.values() if v.name.startswith(st.scope_name() + "generator")]
        self.d_vars = [v for v in st.vars.values() if v.name.startswith(st.scope_name() + "discriminator")]
        st.finalize(OrderedDict(d=self.d_vars, g=self.g_vars))
        self.d_optim = ops.AdamOptimizer(st, "d", learning_rate, beta1)
        self.g_optim = ops.AdamOptimizer(st, "g", learning_rate, beta1)
        self.d_optim.var_list, self.g_optim.var_list = self.d_vars, self.g_vars
        self.bn = batch_norm(name="plain", affine=False, ema=False)      # tf.nn.moments + batch_normalization(None, None, 1e-5)
        self._ones = {}

    # ------------------------------------------------------------------------------
    def generator(self, X):
        """X: [T*B, S, S, 3] fp32 in [0,1) (frame-major).  Returns [T*B, S, S, 3]."""
        T, B, V = self.T, self.B, self.V
        x = X
        for i in range(4):
            x = ops.conv2d_v(x, V[f"d_conv{i}" if self.shared_conv else f"g_conv{i}"], bn=self.bn, act="relu", groups=T)
        if self.shared_conv:      # the encoder's fc layer is the discriminator's (…shared_conv_and_drop_out.py:213-215)
            enc = ops.linear_v(x.reshape(T * B, self.fc), V["d_fc_w"], V["d_fc_b"], out_dtype=torch.float32).reshape(T, B, self.H)
        else:
            enc = x.reshape(T, B, self.fc).float()
        h = enc
        for k in range(self.L):
            h = self._lstm(h, V[f"lstm_m{k}"], V[f"lstm_b{k}"])
            if self.keep < 1.0 and not ops._is_meta(h):
                h = h * self._mask(k)
        s16 = self.S // 16
        d = ops.linear_v(h.reshape(T * B, self.H), V["out_w"], V["out_b"], out_dtype=torch.float32).reshape(T * B, s16, s16, 512)
        for i in range(4):
            d = ops.bn_act(d, self.bn, act="relu", groups=T)
            sz = s16 * 2 ** (i + 1)
            last = i == 3
            d = ops.deconv2d_v(d, V[f"g_deconv{i}"], [T * B, sz, sz, self.CH[3 - i]], act="tanh01" if last else None,
                               out_dtype=torch.float32 if last else None)
        return d

    def _lstm(self, x, mvar, bvar):
        if ops._is_meta(x):
            return torch.empty((x.shape[0], x.shape[1], self.H), dtype=torch.float32, device="meta")
        return _LSTM.apply(x.contiguous(), ops._wtensor(mvar, ops._wants_grad(mvar)), ops._wtensor(bvar, ops._wants_grad(bvar)), mvar, bvar, self.H)

    def _mask(self, k):
        """DropoutWrapper's draw for layer k: [T, B, H] of {0, 1/keep} (tf.nn.dropout: floor(keep + uniform) / keep)."""
        if self.masks is not None:
            return torch.as_tensor(self.masks[k]).to(self.store.device, torch.float32)
        u = torch.rand((self.T, self.B, self.H), device=self.store.device)
        return torch.floor(self.keep + u) / self.keep

    def discriminator(self, frames):
        """frames: [T*B, S, S, 3] -> logits [B, 1]."""
        T, B, V = self.T, self.B, self.V
        x = frames
        for i in range(4):
            x = ops.conv2d_v(x, V[f"d_conv{i}"], bn=self.bn, act="relu" if self.shared_conv else "lrelu", groups=T)
        per = ops.linear_v(x.reshape(T * B, self.fc), V["d_fc_w"], V["d_fc_b"], out_dtype=torch.float32)      # [T*B, 100]
        cat = per.reshape(T, B, self.H).permute(1, 0, 2).reshape(B, T * self.H)                              # tf.concat(1, series)
        return ops.linear_v(cat, V["d_final_w"], V["d_final_b"])

    def _split(self, batch_input):
        """int32/uint8 [B, T+1, S, S, 3] -> X (frames 0..T-1) and Y (frames 1..T), /256, frame-major [T*B, S, S, 3]."""
        f = torch.as_tensor(batch_input).to(self.store.device).float() / 256
        X = f[:, : self.T].permute(1, 0, 2, 3, 4).reshape(self.T * self.B, self.S, self.S, 3).contiguous()
        Y = f[:, 1:].permute(1, 0, 2, 3, 4).reshape(self.T * self.B, self.S, self.S, 3).contiguous()
        return X, Y

    def _one(self, like):
        n = like.numel()
        if n not in self._ones:
            self._ones[n] = torch.ones(n, dtype=torch.float32, device=self.store.device)
        return self._ones[n]

    def update(self, batch_input, which, apply=True):
        """sess.run(d_optim) / sess.run(g_optim) of recurrent_DCGAN.py:353-375."""
        X, Y = self._split(batch_input)
        opt, var_list = (self.d_optim, self.d_vars) if which == "d" else (self.g_optim, self.g_vars)
        opt.zero_grad()
        B = self.B
        with ops.trainable(var_list), ops.overlap_wgrad():
            if which == "d":
                # shared encoder: d_loss reaches the discriminator's filters through the generator as well
                with torch.enable_grad() if self.shared_conv else torch.no_grad():
                    fake = self.generator(X)
                lf, lr = self.discriminator(fake), self.discriminator(Y)
                loss_f = ops.sigmoid_cross_entropy_loss(lf, target=0.0)
                loss_r = ops.sigmoid_cross_entropy_loss(lr, target=1.0)
                torch.autograd.backward([loss_f, loss_r], grad_tensors=[self._one(loss_f), self._one(loss_r)])
                d_loss, g_loss = loss_f[0] + loss_r[0], None
            else:
                fake = self.generator(X)
                lf = self.discriminator(fake)
                loss_g = ops.sigmoid_cross_entropy_loss(lf, target=1.0)
                torch.autograd.backward(loss_g, grad_tensors=self._one(loss_g))
                d_loss, g_loss = None, loss_g[0]
        if apply:
            opt.apply()
        return dict(d_loss=d_loss, g_loss=g_loss)

    def train_step(self, batch_input):
        d = self.update(batch_input, "d")
        self.update(batch_input, "g")
        g = self.update(batch_input, "g")
        return dict(d_loss=float(d["d_loss"]), g_loss=float(g["g_loss"]))
