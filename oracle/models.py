"""CPU oracle -- model graphs and train steps.  TEST INFRASTRUCTURE ONLY
(PARITY UNPINNED by the reference: see oracle/tf_ops.py header).

Restates, on top of oracle/tf_ops.py:
  * DCGAN                 /root/reference/models/recurrent_z/model.py:76-141, 268-389
  * its train step        model.py:226-243  (1 D-update, 2 G-updates with the same z)
  * VID_DCGAN             /root/reference/models/recurrent_z/z_model_lib.py:49-115, 353-418
  * its train step        z_model_lib.py:217-239
  * recurrent_DCGAN       /root/reference/models/recurrent_image/rnn_test/recurrent_DCGAN.py:159-307, 353-375

Variables are kept in a flat ``dict`` keyed by the TensorFlow variable names
 so that the product's state dict can be loaded 1:1.
"""
from __future__ import annotations

import numpy as np
import torch

from . import tf_ops as T


def _bn_vars(vars_, name, C, dtype):
    vars_[f"{name}/beta"] = torch.zeros(C, dtype=dtype)
    vars_[f"{name}/gamma"] = torch.ones(C, dtype=dtype)
    vars_[f"{name}/moving_mean"] = torch.zeros(C, dtype=dtype)
    vars_[f"{name}/moving_variance"] = torch.ones(C, dtype=dtype)


def _is_trainable(name):
    return not (name.endswith("moving_mean") or name.endswith("moving_variance"))


def _bf16(t):
    return t.to(torch.bfloat16).to(t.dtype)


class _RoundBoth(torch.autograd.Function):
    """value rounded to bf16 forward, gradient rounded to bf16 backward (a tensor the product stores in bf16)."""

    @staticmethod
    def forward(ctx, x):
        return _bf16(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


class _RoundGrad(torch.autograd.Function):
    """identity forward, gradient rounded to bf16 backward (fp32 pre-norm tensor whose gradient is a bf16 GEMM operand)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


class _RoundValue(torch.autograd.Function):
    """value rounded to bf16 forward, gradient untouched (the packed bf16 copy of an fp32 master filter)."""

    @staticmethod
    def forward(ctx, x):
        return _bf16(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _Graph:
    """Shared helpers: a variable store + the ops.py wrappers bound to it.

    `quant="bf16"` reproduces the product's bf16-mode quantisation points in the oracle (activations and their
    gradients stored in bf16, bf16 filter copies on the tensor-core layers, bf16 pre-norm gradients) so that the
    end-to-end bf16 parity test compares like with like: against a plain float64 oracle ~0.5 % of the
    ReLU/LeakyReLU masks flip under bf16 rounding noise, which by itself is a 5-10 % L2 gradient difference."""

    def __init__(self, dtype):
        self.dtype = dtype
        self.vars: dict[str, torch.Tensor] = {}
        self.prefix = ""
        self.trace = None  # when a dict, activations are recorded under their layer names
        self.quant = None

    def qa(self, x):
        return _RoundBoth.apply(x) if self.quant else x

    def qg(self, x, on=True):
        return _RoundGrad.apply(x) if (self.quant and on) else x

    def qw(self, w):
        C, K = w.shape[-2], w.shape[-1]
        # tensor-core layers (tcgen05: C, K multiples of 64) and the image-side layers (warp MMA: C == 3, K multiple of 64)
        return _RoundValue.apply(w) if (self.quant and (C % 64 == 0 or C == 3) and K % 64 == 0) else w

    def qimg(self, x):
        """the fp32 image is rounded to bf16 while it is staged as an MMA operand of d_h0_conv (conv_c3_mma.cu)"""
        return _RoundValue.apply(x) if (self.quant and x.shape[-1] == 3 and getattr(self, "df_dim", 0) % 64 == 0 and getattr(self, "df_dim", 0) > 0) else x

    def _v(self, name):
        return self.vars[self.prefix + name]

    def _rec(self, name, t):
        if self.trace is not None:
            self.trace[name] = t
            if t.requires_grad and not t.is_leaf:
                t.retain_grad()          # per-layer activation gradients for the parity diagnostics
        return t

    def bn(self, name, x, train):
        g, b = self._v(f"{name}/gamma"), self._v(f"{name}/beta")
        mm, mv = self._v(f"{name}/moving_mean"), self._v(f"{name}/moving_variance")
        if train:
            y, nmm, nmv = T.batch_norm_train(x, g, b, mm, mv)
            mm.copy_(nmm)
            mv.copy_(nmv)
            return y
        return T.batch_norm_infer(x, g, b, mm, mv)

    def conv2d(self, name, x):
        return self.qg(T.conv2d(x, self.qw(self._v(f"{name}/w")), self._v(f"{name}/biases")))

    def conv3d(self, name, x):
        return self.qg(T.conv3d(x, self.qw(self._v(f"{name}/w")), self._v(f"{name}/biases")))

    def deconv2d(self, name, x, out_shape, grad_round=True):
        return self.qg(T.conv2d_transpose(x, self.qw(self._v(f"{name}/w")), out_shape, self._v(f"{name}/biases")), grad_round)

    def linear(self, name, x, grad_round=False):
        return self.qg(T.linear(x, self._v(f"{name}/Matrix"), self._v(f"{name}/bias")), grad_round)

    def set_requires_grad(self, names):
        for k, v in self.vars.items():
            v.requires_grad_(k in names)
            v.grad = None

    def state_dict(self):
        return {k: v.detach().clone() for k, v in self.vars.items()}

    def load_state_dict(self, sd):
        with torch.no_grad():
            for k, v in sd.items():
                self.vars[k].copy_(torch.as_tensor(v, dtype=self.dtype))


# ==========================================================================
# DCGAN  (model.py)
# ==========================================================================
class DCGAN(_Graph):
    def __init__(self, batch_size=64, output_size=64, y_dim=None, z_dim=100, gf_dim=64,
                 df_dim=64, gfc_dim=1024, dfc_dim=1024, c_dim=3, seed=7,
                 dtype=torch.float32, lr=2e-4, beta1=0.5, vars_=None, prefix=""):
        super().__init__(dtype)
        self.batch_size, self.output_size = batch_size, output_size
        self.y_dim, self.z_dim, self.gf_dim, self.df_dim = y_dim, z_dim, gf_dim, df_dim
        self.gfc_dim, self.dfc_dim, self.c_dim = gfc_dim, dfc_dim, c_dim
        self.prefix = prefix
        if vars_ is not None:
            self.vars = vars_
        self._init_vars(np.random.RandomState(seed))
        names = [k for k in self.vars if k.startswith(prefix)]
        # model.py:138-139 -- split by substring of the variable name
        self.d_vars = [k for k in names if "d_" in k[len(prefix):] and _is_trainable(k)]
        self.g_vars = [k for k in names if "g_" in k[len(prefix):] and _is_trainable(k)]
        self.d_optim = T.TFAdam({k: self.vars[k] for k in self.d_vars}, lr, beta1)
        self.g_optim = T.TFAdam({k: self.vars[k] for k in self.g_vars}, lr, beta1)

    # ---- variables, initialisers of ops.py:55-59, 82-94, 110-113 ----------
    def _init_vars(self, rs):
        V, P, dt = self.vars, self.prefix, self.dtype
        tn = lambda shape: torch.tensor(T.truncated_normal(rs, shape), dtype=dt)
        rn = lambda shape: torch.tensor(T.random_normal(rs, shape), dtype=dt)
        z = lambda n: torch.zeros(n, dtype=dt)
        df, gf, c = self.df_dim, self.gf_dim, self.c_dim
        if not self.y_dim:
            s16 = self.output_size // 16
            for i, (ci, co) in enumerate([(c, df), (df, 2 * df), (2 * df, 4 * df), (4 * df, 8 * df)]):
                V[f"{P}d_h{i}_conv/w"], V[f"{P}d_h{i}_conv/biases"] = tn((5, 5, ci, co)), z(co)
            V[f"{P}d_h3_lin/Matrix"], V[f"{P}d_h3_lin/bias"] = rn((8 * df * s16 * s16, 1)), z(1)
            for i, C in [(1, 2 * df), (2, 4 * df), (3, 8 * df)]:
                _bn_vars(V, f"{P}d_bn{i}", C, dt)
            V[f"{P}g_h0_lin/Matrix"], V[f"{P}g_h0_lin/bias"] = rn((self.z_dim, 8 * gf * s16 * s16)), z(8 * gf * s16 * s16)
            _bn_vars(V, f"{P}g_bn0", 8 * gf, dt)
            for i, (ci, co) in enumerate([(8 * gf, 4 * gf), (4 * gf, 2 * gf), (2 * gf, gf), (gf, c)], start=1):
                V[f"{P}g_h{i}/w"], V[f"{P}g_h{i}/biases"] = rn((5, 5, co, ci)), z(co)
                if i < 4:
                    _bn_vars(V, f"{P}g_bn{i}", co, dt)
        else:
            y, s4 = self.y_dim, self.output_size // 4
            V[f"{P}d_h0_conv/w"], V[f"{P}d_h0_conv/biases"] = tn((5, 5, c + y, c + y)), z(c + y)
            V[f"{P}d_h1_conv/w"], V[f"{P}d_h1_conv/biases"] = tn((5, 5, c + 2 * y, df + y)), z(df + y)
            _bn_vars(V, f"{P}d_bn1", df + y, dt)
            V[f"{P}d_h2_lin/Matrix"], V[f"{P}d_h2_lin/bias"] = rn(((df + y) * s4 * s4 + y, self.dfc_dim)), z(self.dfc_dim)
            _bn_vars(V, f"{P}d_bn2", self.dfc_dim, dt)
            V[f"{P}d_h3_lin/Matrix"], V[f"{P}d_h3_lin/bias"] = rn((self.dfc_dim + y, 1)), z(1)
            V[f"{P}g_h0_lin/Matrix"], V[f"{P}g_h0_lin/bias"] = rn((self.z_dim + y, self.gfc_dim)), z(self.gfc_dim)
            _bn_vars(V, f"{P}g_bn0", self.gfc_dim, dt)
            V[f"{P}g_h1_lin/Matrix"], V[f"{P}g_h1_lin/bias"] = rn((self.gfc_dim + y, 2 * gf * s4 * s4)), z(2 * gf * s4 * s4)
            _bn_vars(V, f"{P}g_bn1", 2 * gf * s4 * s4, dt)
            V[f"{P}g_h2/w"], V[f"{P}g_h2/biases"] = rn((5, 5, 2 * gf, 2 * gf + y)), z(2 * gf)
            _bn_vars(V, f"{P}g_bn2", 2 * gf, dt)
            V[f"{P}g_h3/w"], V[f"{P}g_h3/biases"] = rn((5, 5, c, 2 * gf + y)), z(c)

    # ---- model.py:268-296 --------------------------------------------------
    def discriminator(self, image, y=None, train=True, tag="d"):
        B = image.shape[0]
        if not self.y_dim:
            h0 = self._rec(f"{tag}_h0", self.qa(T.lrelu(self._rec(f"{tag}_h0_conv", self.conv2d("d_h0_conv", self.qimg(image))))))
            h1 = self._rec(f"{tag}_h1", self.qa(T.lrelu(self.bn("d_bn1", self._rec(f"{tag}_h1_conv", self.conv2d("d_h1_conv", h0)), train))))
            h2 = self._rec(f"{tag}_h2", self.qa(T.lrelu(self.bn("d_bn2", self._rec(f"{tag}_h2_conv", self.conv2d("d_h2_conv", h1)), train))))
            h3 = self._rec(f"{tag}_h3", self.qa(T.lrelu(self.bn("d_bn3", self._rec(f"{tag}_h3_conv", self.conv2d("d_h3_conv", h2)), train))))
            h4 = self._rec(f"{tag}_logits", self.linear("d_h3_lin", h3.reshape(B, -1)))
            return torch.sigmoid(h4), h4, h2
        yb = y.reshape(B, 1, 1, self.y_dim)
        x = T.conv_cond_concat(image, yb)
        h0 = T.conv_cond_concat(T.lrelu(self.conv2d("d_h0_conv", x)), yb)
        # model.py:287,291 -- the y branch never passes train=: always batch statistics
        h1 = T.lrelu(self.bn("d_bn1", self.conv2d("d_h1_conv", h0), True))
        h1 = torch.cat([h1.reshape(B, -1), y], 1)
        h2 = torch.cat([T.lrelu(self.bn("d_bn2", self.linear("d_h2_lin", h1), True)), y], 1)
        h3 = self._rec(f"{tag}_logits", self.linear("d_h3_lin", h2))
        return torch.sigmoid(h3), h3, None

    # ---- model.py:298-344 (train=True) and 346-389 (sampler, train=False) ---
    def generator(self, z, y=None, train=True, tag="g"):
        B, s = z.shape[0], self.output_size
        gf = self.gf_dim
        if not self.y_dim:
            s2, s4, s8, s16 = s // 2, s // 4, s // 8, s // 16
            h0 = self._rec(f"{tag}_h0_lin", self.linear("g_h0_lin", z, grad_round=True)).reshape(-1, s16, s16, gf * 8)
            h0 = self._rec(f"{tag}_h0", self.qa(torch.relu(self.bn("g_bn0", h0, train))))
            h1 = self._rec(f"{tag}_h1_deconv", self.deconv2d("g_h1", h0, [B, s8, s8, gf * 4]))
            h1 = self._rec(f"{tag}_h1", self.qa(torch.relu(self.bn("g_bn1", h1, train))))
            h2 = self._rec(f"{tag}_h2_deconv", self.deconv2d("g_h2", h1, [B, s4, s4, gf * 2]))
            h2 = self._rec(f"{tag}_h2", self.qa(torch.relu(self.bn("g_bn2", h2, train))))
            h3 = self._rec(f"{tag}_h3_deconv", self.deconv2d("g_h3", h2, [B, s2, s2, gf]))
            h3 = self._rec(f"{tag}_h3", self.qa(torch.relu(self.bn("g_bn3", h3, train))))
            h4 = self._rec(f"{tag}_h4_deconv", self.deconv2d("g_h4", h3, [B, s, s, self.c_dim], grad_round=(self.c_dim == 3 and gf % 64 == 0)))
            return self._rec(f"{tag}_out", torch.tanh(h4))
        s2, s4 = s // 2, s // 4
        yb = y.reshape(B, 1, 1, self.y_dim)
        z = torch.cat([z, y], 1)
        # model.py:380 -- the sampler's g_bn0 is called without train=False: batch statistics
        h0 = torch.cat([torch.relu(self.bn("g_bn0", self.linear("g_h0_lin", z), True)), y], 1)
        h1 = torch.relu(self.bn("g_bn1", self.linear("g_h1_lin", h0), train)).reshape(B, s4, s4, gf * 2)
        h1 = T.conv_cond_concat(h1, yb)
        h2 = torch.relu(self.bn("g_bn2", self.deconv2d("g_h2", h1, [B, s2, s2, gf * 2]), train))
        h2 = T.conv_cond_concat(h2, yb)
        return self._rec(f"{tag}_out", torch.sigmoid(self.deconv2d("g_h3", h2, [B, s, s, self.c_dim])))

    def sampler(self, z, y=None):
        with torch.no_grad():
            return self.generator(z, y, train=False, tag="s")

    # ---- losses model.py:121-131 -------------------------------------------
    @staticmethod
    def _ce(logits, target):
        return T.sigmoid_cross_entropy_with_logits(logits, torch.full_like(logits, target)).mean()

    def d_update(self, images, z, y=None, apply=True):
        """sess.run(d_optim) at model.py:227: G fwd, D(real), D(fake), grads wrt d_vars."""
        self.set_requires_grad(set(self.d_vars))
        G = self.generator(z, y, train=True)
        _, logits, _ = self.discriminator(images, y, train=True, tag="d_real")
        _, logits_, _ = self.discriminator(G, y, train=True, tag="d_fake")
        d_loss_real, d_loss_fake = self._ce(logits, 1.0), self._ce(logits_, 0.0)
        d_loss = d_loss_real + d_loss_fake
        d_loss.backward()
        grads = {k: self.vars[k].grad.detach().clone() for k in self.d_vars}
        self.set_requires_grad(set())
        if apply:
            self.d_optim.apply(grads)
        return dict(d_loss=d_loss.item(), d_loss_real=d_loss_real.item(), d_loss_fake=d_loss_fake.item(), grads=grads)

    def g_update(self, z, y=None, apply=True):
        """sess.run(g_optim) at model.py:232: G fwd, D(fake), grads wrt g_vars."""
        self.set_requires_grad(set(self.g_vars))
        G = self.generator(z, y, train=True)
        if self.trace is not None:
            G.retain_grad()
        _, logits_, _ = self.discriminator(G, y, train=True, tag="d_fake")
        g_loss = self._ce(logits_, 1.0)
        g_loss.backward()
        grads = {k: self.vars[k].grad.detach().clone() for k in self.g_vars}
        out = dict(g_loss=g_loss.item(), grads=grads)
        if self.trace is not None:
            out["dG"] = G.grad.detach().clone()
        self.set_requires_grad(set())
        if apply:
            self.g_optim.apply(grads)
        return out

    def train_step(self, images, z, y=None, evals=False):
        """model.py:226-243: one D update, two G updates with the same z, and
        (evals=True) the three forward-only loss evaluations that also advance the BN EMAs."""
        d = self.d_update(images, z, y)
        g1 = self.g_update(z, y)
        g2 = self.g_update(z, y)
        out = dict(d_loss=d["d_loss"], g_loss=g2["g_loss"], g_loss_first=g1["g_loss"])
        if evals:
            with torch.no_grad():
                G = self.generator(z, y, train=True)
                out["errD_fake"] = self._ce(self.discriminator(G, y, train=True)[1], 0.0).item()
                out["errD_real"] = self._ce(self.discriminator(images, y, train=True)[1], 1.0).item()
                G = self.generator(z, y, train=True)
                out["errG"] = self._ce(self.discriminator(G, y, train=True)[1], 1.0).item()
        return out


# ==========================================================================
# VID_DCGAN  (z_model_lib.py)
# ==========================================================================
class VID_DCGAN(_Graph):
    def __init__(self, batch_size=32, z_input_size=120, z_output_size=100, vid_length=16,
                 output_image_size=64, c_dim=3, first_frame_loss_scalar=0.0,
                 train_img_gen=False, train_img_disc=False, seed=7, dtype=torch.float32,
                 lr=2e-4, beta1=0.5):
        super().__init__(dtype)
        self.batch_size, self.vid_length = batch_size, vid_length
        self.z_input_size, self.z_output_size = z_input_size, z_output_size
        self.first_frame_loss_scalar = first_frame_loss_scalar
        rs = np.random.RandomState(seed)
        V, dt = self.vars, dtype
        tn = lambda shape: torch.tensor(T.truncated_normal(rs, shape), dtype=dt)
        rn = lambda shape: torch.tensor(T.random_normal(rs, shape), dtype=dt)
        G = "video_gan/video_generator/"
        dims = [z_input_size + 1, 512, 512, 512, z_output_size]
        for i in range(4):
            V[f"{G}gvideo_{i}/Matrix"], V[f"{G}gvideo_{i}/bias"] = rn((dims[i], dims[i + 1])), torch.zeros(dims[i + 1], dtype=dt)
        for i in range(4):  # z_model_lib.py:36-39 -- gvideo_bn3 is constructed but never used
            if i < 3:
                _bn_vars(V, f"{G}gvideo_bn{i}", 512, dt)
        self.img_prefix = "video_gan/image_gan/"
        self.img_dcgan = DCGAN(batch_size * vid_length, output_image_size, z_dim=z_output_size, c_dim=c_dim,
                               seed=seed + 1, dtype=dt, vars_=self.vars, prefix=self.img_prefix)
        D = "video_gan/video_discriminator/"
        act_c = 4 * self.img_dcgan.df_dim
        for i, ci in [(1, act_c), (2, 256), (3, 256)]:
            V[f"{D}dvideo_h{i}/w"], V[f"{D}dvideo_h{i}/biases"] = tn((3, 3, 3, ci, 256)), torch.zeros(256, dtype=dt)
        for i in (2, 3):
            _bn_vars(V, f"{D}dvideo_bn{i}", 256, dt)
        sp = output_image_size // 8  # image-D h2 is [*, s/8, s/8, 256]
        flat = max(vid_length // 8, 1) * max(sp // 8, 1) ** 2 * 256
        V[f"{D}dvideo_h4/Matrix"], V[f"{D}dvideo_h4/bias"] = rn((flat, 1)), torch.zeros(1, dtype=dt)
        self.G, self.D = G, D
        names = list(self.vars)
        self.d_vid_vars = [k for k in names if "dvideo_" in k and _is_trainable(k)]
        self.g_vid_vars = [k for k in names if "gvideo_" in k and _is_trainable(k)]
        d_vars = self.d_vid_vars + (self.img_dcgan.d_vars if train_img_disc else [])
        g_vars = self.g_vid_vars + (self.img_dcgan.g_vars if train_img_gen else [])
        self.d_var_list, self.g_var_list = d_vars, g_vars
        self.d_optim = T.TFAdam({k: self.vars[k] for k in d_vars}, lr, beta1)
        self.g_optim = T.TFAdam({k: self.vars[k] for k in g_vars}, lr, beta1)

    # z_model_lib.py:353-382
    def generator(self, z, train=True):
        Bv, Tn = z.shape[0], self.vid_length
        z_copied = z[:, None, :].expand(Bv, Tn, z.shape[1])
        frame_numbers = torch.tensor(np.linspace(-1.0, 1.0, Tn), dtype=self.dtype)[None, :, None].expand(Bv, Tn, 1)
        h = torch.cat([z_copied, frame_numbers], 2).reshape(Bv * Tn, -1)
        self.prefix = self.G
        for i in range(3):
            h = self._rec(f"gr{i}", torch.relu(self.bn(f"gvideo_bn{i}", self.linear(f"gvideo_{i}", h), train)))
        out = self._rec("gr3", torch.tanh(self.linear("gvideo_3", h)))
        self.prefix = ""
        return out

    # z_model_lib.py:384-418 (batch norm always in train mode)
    def discriminator(self, act, tag="dv"):
        Bv = self.batch_size
        vid = act.reshape(Bv, self.vid_length, act.shape[1], act.shape[2], -1)
        self.prefix = self.D
        dr1 = self._rec(f"{tag}_dr1", T.lrelu(self.conv3d("dvideo_h1", vid)))
        dr2 = self._rec(f"{tag}_dr2", T.lrelu(self.bn("dvideo_bn2", self.conv3d("dvideo_h2", dr1), True)))
        dr3 = self._rec(f"{tag}_dr3", T.lrelu(self.bn("dvideo_bn3", self.conv3d("dvideo_h3", dr2), True)))
        d4 = self._rec(f"{tag}_logits", self.linear("dvideo_h4", dr3.reshape(Bv, -1)))
        self.prefix = ""
        return torch.sigmoid(d4), d4

    def _fake_logits(self, z, train_gvideo=True):
        G_out = self.generator(z, train=train_gvideo)
        img = self.img_dcgan
        frames = img.generator(G_out, train=False, tag="s")          # img_dcgan.sampler(G_out), z_model_lib.py:76 + model.py:111
        _, _, act = img.discriminator(frames, train=False, tag="dinf_fake")  # D_activations_inf_
        return self.discriminator(act, "dv_fake")[1], G_out, frames

    def d_update(self, images, z, apply=True):
        """z_model_lib.py:220-229 with is_training=True."""
        self.set_requires_grad(set(self.d_var_list))
        _, _, act_real = self.img_dcgan.discriminator(images, train=False, tag="dinf_real")
        real_logits = self.discriminator(act_real, "dv_real")[1]
        fake_logits, _, frames = self._fake_logits(z)
        d_loss = DCGAN._ce(real_logits, 1.0) + DCGAN._ce(fake_logits, 0.0)
        d_loss.backward()
        grads = {k: self.vars[k].grad.detach().clone() for k in self.d_var_list}
        self.set_requires_grad(set())
        if apply:
            self.d_optim.apply(grads)
        return dict(d_loss=d_loss.item(), grads=grads,
                    images_std=T.get_std(images).item(), sampler_std=T.get_std(frames.detach()).item())

    def g_update(self, z, apply=True):
        """z_model_lib.py:233-239."""
        self.set_requires_grad(set(self.g_var_list))
        fake_logits, G_out, _ = self._fake_logits(z)
        first = G_out[:: self.vid_length, :]
        g_first = self.first_frame_loss_scalar * ((first - z[:, : self.z_output_size]) ** 2).mean()
        g_loss = DCGAN._ce(fake_logits, 1.0) + g_first
        g_loss.backward()
        grads = {k: self.vars[k].grad.detach().clone() for k in self.g_var_list}
        self.set_requires_grad(set())
        if apply:
            self.g_optim.apply(grads)
        return dict(g_loss=g_loss.item(), g_loss_first_frame=float(g_first.detach()) if torch.is_tensor(g_first) else float(g_first), grads=grads)

    def train_step(self, images, z, disc_updates=1, gen_updates=2):
        d = [self.d_update(images, z) for _ in range(disc_updates)]
        g = [self.g_update(z) for _ in range(gen_updates)]
        return dict(d_loss=d[-1]["d_loss"], g_loss=g[-1]["g_loss"])

    def sample(self, z):
        """model_sampler.py:66-70: gvideo (inference BN) -> image sampler."""
        with torch.no_grad():
            return self.img_dcgan.generator(self.generator(z, train=False), train=False, tag="s")


# ==========================================================================
# recurrent_DCGAN  (rnn_test/recurrent_DCGAN.py)
# ==========================================================================
class RecurrentDCGAN(_Graph):
    CH = [3, 64, 128, 256, 512]

    def __init__(self, batch_size=40, video_length=16, image_dimension=64, state_size=100,
                 seed=7, dtype=torch.float32, lr=2e-4, beta1=0.5, num_layers=1, shared_conv=False, output_keep_prob=1.0):
        """num_layers > 1: multi-layer_recurrent_DCGAN.py:22,206-207 (MultiRNNCell of separate BasicLSTMCells).
        shared_conv: multi-layer_recurrent_DCGAN_with_shared_conv_and_drop_out.py:169-180,207,213-215,280 -- the encoder
        runs the DISCRIMINATOR's conv filters and its d_fc (LSTM input = state_size), the discriminator uses ReLU.
        output_keep_prob < 1: DropoutWrapper on every cell's output (:219), masks supplied by the caller."""
        super().__init__(dtype)
        self.B, self.Tn, self.S, self.H = batch_size, video_length, image_dimension, state_size
        self.L, self.shared_conv, self.keep = num_layers, shared_conv, output_keep_prob
        rs = np.random.RandomState(seed)
        rn = lambda shape: torch.tensor(T.random_normal(rs, shape), dtype=dtype)
        V, CH = self.vars, self.CH
        s16 = image_dimension // 16
        self.fc = fc = s16 * s16 * 512
        if not shared_conv:
            for i in range(4):  # recurrent_DCGAN.py:177-180
                V[f"generator/conv_f{i+1}"] = rn((5, 5, CH[i], CH[i + 1]))
        lstm_in = state_size if shared_conv else fc
        if num_layers == 1:
            V["generator/lstm/Matrix"] = rn((lstm_in + state_size, 4 * state_size))
            V["generator/lstm/Bias"] = torch.zeros(4 * state_size, dtype=dtype)
        else:
            for k in range(num_layers):
                V[f"generator/lstm/Cell{k}/Matrix"] = rn(((lstm_in if k == 0 else state_size) + state_size, 4 * state_size))
                V[f"generator/lstm/Cell{k}/Bias"] = torch.zeros(4 * state_size, dtype=dtype)
        V["generator/output_fc_w"] = rn((state_size, fc))
        V["generator/output_fc_bias"] = torch.zeros(1, fc, dtype=dtype)
        for i in range(4):  # recurrent_DCGAN.py:205-208: [kh,kw,Cout,Cin]
            V[f"generator/deconv_f{i+1}"] = rn((5, 5, CH[3 - i], CH[4 - i]))
        for i in range(4):  # recurrent_DCGAN.py:239-242
            V[f"discriminator/d_conv_f{i+1}"] = rn((5, 5, CH[i], CH[i + 1]))
        V["discriminator/d_fc_w"] = rn((fc, state_size))
        V["discriminator/d_fc_bias"] = torch.zeros(1, state_size, dtype=dtype)
        V["discriminator/d_final_fc_w"] = rn((state_size * video_length, 1))
        V["discriminator/d_final_fc_bias"] = torch.zeros(1, 1, dtype=dtype)
        self.g_vars = [k for k in V if k.startswith("generator")]
        self.d_vars = [k for k in V if k.startswith("discriminator")]
        self.g_optim = T.TFAdam({k: V[k] for k in self.g_vars}, lr, beta1)
        self.d_optim = T.TFAdam({k: V[k] for k in self.d_vars}, lr, beta1)
        self.masks = None      # [L, T, B, H] of {0, 1/keep}: the DropoutWrapper draws, fixed by the caller for parity

    def _cell(self, k):
        V = self.vars
        if self.L == 1:
            return V["generator/lstm/Matrix"], V["generator/lstm/Bias"]
        return V[f"generator/lstm/Cell{k}/Matrix"], V[f"generator/lstm/Cell{k}/Bias"]

    def generator(self, X):
        """X: list of T tensors [B,S,S,3] in [0,1).  recurrent_DCGAN.py:170-225."""
        V, B = self.vars, self.B
        s16 = self.S // 16
        enc = []
        for x in X:
            for i in range(4):
                w = V[f"discriminator/d_conv_f{i+1}"] if self.shared_conv else V[f"generator/conv_f{i+1}"]
                x = torch.relu(T.batch_norm_plain(T.conv2d(x, w)))
            x = x.reshape(B, self.fc)
            if self.shared_conv:
                x = x @ V["discriminator/d_fc_w"] + V["discriminator/d_fc_bias"]
            enc.append(x)
        state = [(torch.zeros(B, self.H, dtype=self.dtype), torch.zeros(B, self.H, dtype=self.dtype)) for _ in range(self.L)]
        outs = []
        for t, e in enumerate(enc):
            inp = e
            for k in range(self.L):      # MultiRNNCell: layer k's (dropped) output feeds layer k+1; states stay undropped
                m, b = self._cell(k)
                c, h = T.basic_lstm_cell(inp, state[k][0], state[k][1], m, b)
                state[k] = (c, h)
                inp = h if self.keep >= 1.0 else h * self.masks[k, t].to(self.dtype)
            self._rec(f"lstm_h{t}", inp)
            d = (inp @ V["generator/output_fc_w"] + V["generator/output_fc_bias"]).reshape(B, s16, s16, 512)
            for i in range(4):
                d = torch.relu(T.batch_norm_plain(d))
                sz = s16 * 2 ** (i + 1)
                d = T.conv2d_transpose(d, V[f"generator/deconv_f{i+1}"], [B, sz, sz, self.CH[3 - i]])
            outs.append((torch.tanh(d) + 1) / 2)
        return outs

    def discriminator(self, frames):
        """recurrent_DCGAN.py:249-266."""
        V, B = self.vars, self.B
        per = []
        for x in frames:
            for i in range(4):
                x = T.batch_norm_plain(T.conv2d(x, V[f"discriminator/d_conv_f{i+1}"]))
                x = torch.relu(x) if self.shared_conv else T.lrelu(x)
            per.append(x.reshape(B, self.fc) @ V["discriminator/d_fc_w"] + V["discriminator/d_fc_bias"])
        return torch.cat(per, 1) @ V["discriminator/d_final_fc_w"] + V["discriminator/d_final_fc_bias"]

    def _split(self, batch_input):
        f = batch_input.to(self.dtype) / 256
        return [f[:, t] for t in range(self.Tn)], [f[:, t + 1] for t in range(self.Tn)]

    def losses(self, batch_input):
        X, Y = self._split(batch_input)
        fake = self.discriminator(self.generator(X))
        real = self.discriminator(Y)
        g_loss = DCGAN._ce(fake, 1.0)
        d_loss = DCGAN._ce(fake, 0.0) + DCGAN._ce(real, 1.0)
        return d_loss, g_loss

    def update(self, batch_input, which, apply=True):
        names = self.d_vars if which == "d" else self.g_vars
        self.set_requires_grad(set(names))
        d_loss, g_loss = self.losses(batch_input)
        (d_loss if which == "d" else g_loss).backward()
        grads = {k: self.vars[k].grad.detach().clone() for k in names}
        self.set_requires_grad(set())
        if apply:
            (self.d_optim if which == "d" else self.g_optim).apply(grads)
        return dict(d_loss=d_loss.item(), g_loss=g_loss.item(), grads=grads)

    def train_step(self, batch_input):
        """recurrent_DCGAN.py:353-375: d_optim, g_optim, g_optim."""
        self.update(batch_input, "d")
        self.update(batch_input, "g")
        return self.update(batch_input, "g")
