"""DCGAN -- drop-in for /root/reference/models/recurrent_z/model.py.

Same constructor signature, method names (`build_model`, `train`, `generator`,
`discriminator`, `sampler`, `save`, `load`, `load_mnist`) and attribute names
(`d_vars`, `g_vars`, `d_loss`, `g_loss`, `G`, `D_logits`, `D_activations(_inf)(_)`,
`h0_w` ... `h4_b`).  TensorFlow's deferred graph + `sess.run` is replaced by eager calls
of the sm_100a kernels (ops.py) and, for training, one CUDA graph of the whole
reference step (1 D update + 2 G updates, model.py:226-239):

    sess.run([d_optim, ...], {images, z})  ->  self.d_update(images, z)
    sess.run([g_optim, ...], {z})          ->  self.g_update(z)
    the three of them per batch            ->  self.train_step(images, z)

`sess` is accepted and ignored.  Symbolic attributes built by `build_model` (self.G,
self.D_logits, ...) are `meta` tensors carrying the static shapes, as the TF graph
nodes did.
"""
from __future__ import annotations

import os
import time
from collections import OrderedDict
from glob import glob

import numpy as np
import torch

from . import ops
from .ops import (batch_norm, conv2d, conv_cond_concat, deconv2d, linear, lrelu, add_noise, get_std,
                  sigmoid_cross_entropy_loss)
from .utils import get_image, save_images


class DCGAN(object):
    def __init__(self, sess=None, image_size=108, is_crop=True,
                 batch_size=64, sample_size=64, output_size=64,
                 y_dim=None, z_dim=100, gf_dim=64, df_dim=64,
                 gfc_dim=1024, dfc_dim=1024, c_dim=3, dataset_name='default',
                 checkpoint_dir=None, sample_dir=None, data_dir='./data',
                 log_dir='./logs', image_glob='*.jpg', shuffle=False,
                 z=None, sample_z=None, noise_std=0.0, *, store=None, standalone=True,
                 learning_rate=0.0002, beta1=0.5, dp=None):
        """Arguments as model.py:13-19.  Keyword-only extensions: `store` (share a
        VariableStore, used by VID_DCGAN), `standalone` (finalize the store and create the
        optimisers here), `dp` (a gifgan.dp.DataParallel gradient reducer)."""
        self.sess = sess
        self.is_crop = is_crop
        self.is_grayscale = (c_dim == 1)
        self.batch_size = batch_size
        self.image_size = image_size
        self.sample_size = sample_size
        self.output_size = output_size
        self.data_dir = data_dir
        self.log_dir = log_dir
        self.image_glob = image_glob
        self.shuffle = shuffle
        self.noise_std = noise_std

        self.y_dim = y_dim
        self.z_dim = z_dim
        self.gf_dim = gf_dim
        self.df_dim = df_dim
        self.gfc_dim = gfc_dim
        self.dfc_dim = dfc_dim
        self.c_dim = c_dim

        self.store = store if store is not None else ops.default_store()
        self.scope_prefix = self.store.scope_name()
        self.dp = dp

        # batch normalization : deals with poor initialization helps gradient flow (model.py:58-70)
        self.d_bn1 = batch_norm(name='d_bn1')
        self.d_bn2 = batch_norm(name='d_bn2')
        if not self.y_dim:
            self.d_bn3 = batch_norm(name='d_bn3')
        self.g_bn0 = batch_norm(name='g_bn0')
        self.g_bn1 = batch_norm(name='g_bn1')
        self.g_bn2 = batch_norm(name='g_bn2')
        if not self.y_dim:
            self.g_bn3 = batch_norm(name='g_bn3')

        self.dataset_name = dataset_name
        self.checkpoint_dir = checkpoint_dir
        self._graph = None
        self.want_sigmoid = True    # D = sigmoid(logits) is only used for summaries; the train step skips it
        self.build_model(z, sample_z)
        if standalone:
            self.store.finalize(OrderedDict(d=self.d_vars, g=self.g_vars))
            self.d_optim = ops.AdamOptimizer(self.store, "d", learning_rate, beta1)
            self.g_optim = ops.AdamOptimizer(self.store, "g", learning_rate, beta1)
            self.d_optim.var_list, self.g_optim.var_list = self.d_vars, self.g_vars

    # ------------------------------------------------------------------------------
    def build_model(self, z, sample_z):
        """model.py:76-141: trace the graph on meta tensors -> creates every variable with its
        TF name, records static shapes, splits d_vars / g_vars by name."""
        B, s, c = self.batch_size, self.output_size, self.c_dim
        meta = lambda *shape: torch.empty(shape, dtype=torch.float32, device="meta")
        n_before = len(self.store.vars)
        if self.y_dim:
            self.y = meta(B, self.y_dim)
        self.images = meta(B, s, s, c)
        self.noisy_images = add_noise(self.images, self.noise_std)
        self.images_std = get_std(self.images)
        self.sample_images = meta(self.sample_size, s, s, c)
        self.z = z if z is not None else meta(B, self.z_dim)
        self.sample_z = sample_z if sample_z is not None else self.z

        if self.y_dim:
            self.G = self.generator(self.z, self.y)
            self.D, self.D_logits = self.discriminator(self.noisy_images, self.y, reuse=False)
            self._sampler_shape = self.sampler(self.z, self.y)
            self.D_, self.D_logits_ = self.discriminator(self.G, self.y, reuse=True)
        else:
            self.G = self.generator(self.z)
            self.noisy_G = add_noise(self.G, self.noise_std)
            self.G_std = get_std(self.G)
            self.D, self.D_logits, self.D_activations = self.discriminator(self.noisy_images)
            self.D_inf, self.D_logits_inf, self.D_activations_inf = self.discriminator(self.noisy_images, reuse=True, train=False)
            self._sampler_shape = self.sampler(self.sample_z)
            self.D_, self.D_logits_, self.D_activations_ = self.discriminator(self.noisy_G, reuse=True)
            self.D_inf_, self.D_logits_inf_, self.D_activations_inf_ = self.discriminator(self._sampler_shape, reuse=True, train=False)

        self.d_loss_real = sigmoid_cross_entropy_loss(self.D_logits, target=1.0)[0]
        self.d_loss_fake = sigmoid_cross_entropy_loss(self.D_logits_, target=0.0)[0]
        self.g_loss = sigmoid_cross_entropy_loss(self.D_logits_, target=1.0)[0]
        self.d_loss = self.d_loss_real  # symbolic placeholder for d_loss_real + d_loss_fake (model.py:131)

        # model.py:136-139: t_vars split by substring of the variable name (within this model's scope)
        mine = [v for v in list(self.store.vars.values())[n_before:]]
        local = lambda v: v.name[len(self.scope_prefix):]
        self.d_vars = [v for v in mine if v.trainable and 'd_' in local(v)]
        self.g_vars = [v for v in mine if v.trainable and 'g_' in local(v)]
        self.all_vars = mine

    # ------------------------------------------------------------------------------
    def discriminator(self, image, y=None, reuse=False, train=True, groups=1, stop_at_h2=False, ce_segments=None):
        with self.store.absolute_scope(self.scope_prefix):
            return self._discriminator(image, y, reuse, train, groups, stop_at_h2, ce_segments)

    def _discriminator(self, image, y=None, reuse=False, train=True, groups=1, stop_at_h2=False, ce_segments=None):
        """model.py:268-296.  `groups=2` runs D(real) and D(fake) as one batch whose halves are
        batch-normalised separately (identical numbers to two calls, half the launches).  `stop_at_h2`
        evaluates only what D_activations needs (what TF's pruning does for VID_DCGAN's fetches)."""
        if not self.y_dim:
            B = image.shape[0]
            h0 = conv2d(image, self.df_dim, name='d_h0_conv', act='lrelu')
            if stop_at_h2:
                h1 = conv2d(h0, self.df_dim * 2, name='d_h1_conv', bn=self.d_bn1, train=train, act='lrelu', groups=groups)
                h2 = conv2d(h1, self.df_dim * 4, name='d_h2_conv', bn=self.d_bn2, train=train, act='lrelu', groups=groups)
                return None, None, h2
            # lrelu(d_bnN(conv2d(...), train=train)) -- conv + batch norm + LeakyReLU as one fused node each
            h1 = conv2d(h0, self.df_dim * 2, name='d_h1_conv', bn=self.d_bn1, train=train, act='lrelu', groups=groups)
            h2 = conv2d(h1, self.df_dim * 4, name='d_h2_conv', bn=self.d_bn2, train=train, act='lrelu', groups=groups)
            h3 = conv2d(h2, self.df_dim * 8, name='d_h3_conv', bn=self.d_bn3, train=train, act='lrelu', groups=groups)
            # `ce_segments`: the train step's cross-entropy means (model.py:121-131) ride in the same launch as the logits, and
            # the launch that back-propagates them carries d_bn3's backward reductions (ops.reshape keeps the hand-over)
            h4 = linear(ops.reshape(h3, (B, -1)), 1, 'd_h3_lin', ce_segments=ce_segments)
            return (ops.sigmoid(h4) if self.want_sigmoid else None), h4, h2
        else:
            B = image.shape[0]
            yb = y.reshape(B, 1, 1, self.y_dim)
            x = conv_cond_concat(image, yb)
            h0 = conv2d(x, self.c_dim + self.y_dim, name='d_h0_conv', act='lrelu')
            h0 = conv_cond_concat(h0, yb)
            # model.py:287: no train= argument -> always batch statistics
            h1 = conv2d(h0, self.df_dim + self.y_dim, name='d_h1_conv', bn=self.d_bn1, act='lrelu', groups=groups)
            h1 = torch.cat([h1.reshape(B, -1), y.to(h1.dtype)], 1)
            h2 = linear(h1, self.dfc_dim, 'd_h2_lin', bn=self.d_bn2, act='lrelu', groups=groups)
            h2 = torch.cat([h2, y.to(h2.dtype)], 1)
            h3 = linear(h2, 1, 'd_h3_lin')
            return (ops.sigmoid(h3) if self.want_sigmoid else None), h3

    def generator(self, z, y=None, train=True, out=None):
        with self.store.absolute_scope(self.scope_prefix):
            return self._generator(z, y, train, out)

    def _generator(self, z, y=None, train=True, out=None):
        """model.py:298-344 (and, with train=False, the sampler graph of model.py:346-389).
        `out=` lets the last layer write the images into a caller buffer (e.g. the fake half of D's batch)."""
        B = z.shape[0]
        if not self.y_dim:
            s = self.output_size
            s2, s4, s8, s16 = int(s / 2), int(s / 4), int(s / 8), int(s / 16)
            # project `z` and reshape
            # relu(g_bnN(...)): linear/deconv + batch norm + ReLU as one fused node each (model.py:304-319)
            h0, self.h0_w, self.h0_b = linear(z, self.gf_dim * 8 * s16 * s16, 'g_h0_lin', with_w=True, bn=self.g_bn0,
                                              bn_channels=self.gf_dim * 8, train=train, act='relu')
            h0 = ops.reshape(h0, (-1, s16, s16, self.gf_dim * 8))
            h1, self.h1_w, self.h1_b = deconv2d(h0, [B, s8, s8, self.gf_dim * 4], name='g_h1', with_w=True, bn=self.g_bn1,
                                                train=train, act='relu')
            h2, self.h2_w, self.h2_b = deconv2d(h1, [B, s4, s4, self.gf_dim * 2], name='g_h2', with_w=True, bn=self.g_bn2,
                                                train=train, act='relu')
            h3, self.h3_w, self.h3_b = deconv2d(h2, [B, s2, s2, self.gf_dim * 1], name='g_h3', with_w=True, bn=self.g_bn3,
                                                train=train, act='relu')
            h4, self.h4_w, self.h4_b = deconv2d(h3, [B, s, s, self.c_dim], name='g_h4', with_w=True, act='tanh',
                                                out_dtype=torch.float32, out=out)
            return h4
        else:
            s = self.output_size
            s2, s4 = int(s / 2), int(s / 4)
            yb = y.reshape(B, 1, 1, self.y_dim)
            z = torch.cat([z, y], 1)
            # model.py:331 / 380: g_bn0 is never given train=False -> batch statistics even in the sampler
            h0 = linear(z, self.gfc_dim, 'g_h0_lin', bn=self.g_bn0, act='relu')
            h0 = torch.cat([h0, y.to(h0.dtype)], 1)
            h1 = linear(h0, self.gf_dim * 2 * s4 * s4, 'g_h1_lin', bn=self.g_bn1, train=train, act='relu')
            h1 = h1.reshape(B, s4, s4, self.gf_dim * 2)
            h1 = conv_cond_concat(h1, yb)
            h2 = deconv2d(h1, [B, s2, s2, self.gf_dim * 2], name='g_h2', bn=self.g_bn2, train=train, act='relu')
            h2 = conv_cond_concat(h2, yb)
            return deconv2d(h2, [B, s, s, self.c_dim], name='g_h3', act='sigmoid', out_dtype=torch.float32, out=out)

    def sampler(self, z, y=None):
        """model.py:346-389: same weights, inference-mode batch norm."""
        return self.generator(z, y, train=False)

    # ------------------------------------------------------------------------------
    # the three sess.run calls of the train loop
    # ------------------------------------------------------------------------------
    def _dev(self, a, dtype=torch.float32):
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        return a.to(self.store.device, dtype=dtype, non_blocking=True)

    def _both_buffer(self, B):
        """[2B, s, s, c] fp32: real images in the first half (the H2D copy lands here), G's output in the second."""
        buf = getattr(self, "_both", None)
        if buf is None or buf.shape[0] != 2 * B:
            buf = torch.empty((2 * B, self.output_size, self.output_size, self.c_dim), dtype=torch.float32, device=self.store.device)
            self._both = buf
        return buf

    def d_update(self, images, z, y=None, apply=True):
        """sess.run([d_optim, d_sum]) (model.py:227-229): G fwd, D(real), D(fake), backward into d_vars, Adam.
        D(real) and D(fake) run as one 2B batch with per-half batch-norm statistics (groups=2)."""
        B = images.shape[0]
        both = self._both_buffer(B)
        if images.data_ptr() != both.data_ptr():
            both[:B].copy_(images)
        if self.dp is not None:
            self.dp.wait_pending()
        self.d_optim.zero_grad(overlap=self.dp is None, tick=apply)
        if self.dp is not None:
            self.dp.begin_update(self.d_optim)
        with ops.trainable(self.d_vars), ops.overlap_wgrad(), ops.stats_arena():
            with torch.no_grad():
                self.generator(z, y, out=both[B:])
            if self.noise_std:
                both = add_noise(both, self.noise_std)
            yy = torch.cat([y, y], 0) if y is not None else None
            segs = [(0, B, 1.0, 1.0), (B, 2 * B, 0.0, 1.0)]
            logits = self.discriminator(both, yy, reuse=True, groups=2, ce_segments=segs)[1]
            losses = sigmoid_cross_entropy_loss(logits, segs)
            ops.join_side()         # the gradient zero-fill (side stream) precedes every gradient kernel
            torch.autograd.backward(losses, grad_tensors=self._ones(losses))
        if self.dp is not None and apply:
            # remaining bucket + Adam on the communication stream: the next update's generator forward overlaps them
            self.dp.finish_update(self.d_optim, lambda: self.d_optim.apply(grad_scale=self._grad_scale()))
        else:
            if self.dp is not None:
                self.dp.allreduce(self.d_optim)
            if apply:
                self.d_optim.apply(grad_scale=self._grad_scale())
        return losses          # [d_loss, d_loss_real, d_loss_fake]

    def g_update(self, z, y=None, apply=True):
        """sess.run([g_optim, g_sum]) (model.py:232-234): G fwd, D(fake), backward through D into g_vars, Adam."""
        self.g_optim.zero_grad(overlap=self.dp is None, tick=apply)
        if self.dp is not None:
            self.dp.begin_update(self.g_optim)
        with ops.trainable(self.g_vars), ops.overlap_wgrad(), ops.stats_arena():
            G = self.generator(z, y)
            if self.dp is not None:
                self.dp.wait_pending()      # the discriminator's update (exchange + Adam) may still be in flight
            segs = [(0, z.shape[0], 1.0, 1.0)]
            logits = self.discriminator(add_noise(G, self.noise_std), y, reuse=True, ce_segments=segs)[1]
            losses = sigmoid_cross_entropy_loss(logits, segs)
            ops.join_side()
            torch.autograd.backward(losses, grad_tensors=self._ones(losses))
        if self.dp is not None:
            self.dp.allreduce(self.g_optim)
        if apply:
            self.g_optim.apply(grad_scale=self._grad_scale())
        return losses          # [g_loss, g_loss]

    def _ones(self, like):
        c = getattr(self, "_ones_cache", None)
        if c is None:
            c = self._ones_cache = {}
        n = like.numel()
        if n not in c:
            c[n] = torch.ones(n, dtype=torch.float32, device=self.store.device)
        return c[n]

    def _grad_scale(self):
        return 1.0 if self.dp is None else 1.0 / self.dp.world_size

    def eval_losses(self, images, z, y=None):
        """model.py:241-243: d_loss_fake.eval, d_loss_real.eval, g_loss.eval -- three forward-only runs in
        train-mode BN (each advances the EMAs, App. A.4)."""
        with torch.no_grad():
            G = self.generator(z, y)
            # the losses are the training graph's tensors: D sees its inputs through add_noise (model.py:105-107)
            errD_fake = sigmoid_cross_entropy_loss(self.discriminator(add_noise(G, self.noise_std), y, reuse=True)[1], target=0.0)
            errD_real = sigmoid_cross_entropy_loss(self.discriminator(add_noise(images, self.noise_std), y, reuse=True)[1], target=1.0)
            G = self.generator(z, y)
            errG = sigmoid_cross_entropy_loss(self.discriminator(add_noise(G, self.noise_std), y, reuse=True)[1], target=1.0)
        return errD_fake, errD_real, errG

    LOSS_KEYS = ("d_loss", "g_loss_first", "g_loss", "errD_fake", "errD_real", "errG")

    def _step_device(self, images, z, y, evals, loss_vec):
        """The loop body of model.py:226-243 on device tensors; scalar losses are gathered into loss_vec."""
        self.want_sigmoid = False
        d = self.d_update(images, z, y)
        g1 = self.g_update(z, y)
        # Run g_optim twice to make sure that d_loss does not go to zero (model.py:236-239)
        g2 = self.g_update(z, y)
        outs = [d, g1, g2]
        if evals:
            outs += list(self.eval_losses(images, z, y))
        ops.cabi.gather_scalars([o.reshape(-1)[0:1] for o in outs], loss_vec)
        self.want_sigmoid = True

    def train_step(self, batch_images, batch_z, batch_labels=None, evals=False, use_graph=True, sync=True):
        """One iteration of the loop body of model.py:226-243.  The batch may be host memory (numpy / pinned
        torch; copied to the device inside this call, like a feed_dict) or device tensors.  With use_graph the
        whole step is captured once in a CUDA graph and replayed.  Returns a dict of Python floats (sync=True,
        includes the device->host read of the losses) or the device loss vector (sync=False)."""
        st = self._static_buffers(np.shape(batch_images)[0], batch_labels is not None)
        B = st["B"]
        if self.dp is not None:
            self.dp.prepare(self.store)        # peer-memory exchange: collective handshake, first call only (never inside a capture)
        st["both"][:B].copy_(torch.as_tensor(batch_images), non_blocking=True)
        st["z"].copy_(torch.as_tensor(batch_z), non_blocking=True)
        if batch_labels is not None:
            st["y"].copy_(torch.as_tensor(batch_labels), non_blocking=True)
        n = 6 if evals else 3
        if use_graph:
            key = (B, evals, batch_labels is not None)
            g = self._graph if (self._graph is not None and self._graph["key"] == key) else self._capture(st, evals, key)
            ops.refresh_packs(self.store)      # no-op unless weights were changed from outside (checkpoint load)
            g["graph"].replay()
            self.d_optim.t += 1
            self.g_optim.t += 2
        else:
            self._step_device(st["both"][:B], st["z"], st.get("y"), evals, st["loss_dev"])
        if not sync:
            return st["loss_dev"][:n]
        st["loss_host"].copy_(st["loss_dev"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return {k: float(st["loss_host"][i]) for i, k in enumerate(self.LOSS_KEYS[:n])}

    def _static_buffers(self, B, has_y):
        st = getattr(self, "_static", None)
        if st is None or st["B"] != B or (("y" in st) != has_y):
            dev = self.store.device
            st = dict(B=B, both=self._both_buffer(B), z=torch.empty((B, self.z_dim), dtype=torch.float32, device=dev),
                      loss_dev=torch.zeros(6, dtype=torch.float32, device=dev),
                      loss_host=torch.zeros(6, dtype=torch.float32).pin_memory())
            if has_y:
                st["y"] = torch.empty((B, self.y_dim), dtype=torch.float32, device=dev)
            self._static = st
            self._graph = None
        return st

    def _capture(self, st, evals, key):
        """Capture d_update + 2 x g_update (+ the three evals) into one CUDA graph."""
        B = st["B"]
        args = (st["both"][:B], st["z"], st.get("y"), evals, st["loss_dev"])
        # snapshot: the warm-up run (allocator, lazy packs) must not change the training trajectory
        snap = {k: t.clone() for k, t in self.store.flat.items()}
        states = (self.d_optim.state.clone(), self.g_optim.state.clone(), self.d_optim.t, self.g_optim.t)

        def restore():
            for k, t in snap.items():
                self.store.flat[k].copy_(t)
            self.d_optim.state.copy_(states[0]); self.g_optim.state.copy_(states[1])
            self.d_optim.t, self.g_optim.t = states[2], states[3]
            for v in self.store.vars.values():
                v.invalidate_packed()
            ops.refresh_packs(self.store)      # the captured step is entered with current bf16 copies (Adam keeps them current inside it)

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._step_device(*args)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        restore()
        graph = torch.cuda.CUDAGraph()
        n0 = ops.cabi.launch_count()
        with torch.cuda.graph(graph):
            self._step_device(*args)
        launches = ops.cabi.launch_count() - n0
        restore()
        torch.cuda.synchronize()
        self._graph = dict(graph=graph, key=key, launches=launches)
        return self._graph

    # ------------------------------------------------------------------------------
    def train(self, config):
        """Train DCGAN (model.py:143-266).  `config` carries the flags of main.py:10-29."""
        if config.dataset == 'mnist':
            data_X, data_y = self.load_mnist()
        elif config.dataset == 'synthetic':
            # the flag's default train_size is "everything" (1 << 62): a synthetic epoch is 4096 samples unless a smaller size is asked
            # for, drawn in float32 (a float64 draw of 65536 x 64 x 64 x 3 would take 6.4 GB of host memory)
            n = int(min(getattr(config, "train_size", 4096), 4096))
            gen = np.random.Generator(np.random.PCG64(102))
            data = None
            data_X = gen.random((n, self.output_size, self.output_size, self.c_dim), dtype=np.float32) * 2.0 - 1.0
        else:
            data = glob(os.path.join(self.data_dir, config.dataset, self.image_glob))
            if self.shuffle:
                np.random.shuffle(data)
        self.d_optim.lr = self.g_optim.lr = config.learning_rate
        self.d_optim.b1 = self.g_optim.b1 = config.beta1

        sample_z = np.random.uniform(-1, 1, size=(self.sample_size, self.z_dim)).astype(np.float32)
        counter = 1
        start_time = time.time()
        if self.load(self.checkpoint_dir):
            print(" [*] Load SUCCESS")
        else:
            print(" [!] Load failed...")

        last = None
        for epoch in range(config.epoch):
            if config.dataset in ('mnist', 'synthetic'):
                batch_idxs = min(len(data_X), config.train_size) // config.batch_size
            else:
                batch_idxs = min(len(data), config.train_size) // config.batch_size
            loader = None
            if config.dataset not in ('mnist', 'synthetic'):
                loader = self.file_batches(data[:int(min(len(data), config.train_size))], config.batch_size)
            for idx in range(0, int(batch_idxs)):
                batch_labels = None
                if config.dataset in ('mnist', 'synthetic'):
                    batch_images = data_X[idx * config.batch_size:(idx + 1) * config.batch_size]
                    if config.dataset == 'mnist':
                        batch_labels = data_y[idx * config.batch_size:(idx + 1) * config.batch_size]
                else:
                    batch_images = next(loader)      # decoded ahead of the step by input_pipeline.Prefetcher
                batch_z = np.random.uniform(-1, 1, [config.batch_size, self.z_dim]).astype(np.float32)

                last = self.train_step(batch_images, batch_z, batch_labels, evals=True)
                counter += 1
                print("Epoch: [%2d] [%4d/%4d] time: %4.4f, d_loss: %.8f, g_loss: %.8f"
                      % (epoch, idx, batch_idxs, time.time() - start_time, last["errD_fake"] + last["errD_real"], last["errG"]))

                if np.mod(counter, 100) == 1 and getattr(config, "sample_dir", None):
                    with torch.no_grad():
                        samples = self.sampler(self._dev(sample_z)).float().cpu().numpy()
                    save_images(samples, [8, 8], '{}/train_{:02d}_{:04d}.png'.format(config.sample_dir, epoch, idx))
                if np.mod(counter, 500) == 2:
                    self.save(config.checkpoint_dir, counter)
        return last

    def file_batches(self, files, batch_size, depth=2, workers=8):
        """The image batches of model.py:212-219 (get_image per file, consecutive slices of the file list) as an iterator
        that decodes `depth` batches ahead on worker threads into page-locked buffers (input_pipeline.Prefetcher)."""
        from .input_pipeline import Prefetcher, chunks
        s = self.output_size
        return Prefetcher(chunks(files, batch_size),
                          lambda f: get_image(f, self.image_size, is_crop=self.is_crop, resize_w=s, is_grayscale=self.is_grayscale),
                          (s, s, self.c_dim), depth=depth, workers=workers)

    def load_mnist(self):
        """model.py:391-426."""
        data_dir = os.path.join(self.data_dir, self.dataset_name)

        def rd(name, off, shape):
            with open(os.path.join(data_dir, name), "rb") as fd:
                return np.frombuffer(fd.read(), dtype=np.uint8)[off:].reshape(shape).astype(np.float64)
        trX, trY = rd('train-images-idx3-ubyte', 16, (60000, 28, 28, 1)), rd('train-labels-idx1-ubyte', 8, (60000,))
        teX, teY = rd('t10k-images-idx3-ubyte', 16, (10000, 28, 28, 1)), rd('t10k-labels-idx1-ubyte', 8, (10000,))
        X = np.concatenate((trX, teX), axis=0)
        y = np.concatenate((trY, teY), axis=0)
        seed = 547
        np.random.seed(seed); np.random.shuffle(X)
        np.random.seed(seed); np.random.shuffle(y)
        y_vec = np.zeros((len(y), self.y_dim), dtype=np.float32)
        y_vec[np.arange(len(y)), y.astype(np.int64)] = 1.0
        return (X / 255.).astype(np.float32), y_vec

    # ------------------------------------------------------------------------------
    def _ckpt_dir(self, checkpoint_dir):
        model_dir = "%s_%s_%s" % (self.dataset_name, self.batch_size, self.output_size)
        return os.path.join(checkpoint_dir, model_dir)

    def save(self, checkpoint_dir, step):
        """model.py:428-439: <checkpoint_dir>/<dataset>_<batch>_<output>/DCGAN.model-<step>; the payload is a
        torch-serialised dict keyed by the TF variable names (+ Adam slots and step counters)."""
        model_name = "DCGAN.model"
        checkpoint_dir = self._ckpt_dir(checkpoint_dir)
        os.makedirs(checkpoint_dir, exist_ok=True)
        path = os.path.join(checkpoint_dir, "%s-%d" % (model_name, step))
        torch.save(self.checkpoint_payload(), path)
        with open(os.path.join(checkpoint_dir, "checkpoint"), "w") as f:
            f.write('model_checkpoint_path: "%s-%d"\n' % (model_name, step))
        return path

    def checkpoint_payload(self):
        f = self.store.flat
        return dict(variables=self.store.state_dict(), adam_m=f["m"].cpu(), adam_v=f["v"].cpu(),
                    d_t=self.d_optim.t, g_t=self.g_optim.t, ranges=dict(self.store.ranges))

    def load_payload(self, payload):
        self.store.load_state_dict(payload["variables"])
        f = self.store.flat
        if payload.get("adam_m") is not None and payload["adam_m"].numel() == f["m"].numel():
            f["m"].copy_(payload["adam_m"]); f["v"].copy_(payload["adam_v"])
            self.d_optim.t, self.g_optim.t = int(payload["d_t"]), int(payload["g_t"])
            self.d_optim.state[0] = self.d_optim.t
            self.g_optim.state[0] = self.g_optim.t

    def load(self, checkpoint_dir):
        """model.py:441-452."""
        print(" [*] Reading checkpoints...")
        if not checkpoint_dir:
            return False
        checkpoint_dir = self._ckpt_dir(checkpoint_dir)
        index = os.path.join(checkpoint_dir, "checkpoint")
        if not os.path.exists(index):
            return False
        with open(index) as f:
            name = os.path.basename(f.readline().split('"')[1])
        from . import checkpoint_io
        if checkpoint_io.tf_format(os.path.join(checkpoint_dir, name)):        # written by TensorFlow (or checkpoint_io): V2 bundle / V1 file
            checkpoint_io.load_tf_checkpoint(os.path.join(checkpoint_dir, name), self.store, (self.d_optim, self.g_optim))
            return True
        self.load_payload(torch.load(os.path.join(checkpoint_dir, name), map_location="cpu", weights_only=True))
        return True
