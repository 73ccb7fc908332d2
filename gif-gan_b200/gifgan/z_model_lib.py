"""VID_DCGAN -- drop-in for /root/reference/models/recurrent_z/z_model_lib.py.

Same constructor signature, method names (`build_model`, `generator`, `discriminator`, `train`, `dump_sample`,
`load_image_gan`, `load_checkpoint`, `load_videos`) and attributes (`G`, `G_sampler`, `G_out`, `img_dcgan`,
`d_fake_out`, `d_real_out`, `d_vid_vars`, `g_vid_vars`, `d_img_vars`, `g_img_vars`).  The latent generator is
the 4-layer MLP over [z || frame index] (z_model_lib.py:353-382); the video discriminator is three stride-2
3x3x3 conv3d layers over the image discriminator's `h2` activations (z_model_lib.py:384-418); the nested image
DCGAN runs in inference-mode batch norm and, with the default flags, is frozen (z_model.py:46-47).

    sess.run([d_optim, d_loss, stds...], {images, z, is_training: True})  ->  self.d_update(images, z)
    sess.run([g_optim, g_loss, g_loss_first_frame], {z, is_training: True})  ->  self.g_update(z)
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np
import torch

from . import ops
from .model import DCGAN
from .ops import batch_norm, conv3d, linear, add_noise, get_std, sigmoid_cross_entropy_loss, variable_scope
from .utils import inverse_transform, open_video_writer, transform


class Layers(object):
    pass


class _FirstFrameMSE(torch.autograd.Function):
    """first_frame_loss_scalar * mean((G_out[::T] - z[:, :z_out])^2)  (z_model_lib.py:109-111)."""

    @staticmethod
    def forward(ctx, g_out, z, T, z_out, scalar):
        rows = z.shape[0]
        loss = torch.empty(1, dtype=torch.float32, device=g_out.device)
        da = torch.empty((rows, z_out), dtype=torch.float32, device=g_out.device)
        ops.check(ops.cabi.lib().gg_mse(ops.ptr(g_out), T * g_out.shape[1], ops.ptr(z), z.shape[1], rows, z_out, float(scalar),
                                        ops.ptr(loss), 0, ops.ptr(da), ops.stream()), "gg_mse")
        ctx.T, ctx.shape, ctx.da = T, g_out.shape, da
        return loss

    @staticmethod
    def backward(ctx, g):
        dg = torch.zeros(ctx.shape, dtype=torch.float32, device=ctx.da.device)
        dg[:: ctx.T] = ctx.da
        return dg, None, None, None, None


class VID_DCGAN(object):
    def __init__(self, sess, batch_size, z_input_size, z_output_size, vid_length,
                 input_image_size, output_image_size, c_dim,
                 sample_cols=8, image_noise_std=0.0, activation_noise_std=0.0,
                 first_frame_loss_scalar=0.0, z=None, *, store=None, learning_rate=0.0002, beta1=0.5,
                 train_img_gen=False, train_img_disc=False, dp=None):
        # Member vars (z_model_lib.py:20-33)
        self.batch_size = batch_size
        self.z_input_size = z_input_size
        self.z_output_size = z_output_size
        self.vid_length = vid_length
        self.input_image_size = input_image_size
        self.output_image_size = output_image_size
        self.c_dim = c_dim
        self.sample_cols = sample_cols
        assert batch_size % sample_cols == 0
        self.sample_rows = batch_size // sample_cols
        self.image_noise_std = image_noise_std
        self.activation_noise_std = activation_noise_std
        self.first_frame_loss_scalar = first_frame_loss_scalar
        self.store = store if store is not None else ops.default_store()
        self.scope_prefix = self.store.scope_name()      # e.g. 'video_gan/' (z_model.py:63)
        self.dp = dp

        # Batch norm layers (z_model_lib.py:36-44; bn3 / bn0,1,4 are constructed but never used there either)
        self.g_bn0 = batch_norm(name='gvideo_bn0')
        self.g_bn1 = batch_norm(name='gvideo_bn1')
        self.g_bn2 = batch_norm(name='gvideo_bn2')
        self.g_bn3 = batch_norm(name='gvideo_bn3')
        self.d_bn0 = batch_norm(name='dvideo_bn0')
        self.d_bn1 = batch_norm(name='dvideo_bn1')
        self.d_bn2 = batch_norm(name='dvideo_bn2')
        self.d_bn3 = batch_norm(name='dvideo_bn3')
        self.d_bn4 = batch_norm(name='dvideo_bn4')

        self._frame_numbers = None
        self._graphs = {}
        self.build_model(sess, z)

        # optimiser groups: contiguous ranges so that "video vars (+ image vars)" is one Adam launch / one all-reduce
        self.store.finalize(OrderedDict(dvideo=self.d_vid_vars, d_img=self.d_img_vars, gvideo=self.g_vid_vars, g_img=self.g_img_vars))
        self.learning_rate, self.beta1 = learning_rate, beta1
        self.set_trainable(train_img_gen, train_img_disc)

    def set_trainable(self, train_img_gen=False, train_img_disc=False):
        """z_model_lib.py:166-179: d_vars = dvideo (+ d_ if --train_img_disc), g_vars = gvideo (+ g_ if --train_img_gen)."""
        self.train_img_gen, self.train_img_disc = train_img_gen, train_img_disc
        self.d_var_list = self.d_vid_vars + (self.d_img_vars if train_img_disc else [])
        self.g_var_list = self.g_vid_vars + (self.g_img_vars if train_img_gen else [])
        self.d_optim = ops.AdamOptimizer(self.store, ("dvideo", "d_img") if train_img_disc else "dvideo", self.learning_rate, self.beta1)
        self.g_optim = ops.AdamOptimizer(self.store, ("gvideo", "g_img") if train_img_gen else "gvideo", self.learning_rate, self.beta1)
        self.d_optim.var_list, self.g_optim.var_list = self.d_var_list, self.g_var_list
        self._graphs = {}

    # ------------------------------------------------------------------------------
    def build_model(self, sess, z):
        """z_model_lib.py:49-115, traced on meta tensors (creates the variables under the reference's scopes)."""
        meta = lambda *shape: torch.empty(shape, dtype=torch.float32, device="meta")
        self.z = z if z is not None else meta(self.batch_size, self.z_input_size)
        with variable_scope('video_generator'):
            self.G, self.G_layers = self.generator(self.z, reuse=False, train=True)
            self.G_sampler, self.G_sampler_layers = self.generator(self.z, reuse=True, train=False)
            self.G_out = self.G
            self.first_frames = self.G_out[::self.vid_length, :]
        # Build the inner image gan (z_model_lib.py:68-77)
        with variable_scope('image_gan'):
            self.img_dcgan = DCGAN(sess, image_size=self.input_image_size,
                                   batch_size=self.batch_size * self.vid_length,
                                   output_size=self.output_image_size,
                                   z_dim=self.z_output_size, c_dim=self.c_dim,
                                   dataset_name='', is_crop=False,
                                   checkpoint_dir='', sample_dir='',
                                   data_dir='', log_dir='', image_glob='', shuffle=False,
                                   z=self.G_out, noise_std=self.image_noise_std, store=self.store, standalone=False)
            self.image_gan_scope_name = self.store.scope_name()
        # Build discriminator (z_model_lib.py:80-93)
        with variable_scope('video_discriminator'):
            self.noisy_D_activations_inf = add_noise(self.img_dcgan.D_activations_inf, self.activation_noise_std)
            self.D_activations_inf_std = get_std(self.img_dcgan.D_activations_inf)
            self.d_real_out, self.d_real_out_logits, self.D_real_layers = self.discriminator(self.noisy_D_activations_inf, reuse=False)
            self.noisy_D_activations_inf_ = add_noise(self.img_dcgan.D_activations_inf_, self.activation_noise_std)
            self.D_activations_inf_std_ = get_std(self.img_dcgan.D_activations_inf_)
            self.d_fake_out, self.d_fake_out_logits, self.D_fake_layers = self.discriminator(self.noisy_D_activations_inf_, reuse=True)
        # trainable variables (z_model_lib.py:95-99)
        t_vars = [v for v in self.store.vars.values() if v.trainable]
        self.d_vid_vars = [v for v in t_vars if 'dvideo_' in v.name]
        self.g_vid_vars = [v for v in t_vars if 'gvideo_' in v.name]
        self.d_img_vars = self.img_dcgan.d_vars
        self.g_img_vars = self.img_dcgan.g_vars

    # ------------------------------------------------------------------------------
    def _z_with_numbers(self, z):
        """z_model_lib.py:358-370: tile z over T and append the frame index linspace(-1, 1, T)."""
        Bv, T = z.shape[0], self.vid_length
        if z.device.type == "meta":
            return torch.empty((Bv * T, z.shape[1] + 1), dtype=torch.float32, device="meta")
        if self._frame_numbers is None or self._frame_numbers.device != z.device or self._frame_numbers.shape[0] != Bv:
            fn = torch.tensor(np.linspace(-1.0, 1.0, T), dtype=torch.float32, device=z.device)
            self._frame_numbers = fn[None, :, None].expand(Bv, T, 1).contiguous()
        return torch.cat([z[:, None, :].expand(Bv, T, z.shape[1]), self._frame_numbers], 2).reshape(Bv * T, -1)

    def generator(self, z, reuse=False, train=True):
        with self.store.absolute_scope(self.scope_prefix + 'video_generator/'):
            return self._generator(z, reuse, train)

    def _generator(self, z, reuse=False, train=True):
        layers = Layers()
        z_reshaped = self._z_with_numbers(z)
        layers.gr0 = linear(z_reshaped, 512, 'gvideo_0', bn=self.g_bn0, train=train, act='relu')
        layers.gr1 = linear(layers.gr0, 512, 'gvideo_1', bn=self.g_bn1, train=train, act='relu')
        layers.gr2 = linear(layers.gr1, 512, 'gvideo_2', bn=self.g_bn2, train=train, act='relu')
        layers.gr3 = linear(layers.gr2, self.z_output_size, 'gvideo_3', act='tanh', out_dtype=torch.float32)
        return layers.gr3, layers

    def discriminator(self, vid, reuse=False, groups=1, ce_segments=None):
        with self.store.absolute_scope(self.scope_prefix + 'video_discriminator/'):
            return self._discriminator(vid, reuse, groups, ce_segments)

    def _discriminator(self, vid, reuse=False, groups=1, ce_segments=None):
        """z_model_lib.py:384-418 (batch norm always in train mode).  `groups=2`: real and fake clips as one batch."""
        layers = Layers()
        nclips = vid.shape[0] // self.vid_length
        vid = vid.reshape(nclips, self.vid_length, vid.shape[1], vid.shape[2], -1)
        layers.dr0 = vid
        layers.dr1 = conv3d(layers.dr0, 256, name='dvideo_h1', act='lrelu')
        layers.dr2 = conv3d(layers.dr1, 256, name='dvideo_h2', bn=self.d_bn2, act='lrelu', groups=groups)
        layers.dr3 = conv3d(layers.dr2, 256, name='dvideo_h3', bn=self.d_bn3, act='lrelu', groups=groups)
        layers.d4 = linear(ops.reshape(layers.dr3, (nclips, -1)), 1, 'dvideo_h4', ce_segments=ce_segments)   # (see DCGAN._discriminator)
        return None, layers.d4, layers

    # ------------------------------------------------------------------------------
    def _both(self):
        n = self.batch_size * self.vid_length
        s, c = self.output_image_size, self.c_dim
        buf = getattr(self, "_both_buf", None)
        if buf is None:
            buf = self._both_buf = torch.empty((2 * n, s, s, c), dtype=torch.float32, device=self.store.device)
        return buf

    def _ones(self, like):
        c = getattr(self, "_ones_cache", None)
        if c is None:
            c = self._ones_cache = {}
        n = like.numel()
        if n not in c:
            c[n] = torch.ones(n, dtype=torch.float32, device=self.store.device)
        return c[n]

    def d_update(self, images, z, apply=True, diagnostics=True):
        """One discriminator update (z_model_lib.py:219-229): real clips and sampler(G(z)) clips through the
        inference-mode image discriminator up to h2, then the video discriminator; backward into d_var_list."""
        img = self.img_dcgan
        n, Bv = self.batch_size * self.vid_length, self.batch_size
        both = self._both()
        if images.data_ptr() != both.data_ptr():
            both[:n].copy_(images)
        if self.dp is not None:
            self.dp.wait_pending()
        self.d_optim.zero_grad(overlap=self.dp is None, tick=apply)
        if self.dp is not None:
            self.dp.begin_update(self.d_optim)
        with ops.trainable(self.d_var_list), ops.overlap_wgrad(), ops.stats_arena():
            with torch.no_grad():
                G_out, _ = self.generator(z, train=True)
                img.generator(G_out, train=False, out=both[n:])                  # img_dcgan.sampler(G_out)
            act = img.discriminator(add_noise(both, self.image_noise_std), reuse=True, train=False, stop_at_h2=True)[2]   # D_activations_inf(_)
            act = add_noise(act, self.activation_noise_std)
            segs = [(0, Bv, 1.0, 1.0), (Bv, 2 * Bv, 0.0, 1.0)]
            logits = self.discriminator(act, reuse=True, groups=2, ce_segments=segs)[1]
            losses = sigmoid_cross_entropy_loss(logits, segs)
            ops.join_side()         # the gradient zero-fill / Adam tick (side stream) precede every gradient kernel
            torch.autograd.backward(losses, grad_tensors=self._ones(losses))
        if self.dp is not None and apply:
            # remaining bucket + Adam on the communication stream: the next update's generator / image-GAN forward overlaps them
            self.dp.finish_update(self.d_optim, lambda: self.d_optim.apply(grad_scale=1.0 / self.dp.world_size))
        else:
            if self.dp is not None:
                self.dp.allreduce(self.d_optim)
            if apply:
                self.d_optim.apply(grad_scale=1.0 if self.dp is None else 1.0 / self.dp.world_size)
        out = dict(losses=losses)
        if diagnostics:   # the std fetches of z_model_lib.py:220-222
            out.update(images_std=get_std(both[:n]), sampler_std=get_std(both[n:]),
                       real_D_std=get_std(act[:n].detach()), fake_D_std=get_std(act[n:].detach()))
        return out

    def g_update(self, z, apply=True):
        """One generator update (z_model_lib.py:233-239)."""
        img = self.img_dcgan
        self.g_optim.zero_grad(overlap=self.dp is None, tick=apply)
        if self.dp is not None:
            self.dp.begin_update(self.g_optim)
        with ops.trainable(self.g_var_list), ops.overlap_wgrad(), ops.stats_arena():
            G_out, _ = self.generator(z, train=True)
            frames = img.generator(G_out, train=False)
            if self.dp is not None and self.train_img_disc:
                self.dp.wait_pending()      # --train_img_disc: the pending D update also rewrites the image discriminator
            act = img.discriminator(add_noise(frames, self.image_noise_std), reuse=True, train=False, stop_at_h2=True)[2]
            if self.dp is not None:
                self.dp.wait_pending()      # the video discriminator's update (exchange + Adam) may still be in flight
            segs = [(0, act.shape[0] // self.vid_length, 1.0, 1.0)]
            logits = self.discriminator(add_noise(act, self.activation_noise_std), reuse=True, ce_segments=segs)[1]
            losses = sigmoid_cross_entropy_loss(logits, segs)
            roots, grads = [losses], [self._ones(losses)]
            first = None
            if self.first_frame_loss_scalar:
                first = _FirstFrameMSE.apply(G_out, z, self.vid_length, self.z_output_size, self.first_frame_loss_scalar)
                roots.append(first); grads.append(self._ones(first))
            ops.join_side()
            torch.autograd.backward(roots, grad_tensors=grads)
        if self.dp is not None:
            self.dp.allreduce(self.g_optim)
        if apply:
            self.g_optim.apply(grad_scale=1.0 if self.dp is None else 1.0 / self.dp.world_size)
        return dict(losses=losses, first_frame=first)

    LOSS_KEYS = ("d_loss", "g_loss", "g_loss_first_frame", "images_std", "sampler_std", "real_D_std", "fake_D_std")

    def _step_device(self, images, z, disc_updates, gen_updates, loss_vec):
        self.img_dcgan.want_sigmoid = False
        for _ in range(disc_updates):
            d = self.d_update(images, z)
        for _ in range(gen_updates):
            g = self.g_update(z)
        self.img_dcgan.want_sigmoid = True
        one = lambda t: None if t is None else t.reshape(-1)[0:1]
        ops.cabi.gather_scalars([one(d["losses"]), one(g["losses"]), one(g["first_frame"])] +
                                [one(d[k]) for k in ("images_std", "sampler_std", "real_D_std", "fake_D_std")], loss_vec)

    def train_step(self, batch_images, batch_z, disc_updates=1, gen_updates=2, use_graph=True, sync=True):
        """The loop body of z_model_lib.py:217-239 for one batch of clips ([Bv*T, s, s, c] frames, [Bv, z_in] latents;
        host or device memory)."""
        n = self.batch_size * self.vid_length
        st = getattr(self, "_static", None)
        if st is None:
            dev = self.store.device
            st = self._static = dict(z=torch.empty((self.batch_size, self.z_input_size), dtype=torch.float32, device=dev),
                                     loss_dev=torch.zeros(7, dtype=torch.float32, device=dev),
                                     loss_host=torch.zeros(7, dtype=torch.float32).pin_memory())
        both = self._both()
        if self.dp is not None:
            self.dp.prepare(self.store)        # peer-memory exchange: collective handshake, first call only (never inside a capture)
        both[:n].copy_(torch.as_tensor(batch_images), non_blocking=True)
        st["z"].copy_(torch.as_tensor(batch_z), non_blocking=True)
        args = (both[:n], st["z"], disc_updates, gen_updates, st["loss_dev"])
        if use_graph:
            key = (disc_updates, gen_updates)
            g = self._graphs.get(key)
            if g is None:
                g = self._graphs[key] = self._capture(args)
            ops.refresh_packs(self.store)
            g["graph"].replay()
            self.d_optim.t += disc_updates
            self.g_optim.t += gen_updates
        else:
            self._step_device(*args)
        if not sync:
            return st["loss_dev"]
        st["loss_host"].copy_(st["loss_dev"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return {k: float(st["loss_host"][i]) for i, k in enumerate(self.LOSS_KEYS)}

    def _capture(self, args):
        snap = {k: t.clone() for k, t in self.store.flat.items()}
        states = (self.d_optim.state.clone(), self.g_optim.state.clone(), self.d_optim.t, self.g_optim.t)

        def restore():
            for k, t in snap.items():
                self.store.flat[k].copy_(t)
            self.d_optim.state.copy_(states[0]); self.g_optim.state.copy_(states[1])
            self.d_optim.t, self.g_optim.t = states[2], states[3]
            for v in self.store.vars.values():
                v.invalidate_packed()
            ops.refresh_packs(self.store)

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
        return dict(graph=graph, launches=launches)

    def sample(self, z, is_training=False):
        """sess.run(img_dcgan.sampler, {z, is_training}) (model_sampler.py:66-70, dump_sample)."""
        with torch.no_grad():
            G_out, _ = self.generator(z, train=is_training)
            return self.img_dcgan.generator(G_out, train=False)

    # ------------------------------------------------------------------------------
    def train(self, sess, config):
        """z_model_lib.py:148-259.  `config.video_list == ['synthetic']` trains on seeded random clips."""
        synthetic = list(getattr(config, "video_list", [])) in ([], ["synthetic"])
        files = []
        if not synthetic:
            for lst in config.video_list:
                with open(lst, 'r') as f:
                    for video in f:
                        video = video.strip()
                        if video:
                            files.append(os.path.join(config.video_data_dir, config.video_dataset, video))
            print("Total video files found:", len(files))
            if config.video_shuffle:
                np.random.shuffle(files)
        self.set_trainable(getattr(config, "train_img_gen", False), getattr(config, "train_img_disc", False))
        self.d_optim.lr = self.g_optim.lr = config.learning_rate
        self.d_optim.b1 = self.g_optim.b1 = config.beta1
        sample_z = np.random.uniform(-1, 1, size=(self.sample_rows * self.sample_cols, self.z_input_size)).astype(np.float32)
        counter = 0
        batch_size = self.batch_size
        n_batches = (len(files) // batch_size) if not synthetic else int(getattr(config, "synthetic_batches", 4))
        rs = np.random.RandomState(103)
        last = None
        for epoch in range(config.epoch):
            loader = None if synthetic else self.video_batches(files, batch_size)
            for i in range(n_batches):
                if synthetic:
                    s = self.output_image_size
                    batch_images = rs.uniform(-1, 1, (batch_size * self.vid_length, s, s, self.c_dim)).astype(np.float32)
                else:
                    batch_images = next(loader).reshape(-1, self.input_image_size, self.input_image_size, self.c_dim)
                batch_z = np.random.uniform(-1, 1, size=(self.batch_size, self.z_input_size)).astype(np.float32)
                last = self.train_step(batch_images, batch_z, config.disc_updates, config.gen_updates)
                counter += 1
                print("Epoch: [%2d] [%4d/%4d] d_loss: %s, g_loss: %s, first_frame_loss: %s"
                      % (epoch, i + 1, n_batches, [last["d_loss"]], [last["g_loss"]], last["g_loss_first_frame"]))
                print("Images std: %0.3f, sampler std: %0.3f | Real D std: %0.3f, fake D std: %0.3f" % (
                    last["images_std"], last["sampler_std"], last["real_D_std"], last["fake_D_std"]))
                if counter % config.sample_frequency == 0:
                    if getattr(config, "video_sample_dir", None):
                        self.dump_sample(sample_z, sess, config, epoch, i, is_training=False)
                    if getattr(config, "video_checkpoint_dir", None):
                        self.save_checkpoint(config.video_checkpoint_dir, counter)
        return last

    def dump_sample(self, sample_z, sess, config, epoch, idx, is_training=False, prefix=""):
        """z_model_lib.py:261-330: sample_rows x sample_cols grid of clips as an mp4 (25 fps)."""
        import cv2
        sz = self.output_image_size
        samples = self.sample(torch.as_tensor(sample_z, dtype=torch.float32).to(self.store.device), is_training).float().cpu().numpy()
        videos = np.reshape(samples, [self.sample_rows, self.sample_cols, self.vid_length, sz, sz, self.c_dim])
        folder = os.path.join(config.video_sample_dir, "train" if is_training else "inference")
        os.makedirs(folder, exist_ok=True)
        filename = '{}/{}train_{:02d}_{:04d}.mp4'.format(folder, prefix, epoch, idx)
        w = open_video_writer(filename, 25.0, (self.sample_cols * sz, self.sample_rows * sz))
        for t in range(self.vid_length):
            frame = np.zeros(shape=[self.sample_rows * sz, self.sample_cols * sz, self.c_dim], dtype=np.uint8)
            for r in range(self.sample_rows):
                for c in range(self.sample_cols):
                    im = np.around(inverse_transform(videos[r, c, t]) * 255).astype('uint8')
                    frame[r * sz:(r + 1) * sz, c * sz:(c + 1) * sz, :] = cv2.cvtColor(im, cv2.COLOR_RGB2BGR)
            w.write(frame)
        w.release()
        return filename

    def video_batches(self, files, batch_size, depth=2, workers=8):
        """The clip batches of z_model_lib.py:226-228 (get_videos on consecutive slices of the file list), decoded
        `depth` batches ahead on worker threads into page-locked buffers; each batch is [clips, T, s, s, c]."""
        from .input_pipeline import Prefetcher, chunks
        s = self.input_image_size
        return Prefetcher(chunks(files, batch_size), lambda f: self.load_videos([f]), (self.vid_length, s, s, self.c_dim),
                          depth=depth, workers=workers)

    def raw_video_batches(self, files, batch_size, frame_hw, depth=2, workers=8):
        """video_batches with the decode TAIL left to the GPU: the workers only decode (cv2.VideoCapture.read) and the batches are
        the raw uint8 BGR frames [clips, T, H0, W0, 3] at the files' own resolution `frame_hw` (one resolution per data set);
        `frames_from_raw` turns a batch into what `load_videos` returns -- resize, BGR -> RGB, / 127.5 - 1 -- in one launch."""
        from .input_pipeline import Prefetcher, chunks
        H0, W0 = frame_hw
        return Prefetcher(chunks(files, batch_size), lambda f: self.load_videos_raw([f], frame_hw)[0], (self.vid_length, H0, W0, self.c_dim),
                          depth=depth, workers=workers, dtype=torch.uint8)

    def load_videos_raw(self, files, frame_hw):
        """The decode half of z_model_lib.py:332-351: uint8 BGR frames [n, T, H0, W0, 3] exactly as cap.read() returns them."""
        import cv2
        H0, W0 = frame_hw
        videos = np.zeros((len(files), self.vid_length, H0, W0, self.c_dim), dtype=np.uint8)
        for (i, f) in enumerate(files):
            cap = cv2.VideoCapture(f)
            frame = 0
            while cap.isOpened() and frame < self.vid_length:
                ret, im = cap.read()
                if not ret:
                    break
                if im.shape != (H0, W0, self.c_dim):
                    raise ValueError(f"{f}: frame shape {im.shape}, expected {(H0, W0, self.c_dim)}")
                videos[i, frame] = im
                frame += 1
            assert frame == self.vid_length
        return videos

    def frames_from_raw(self, raw_u8, out=None):
        """raw uint8 BGR frames [clips, T, H0, W0, 3] (host, page-locked, or device) -> the float32 batch
        [clips * T, s, s, 3] of load_videos, bit for bit (ops.frames_to_input: the H2D copy moves bytes, the rest is one launch)."""
        t = raw_u8 if isinstance(raw_u8, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(raw_u8))
        t = t.reshape(-1, t.shape[-3], t.shape[-2], t.shape[-1])
        if not t.is_cuda:
            t = t.to(self.store.device, non_blocking=True)
        return ops.frames_to_input(t, self.input_image_size, swap_rb=True, out=out)

    def load_videos(self, files):
        """z_model_lib.py:332-351."""
        import cv2
        n = len(files)
        videos = np.zeros(shape=(n, self.vid_length, self.input_image_size, self.input_image_size, self.c_dim))
        for (i, f) in enumerate(files):
            cap = cv2.VideoCapture(f)
            frame = 0
            while cap.isOpened() and frame < self.vid_length:
                ret, im = cap.read()
                if not ret:
                    break
                im = cv2.resize(im, (self.input_image_size, self.input_image_size), interpolation=cv2.INTER_LINEAR)
                im = cv2.cvtColor(im, cv2.COLOR_BGR2RGB)
                videos[i, frame] = transform(im, is_crop=False)
                frame += 1
            assert frame == self.vid_length
        return np.reshape(videos, [n * self.vid_length, self.input_image_size, self.input_image_size, self.c_dim])

    # ---- checkpoints (keys = TF variable names) -----------------------
    def save_checkpoint(self, checkpoint_dir, step):
        os.makedirs(checkpoint_dir, exist_ok=True)
        name = "VID_DCGAN.model-%d" % step
        f = self.store.flat
        torch.save(dict(variables=self.store.state_dict(), adam_m=f["m"].cpu(), adam_v=f["v"].cpu(), d_t=self.d_optim.t, g_t=self.g_optim.t),
                   os.path.join(checkpoint_dir, name))
        with open(os.path.join(checkpoint_dir, "checkpoint"), "w") as fh:
            fh.write('model_checkpoint_path: "%s"\n' % name)

    @staticmethod
    def _latest(checkpoint_dir):
        index = os.path.join(checkpoint_dir, "checkpoint")
        if not os.path.exists(index):
            return None
        with open(index) as f:
            return os.path.join(checkpoint_dir, os.path.basename(f.readline().split('"')[1]))   # z_model_lib.py:122-123

    def load_image_gan(self, sess, checkpoint_dir):
        """z_model_lib.py:117-134: restore the nested image GAN from a DCGAN checkpoint (prefix-stripped names)."""
        path = self._latest(checkpoint_dir)
        if path is None:
            print("FAIL!")
            return False
        from . import checkpoint_io
        if checkpoint_io.tf_format(path):                      # a TensorFlow checkpoint of the image GAN
            checkpoint_io.load_tf_checkpoint(path, self.store, prefix=self.image_gan_scope_name)
        else:
            payload = torch.load(path, map_location="cpu", weights_only=True)
            self.store.load_state_dict(payload["variables"], strict=True, prefix=self.image_gan_scope_name)
        print("Success!")
        return True

    def load_checkpoint(self, sess, checkpoint_dir):
        """z_model_lib.py:136-146."""
        path = self._latest(checkpoint_dir)
        if path is None:
            print("FAIL!")
            return False
        from . import checkpoint_io
        if checkpoint_io.tf_format(path):
            checkpoint_io.load_tf_checkpoint(path, self.store)
        else:
            payload = torch.load(path, map_location="cpu", weights_only=True)
            self.store.load_state_dict(payload["variables"])
        print("Success!")
        return True
