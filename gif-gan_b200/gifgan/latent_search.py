"""Latent search: find the z whose generated image matches a target image, in pixels and / or in the image
discriminator's `h2` activations -- the loss graph and optimiser that /root/reference/models/recurrent_z/z_space_finder.py
(:226-298) and discriminator_activation_optimizer*.py (:151-217) build on top of a trained DCGAN:

    loss = aL2 * mean((D_h2(G(z)) - D_h2(target))^2) + aL1 * mean|...|
         + pL2 * mean((G(z) - target)^2)             + pL1 * mean|...|
         + gW  * mean(sigmoid_ce(D_logits(G(z)), 1))
    optim = tf.train.AdamOptimizer(lr_tensor, beta1).minimize(loss, var_list=[z])

with the five weights normalised to sum 1 and `discriminator_mode` choosing the train-mode graph (G, D_activations_,
g_loss: batch statistics) or the inference one (sampler, D_activations_inf_, g_loss_inf: moving averages).  On the device
one step is: generator forward, discriminator forward (to h2, or to the logits when gW > 0), the distance-loss launches
(gg_distance_loss writes the loss and its gradient together), the input-gradient kernels back to z -- no filter
gradients: the var_list is [z] -- and one TF-Adam launch on z.

Terms whose normalised weight is zero are not evaluated (the reference still evaluates them and multiplies by 0.0; the
only side effect of that is a d_bn3 moving-average update in train mode, which none of these tools read back).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .ops import add_noise, distance_loss, sigmoid_cross_entropy_loss

WEIGHT_NAMES = ("pixel_L2_weight", "pixel_L1_weight", "activations_L2_weight", "activations_L1_weight", "generator_loss_weight")


def normalised_weights(opts):
    """z_space_finder.py:229-238: divide the five loss weights by their sum."""
    total = sum(float(getattr(opts, k)) for k in WEIGHT_NAMES)
    if total <= 0:
        raise ValueError("the loss weights must have a positive sum")
    return {k: float(getattr(opts, k)) / total for k in WEIGHT_NAMES}


class LatentSearch(object):
    def __init__(self, dcgan, discriminator_mode="inference", pixel_L2_weight=0.0, pixel_L1_weight=0.0, activations_L2_weight=1.0,
                 activations_L1_weight=0.0, generator_loss_weight=0.0, beta1=0.5, beta2=0.999, epsilon=1e-8, random_seed=0, z=None,
                 use_graph=False):
        """`dcgan`: a gifgan.model.DCGAN (unconditional) holding the trained weights; its batch_size is the number of
        latents searched at once.  The weights are taken as given (call normalised_weights first for CLI behaviour).
        `z`: initial latents [batch, z_dim]; default uniform(-1, 1) (tf.random_uniform_initializer, z_space_finder.py:46).
        `use_graph`: replay one captured CUDA graph per step (one capture per learning-rate value) instead of ~60 eager
        launches; the targets are then copied into fixed device buffers before each replay."""
        if discriminator_mode not in ("train", "inference"):
            raise ValueError("discriminator_mode must be 'train' or 'inference'")
        if dcgan.y_dim:
            raise ValueError("latent search is defined for the unconditional DCGAN only")
        self.dcgan, self.train = dcgan, discriminator_mode == "train"
        self.w = dict(pixel_L2_weight=float(pixel_L2_weight), pixel_L1_weight=float(pixel_L1_weight),
                      activations_L2_weight=float(activations_L2_weight), activations_L1_weight=float(activations_L1_weight),
                      generator_loss_weight=float(generator_loss_weight))
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        dev = dcgan.store.device
        zshape = self._z_shape()
        if z is None:
            z = np.random.RandomState(random_seed).uniform(-1.0, 1.0, size=zshape)
        self.z = torch.as_tensor(np.asarray(z, dtype=np.float32)).to(dev).contiguous().requires_grad_(True)
        if tuple(self.z.shape) != tuple(zshape):
            raise ValueError(f"z must have shape {tuple(zshape)}")
        # Adam slots of the optimiser built once for the whole run (z_space_finder.py:294-298): they, like z, carry over
        # from one target to the next
        self.m, self.v, self.t = torch.zeros_like(self.z), torch.zeros_like(self.z), 0
        self.state = torch.zeros(4, dtype=torch.int32, device=dev)      # [t, lr_t bits, ticket, -] on the device (gg_adam_graph)
        self.use_graph, self._graphs, self._bufs = use_graph, {}, {}
        self._ones = torch.ones(1, dtype=torch.float32, device=dev)
        self._loss_vec = torch.zeros(8, dtype=torch.float32, device=dev)

    def _z_shape(self):
        return (self.dcgan.batch_size, self.dcgan.z_dim)

    # ------------------------------------------------------------------------------------------
    def _target(self, a):
        return torch.as_tensor(np.asarray(a, dtype=np.float32) if not torch.is_tensor(a) else a).to(self.z.device, torch.float32).contiguous()

    def target_activations(self, images):
        """sess.run(D_activations | D_activations_inf, {images: targets}) (z_space_finder.py:127-131): float32 [B, s/8, s/8, 4*df]."""
        with torch.no_grad():
            h2 = self.dcgan.discriminator(add_noise(self._target(images), self.dcgan.noise_std), reuse=True, train=self.train, stop_at_h2=True)[2]
        return h2.float().contiguous()

    def images(self):
        """sess.run(G | sampler): the images of the current latents, float32 [B, s, s, c] in (-1, 1)."""
        with torch.no_grad():
            return self.dcgan.generator(self.z.detach(), train=self.train)

    def assign(self, z):
        """sess.run(dcgan.z.assign(z)): overwrite the latents (the Adam slots keep running, as in the reference's
        discriminator_activation_optimizer_video.py:228-231, 243-248, which copies frame 0's latents over the later frames)."""
        with torch.no_grad():
            self.z.copy_(torch.as_tensor(np.asarray(z, dtype=np.float32)).reshape(self.z.shape))

    def lr_t(self, lr):
        return lr * float(np.sqrt(1.0 - self.beta2 ** self.t)) / (1.0 - self.beta1 ** self.t)

    def loss_and_grad(self, target_images, target_activations):
        """Loss terms (device float32 [1] each, already weighted) and d loss / d z in self.z.grad."""
        w, d = self.w, self.dcgan
        tgt_img = self._target(target_images) if (w["pixel_L2_weight"] or w["pixel_L1_weight"]) else None
        need_act = bool(w["activations_L2_weight"] or w["activations_L1_weight"])
        need_gen = bool(w["generator_loss_weight"])
        self.z.grad = None
        roots = []
        with ops.trainable([]), ops.stats_arena():
            G = d.generator(self.z, train=self.train)
            if need_act or need_gen:
                out = d.discriminator(add_noise(G, d.noise_std), reuse=True, train=self.train, stop_at_h2=not need_gen)
                if need_act:
                    roots.append(distance_loss(out[2], self._target(target_activations), w["activations_L2_weight"], w["activations_L1_weight"]))
                if need_gen:
                    B = out[1].shape[0]
                    roots.append(sigmoid_cross_entropy_loss(out[1], [(0, B, 1.0, w["generator_loss_weight"])])[0:1])
            if tgt_img is not None:
                roots.append(distance_loss(G, tgt_img, w["pixel_L2_weight"], w["pixel_L1_weight"]))
            torch.autograd.backward(roots, grad_tensors=[self._ones] * len(roots))
        return roots

    def _step_device(self, target_images, target_activations, lr):
        roots = self.loss_and_grad(target_images, target_activations)
        z = self.z.detach()
        ops.check(ops.cabi.lib().gg_adam_graph(ops.ptr(z), None, ops.ptr(self.z.grad), ops.ptr(self.m), ops.ptr(self.v), z.numel(),
                                               ops.ptr(self.state), float(lr), self.beta1, self.beta2, self.epsilon, 1.0, ops.stream()),
                  "gg_adam_graph")
        ops.cabi.gather_scalars(roots, self._loss_vec)
        return len(roots)

    def _static(self, name, src):
        t = self._target(src)
        buf = self._bufs.get(name)
        if buf is None or buf.shape != t.shape:
            buf = self._bufs[name] = torch.empty_like(t)
            self._graphs.clear()
        buf.copy_(t, non_blocking=True)
        return buf

    def _capture(self, lr):
        """Capture one step at learning rate `lr`.  The warm-up run (allocator, lazy filter packs) and the capture must
        not move the search: latents, Adam state and the moving averages are put back afterwards."""
        img, act = self._bufs.get("img"), self._bufs.get("act")      # None for terms that are switched off
        flat = self.dcgan.store.flat["params"]
        snap = [t.clone() for t in (flat, self.z.detach(), self.m, self.v, self.state)]

        def restore():
            for dst, src in zip((flat, self.z.detach(), self.m, self.v, self.state), snap):
                dst.copy_(src)

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._step_device(img, act, lr)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        restore()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            n = self._step_device(img, act, lr)
        restore()
        torch.cuda.synchronize()
        g = self._graphs[float(lr)] = dict(graph=graph, nroots=n)
        return g

    def step(self, target_images, target_activations, lr, fetch_loss=True, use_graph=None):
        """sess.run([optim, loss], {lr_tensor, activations_placeholder, target_placeholder}): one Adam step on z.
        Returns the loss at the latents BEFORE the update (a Python float; None with fetch_loss=False)."""
        if self.use_graph if use_graph is None else use_graph:
            if self.w["pixel_L2_weight"] or self.w["pixel_L1_weight"]:
                self._static("img", target_images)
            if self.w["activations_L2_weight"] or self.w["activations_L1_weight"]:
                self._static("act", target_activations)
            g = self._graphs.get(float(lr)) or self._capture(lr)
            g["graph"].replay()
            n = g["nroots"]
        else:
            n = self._step_device(target_images, target_activations, lr)
        self.t += 1
        if not fetch_loss:
            return None
        return float(self._loss_vec[:n].sum().item())

    # ------------------------------------------------------------------------------------------
    def fit_video(self, targets, num_initial_steps=500, num_steps_per_frame=100, learning_rate=0.05, lr_decay_amount=0.5, log=None):
        """z_space_finder.py:122-160 (process_batch): `targets` [B, T, s, s, c] in (-1, 1).  Searches frame 0 from the
        current latents for num_initial_steps, multiplies the learning rate by lr_decay_amount, then tracks every frame
        in turn (frame 0 again first) for num_steps_per_frame steps each, warm-starting from the previous frame's
        latents.  Returns (images [B, T, s, s, c], latents [B, T, z_dim]) as float32 numpy arrays."""
        targets = np.asarray(targets, dtype=np.float32)
        B, T = targets.shape[:2]
        acts = [self.target_activations(targets[:, f]) for f in range(T)]
        imgs = [self._target(targets[:, f]) for f in range(T)]
        results = np.zeros(targets.shape, dtype=np.float32)
        zs = np.zeros((B, T, self.dcgan.z_dim), dtype=np.float32)
        total = num_initial_steps + num_steps_per_frame * T
        lr = learning_rate
        for i in range(num_initial_steps):
            loss = self.step(imgs[0], acts[0], lr, fetch_loss=log is not None)
            if log is not None:
                log("Step %d/%d: loss %f" % (i, total, loss))
        results[:, 0], zs[:, 0] = self.images().float().cpu().numpy(), self.z.detach().cpu().numpy()
        lr *= lr_decay_amount
        for f in range(T):
            for i in range(num_steps_per_frame):
                loss = self.step(imgs[f], acts[f], lr, fetch_loss=log is not None)
                if log is not None:
                    log("Step %d/%d: loss %f" % (num_initial_steps + num_steps_per_frame * f + i, total, loss))
            results[:, f], zs[:, f] = self.images().float().cpu().numpy(), self.z.detach().cpu().numpy()
        return results, zs

    def optimise(self, targets, num_steps=1000, learning_rate=0.0002, lr_decay_frequency=0, lr_decay_amount=0.9, on_step=None):
        """discriminator_activation_optimizer.py:151-157, 231-276: match one batch of target images for num_steps steps;
        every lr_decay_frequency steps (when > 0) the learning rate is multiplied by lr_decay_amount.  `on_step(i, loss,
        images)` is called after each step when given (sample / progress-video writers).  Returns the final images."""
        acts = self.target_activations(targets)
        tgt = self._target(targets)
        lr = learning_rate
        for i in range(num_steps):
            loss = self.step(tgt, acts, lr, fetch_loss=on_step is not None)
            if on_step is not None:
                on_step(i, loss, self)
            if lr_decay_frequency > 0 and i % lr_decay_frequency == lr_decay_frequency - 1:
                lr *= lr_decay_amount
        return self.images().float().cpu().numpy()


class NestedLatentSearch(LatentSearch):
    """discriminator_activation_optimizer_nested.py: the search runs THROUGH the video generator.  The variable is the video
    latent z [clips, z_input_size] (120); the clip is vid.generator(z) -> [clips * T, 100] image latents -> the image GAN's
    generator -> [clips * T, s, s, c] frames -> the image discriminator; the activation and pixel terms compare only the FIRST
    frame of every clip with the target (`[::vid_length]`, reference lines 180-198), the generator term is the image GAN's g_loss
    over all frames (line 192).  `discriminator_mode` also sets the video generator's batch-norm mode (`is_training`, line 242).
    Batch-norm statistics in train mode run over all clips * T frames, so the target activations are taken, as in the reference
    (lines 148-157), from a batch that holds the targets in the frame-0 slots and zeros elsewhere."""

    def __init__(self, vid, **kw):
        self.vid, self.T = vid, vid.vid_length
        super().__init__(vid.img_dcgan, **kw)

    def _z_shape(self):
        return (self.vid.batch_size, self.vid.z_input_size)

    def _frames(self, z):
        G_out, _ = self.vid.generator(z, train=self.train)
        return self.vid.img_dcgan.generator(G_out, train=self.train)

    def target_activations(self, images):
        t = self._target(images)
        full = torch.zeros((t.shape[0] * self.T,) + tuple(t.shape[1:]), dtype=torch.float32, device=t.device)
        full[::self.T] = t
        with torch.no_grad():
            h2 = self.dcgan.discriminator(add_noise(full, self.dcgan.noise_std), reuse=True, train=self.train, stop_at_h2=True)[2]
        return h2[::self.T].float().contiguous()

    def images(self):
        """All frames of the current clips, float32 [clips * T, s, s, c]; `[::T]` are the searched first frames."""
        with torch.no_grad():
            return self._frames(self.z.detach())

    def loss_and_grad(self, target_images, target_activations):
        w, d, T_ = self.w, self.dcgan, self.T
        tgt_img = self._target(target_images) if (w["pixel_L2_weight"] or w["pixel_L1_weight"]) else None
        need_act = bool(w["activations_L2_weight"] or w["activations_L1_weight"])
        need_gen = bool(w["generator_loss_weight"])
        self.z.grad = None
        roots = []
        with ops.trainable([]), ops.stats_arena():
            G = self._frames(self.z)
            if need_act or need_gen:
                out = d.discriminator(add_noise(G, d.noise_std), reuse=True, train=self.train, stop_at_h2=not need_gen)
                if need_act:
                    roots.append(distance_loss(out[2][::T_].contiguous(), self._target(target_activations), w["activations_L2_weight"],
                                               w["activations_L1_weight"]))
                if need_gen:
                    n = out[1].shape[0]
                    roots.append(sigmoid_cross_entropy_loss(out[1], [(0, n, 1.0, w["generator_loss_weight"])])[0:1])
            if tgt_img is not None:
                roots.append(distance_loss(G[::T_].contiguous(), tgt_img, w["pixel_L2_weight"], w["pixel_L1_weight"]))
            torch.autograd.backward(roots, grad_tensors=[self._ones] * len(roots))
        return roots


# ---- what the two programs share ----------------------------------------------------------------------
def load_dcgan(opts, batch_size):
    """load_dcgan of z_space_finder.py:43-68 / discriminator_activation_optimizer.py:57-82: an image DCGAN of
    `batch_size` latents restored from the newest checkpoint named by <checkpoint_directory>/checkpoint."""
    import os
    from .model import DCGAN
    ops.set_precision(opts.precision)
    ops.reset_default_store()
    dcgan = DCGAN(None, image_size=opts.image_size, batch_size=batch_size, output_size=opts.output_size, c_dim=opts.c_dim,
                  dataset_name='', is_crop=False, checkpoint_dir='', sample_dir='', data_dir='', log_dir='', image_glob='', shuffle=False)
    if opts.checkpoint_directory:
        index = os.path.join(opts.checkpoint_directory, "checkpoint")
        if not os.path.exists(index):
            raise IOError("no checkpoint index in %s" % opts.checkpoint_directory)
        with open(index) as f:
            name = os.path.basename(f.readline().split('"')[1])
        path = os.path.join(opts.checkpoint_directory, name)
        from . import checkpoint_io
        if checkpoint_io.tf_format(path):                      # a TensorFlow checkpoint (V2 bundle or V1 file)
            checkpoint_io.load_tf_checkpoint(path, dcgan.store, (dcgan.d_optim, dcgan.g_optim))
        else:
            dcgan.load_payload(torch.load(path, map_location="cpu", weights_only=True))
    elif not opts.synthetic:
        raise ValueError("--checkpoint_directory is required (or --synthetic n to run on random weights and targets)")
    return dcgan


def search_from_options(dcgan, opts):
    if opts.discriminator_mode not in ("train", "inference"):
        raise ValueError("--discriminator_mode must be train or inference")
    w = normalised_weights(opts)
    print("Normalized loss weights:")
    for k in WEIGHT_NAMES:
        print(k, w[k])
    return LatentSearch(dcgan, opts.discriminator_mode, beta1=opts.beta1, random_seed=opts.random_seed, use_graph=opts.cuda_graph, **w)


def read_video_frames(path, image_size, vid_length, skip):
    """load_video of z_space_finder.py:70-88: every `skip`-th frame, resized to image_size (bilinear), RGB, x/127.5 - 1.
    None when the file has fewer than vid_length * skip frames."""
    import cv2
    cap = cv2.VideoCapture(path)
    frames = []
    for _ in range(vid_length):
        im = None
        for _ in range(skip):
            ok, im = cap.read() if cap.isOpened() else (False, None)
            if not ok:
                print("Video %s not long enough! Skipping!" % path)
                return None
        im = cv2.cvtColor(cv2.resize(im, (image_size, image_size), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2RGB)
        frames.append(im.astype(np.float64) / 127.5 - 1.)
    return frames
