"""CPU oracle -- latent search.  TEST INFRASTRUCTURE ONLY (PARITY UNPINNED by the reference: the reference ships no
vectors for this path and TensorFlow 0.12 cannot run here; see oracle/tf_ops.py header).

Restates, on top of oracle/models.py::DCGAN and oracle/tf_ops.py, the loss graph and optimiser that
  * /root/reference/models/recurrent_z/z_space_finder.py:226-298 (losses, optimiser) and :122-160 (schedule)
  * /root/reference/models/recurrent_z/discriminator_activation_optimizer.py:151-217, 231-276
build on a trained DCGAN:  loss = sum_k w_k * term_k over
  activations L2 / L1:  reduce_mean(reduce_mean(square|abs(D_h2(G(z)) - D_h2(target)), [1,2,3]))
  pixel       L2 / L1:  the same on G(z) - target
  generator loss:       reduce_mean(sigmoid_cross_entropy_with_logits(D_logits(G(z)), 1))
minimised over z alone with tf.train.AdamOptimizer(lr_tensor, beta1); `train` mode uses the batch-statistics graph
(G, D_activations_, g_loss), `inference` the moving-average one (sampler, D_activations_inf_, g_loss_inf).
"""
from __future__ import annotations

import numpy as np
import torch

from . import tf_ops as T

WEIGHT_NAMES = ("pixel_L2_weight", "pixel_L1_weight", "activations_L2_weight", "activations_L1_weight", "generator_loss_weight")


def make_trained_like(dcgan, seed=11, gain=8.0):
    """Test helper: give a freshly initialised oracle DCGAN the look of a trained one -- seeded moving averages
    (mean ~ N(0, 0.1), variance ~ U(0.5, 1.5)), gamma ~ N(1, 0.1) and filters scaled by `gain` (sigma
    0.02 -> 0.16) so that inference-mode batch norm, activations and latent gradients have a useful size."""
    rs = np.random.RandomState(seed)
    with torch.no_grad():
        for k, v in dcgan.vars.items():
            if k.endswith("moving_mean"):
                v.copy_(torch.tensor(rs.normal(0, 0.1, tuple(v.shape)), dtype=v.dtype))
            elif k.endswith("moving_variance"):
                v.copy_(torch.tensor(rs.uniform(0.5, 1.5, tuple(v.shape)), dtype=v.dtype))
            elif k.endswith("gamma"):
                v.copy_(torch.tensor(rs.normal(1, 0.1, tuple(v.shape)), dtype=v.dtype))
            elif k.endswith("/w") or k.endswith("Matrix"):
                v.mul_(gain)
    return dcgan


class LatentSearch:
    def __init__(self, dcgan, discriminator_mode="inference", pixel_L2_weight=0.0, pixel_L1_weight=0.0, activations_L2_weight=1.0,
                 activations_L1_weight=0.0, generator_loss_weight=0.0, beta1=0.5, z=None, random_seed=0):
        self.dcgan, self.train = dcgan, discriminator_mode == "train"
        self.w = dict(zip(WEIGHT_NAMES, (pixel_L2_weight, pixel_L1_weight, activations_L2_weight, activations_L1_weight, generator_loss_weight)))
        if z is None:
            z = np.random.RandomState(random_seed).uniform(-1.0, 1.0, size=(dcgan.batch_size, dcgan.z_dim))
        self.z = torch.tensor(np.asarray(z), dtype=dcgan.dtype)
        self.optim = T.TFAdam({"z": self.z}, 0.0, beta1)        # one optimiser for the whole run; lr is fed per step

    def _t(self, a):
        return torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).to(self.dcgan.dtype)

    def target_activations(self, images):
        """sess.run(D_activations | D_activations_inf, {images: targets})  (z_space_finder.py:127-131)"""
        with torch.no_grad():
            return self.dcgan.discriminator(self._t(images), train=self.train, tag="d_target")[2].clone()

    def images(self):
        with torch.no_grad():
            return self.dcgan.generator(self.z, train=self.train, tag="s")

    def loss_terms(self, z, target_images, target_activations):
        """The five terms of z_space_finder.py:258-292, each already multiplied by its weight (zero-weight terms are
        left out: their gradient contribution in the reference is exactly zero)."""
        w, d = self.w, self.dcgan
        G = d.generator(z, train=self.train)
        terms = {}
        if w["activations_L2_weight"] or w["activations_L1_weight"] or w["generator_loss_weight"]:
            _, logits, h2 = d.discriminator(G, train=self.train, tag="d_fake")
            diff = h2 - self._t(target_activations)
            if w["activations_L2_weight"]:
                terms["activations_L2"] = w["activations_L2_weight"] * diff.square().mean(dim=(1, 2, 3)).mean()
            if w["activations_L1_weight"]:
                terms["activations_L1"] = w["activations_L1_weight"] * diff.abs().mean(dim=(1, 2, 3)).mean()
            if w["generator_loss_weight"]:
                terms["generator"] = w["generator_loss_weight"] * d._ce(logits, 1.0)
        if w["pixel_L2_weight"] or w["pixel_L1_weight"]:
            pd = G - self._t(target_images)
            if w["pixel_L2_weight"]:
                terms["pixel_L2"] = w["pixel_L2_weight"] * pd.square().mean(dim=(1, 2, 3)).mean()
            if w["pixel_L1_weight"]:
                terms["pixel_L1"] = w["pixel_L1_weight"] * pd.abs().mean(dim=(1, 2, 3)).mean()
        return terms

    def loss_and_grad(self, target_images, target_activations):
        self.dcgan.set_requires_grad(set())
        z = self.z.detach().clone().requires_grad_(True)
        loss = sum(self.loss_terms(z, target_images, target_activations).values())
        loss.backward()
        return float(loss.item()), z.grad.detach().clone()

    def step(self, target_images, target_activations, lr):
        """sess.run([optim, loss], {lr_tensor: lr, ...}): the loss before the update; z advanced by one TF-Adam step."""
        loss, g = self.loss_and_grad(target_images, target_activations)
        self.optim.lr = lr
        self.optim.apply({"z": g})
        return loss

    def fit_video(self, targets, num_initial_steps, num_steps_per_frame, learning_rate, lr_decay_amount):
        """process_batch, z_space_finder.py:122-160."""
        targets = np.asarray(targets)
        B, Tn = targets.shape[:2]
        acts = [self.target_activations(targets[:, f]) for f in range(Tn)]
        results = np.zeros(targets.shape, dtype=np.float64)
        zs = np.zeros((B, Tn, self.dcgan.z_dim), dtype=np.float64)
        lr = learning_rate
        losses = []
        for _ in range(num_initial_steps):
            losses.append(self.step(targets[:, 0], acts[0], lr))
        results[:, 0], zs[:, 0] = self.images().numpy(), self.z.numpy()
        lr *= lr_decay_amount
        for f in range(Tn):
            for _ in range(num_steps_per_frame):
                losses.append(self.step(targets[:, f], acts[f], lr))
            results[:, f], zs[:, f] = self.images().numpy(), self.z.numpy()
        return results, zs, losses


class NestedLatentSearch(LatentSearch):
    """/root/reference/models/recurrent_z/discriminator_activation_optimizer_nested.py:146-217: the variable is the VIDEO latent
    z [clips, 120]; frames = image_gan.generator(video_generator(z)); activation and pixel terms on the first frame of every clip
    (`[::vid_length]`), generator term = the image GAN's g_loss over all frames; target activations from a batch holding the targets
    in the frame-0 slots and zeros elsewhere (lines 148-157); `train` also switches the video generator's batch norm (line 242)."""

    def __init__(self, vid, discriminator_mode="inference", z=None, random_seed=0, **kw):
        self.vid, self.Tn = vid, vid.vid_length
        if z is None:
            z = np.random.RandomState(random_seed).uniform(-1.0, 1.0, size=(vid.batch_size, vid.z_input_size))
        super().__init__(vid.img_dcgan, discriminator_mode=discriminator_mode, z=z, **kw)

    def _frames(self, z):
        return self.vid.img_dcgan.generator(self.vid.generator(z, train=self.train), train=self.train, tag="s")

    def target_activations(self, images):
        t = self._t(images)
        full = torch.zeros((t.shape[0] * self.Tn,) + tuple(t.shape[1:]), dtype=t.dtype)
        full[::self.Tn] = t
        with torch.no_grad():
            return self.dcgan.discriminator(full, train=self.train, tag="d_target")[2][::self.Tn].clone()

    def images(self):
        with torch.no_grad():
            return self._frames(self.z)

    def loss_terms(self, z, target_images, target_activations):
        w, d, Tn = self.w, self.dcgan, self.Tn
        G = self._frames(z)
        terms = {}
        if w["activations_L2_weight"] or w["activations_L1_weight"] or w["generator_loss_weight"]:
            _, logits, h2 = d.discriminator(G, train=self.train, tag="d_fake")
            diff = h2[::Tn] - self._t(target_activations)
            if w["activations_L2_weight"]:
                terms["activations_L2"] = w["activations_L2_weight"] * diff.square().mean(dim=(1, 2, 3)).mean()
            if w["activations_L1_weight"]:
                terms["activations_L1"] = w["activations_L1_weight"] * diff.abs().mean(dim=(1, 2, 3)).mean()
            if w["generator_loss_weight"]:
                terms["generator"] = w["generator_loss_weight"] * d._ce(logits, 1.0)
        if w["pixel_L2_weight"] or w["pixel_L1_weight"]:
            pd = G[::Tn] - self._t(target_images)
            if w["pixel_L2_weight"]:
                terms["pixel_L2"] = w["pixel_L2_weight"] * pd.square().mean(dim=(1, 2, 3)).mean()
            if w["pixel_L1_weight"]:
                terms["pixel_L1"] = w["pixel_L1_weight"] * pd.abs().mean(dim=(1, 2, 3)).mean()
        return terms

    def loss_and_grad(self, target_images, target_activations):
        self.vid.set_requires_grad(set())
        z = self.z.detach().clone().requires_grad_(True)
        loss = sum(self.loss_terms(z, target_images, target_activations).values())
        loss.backward()
        return float(loss.item()), z.grad.detach().clone()
