"""Entry point of the video GAN (the role of models/recurrent_z/z_model.py; options: flags.TABLES["video_gan"]).
Without a clip list it trains on seeded random clips."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gifgan import flags, ops  # noqa: E402
from gifgan.utils import pp  # noqa: E402
from gifgan.z_model_lib import VID_DCGAN  # noqa: E402

VID_Z_DIM, IMAGE_Z_DIM = 120, 100          # latent sizes fixed by the reference (z_model.py:64-65)


def build(opts):
    """VID_DCGAN under the 'video_gan' scope (the checkpoint keys start with it), image GAN loaded from image_model_dir."""
    with ops.variable_scope("video_gan"):
        gan = VID_DCGAN(None, opts.vid_batch_size, VID_Z_DIM, IMAGE_Z_DIM, opts.vid_length, opts.image_size, opts.output_size,
                        c_dim=opts.c_dim, image_noise_std=opts.image_noise, activation_noise_std=opts.activation_noise,
                        first_frame_loss_scalar=opts.first_frame_loss_scalar)
    gan.load_image_gan(None, opts.image_model_dir)      # after initialisation, so the loaded weights survive
    return gan


def shape_check(gan, opts):
    """The reference prints the shapes of one latent -> clip pass before training; same here."""
    z = torch.as_tensor(np.random.uniform(-1, 1, size=(opts.vid_batch_size, VID_Z_DIM)).astype(np.float32)).to(gan.store.device)
    with torch.no_grad():
        latents = gan.generator(z)[0]
        frames = gan.img_dcgan.sampler(latents)
    print(tuple(latents.shape), tuple(frames.shape), "expected clip tensor",
          (opts.vid_batch_size, opts.vid_length, opts.output_size, opts.output_size, opts.c_dim))


def main(argv=None):
    opts = flags.parse("video_gan", argv)
    pp.pprint(vars(opts))
    ops.set_precision(opts.precision)
    ops.reset_default_store()
    gan = build(opts)
    shape_check(gan, opts)
    if opts.is_train:
        gan.train(None, opts)
    return gan


if __name__ == "__main__":
    main()
