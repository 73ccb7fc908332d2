"""Sampling entry point (the role of models/recurrent_z/model_sampler.py; options: flags.TABLES["sampler"]): loads a
video-GAN checkpoint and writes one animated GIF per generated clip."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gifgan import flags, ops, utils  # noqa: E402
from gifgan.z_model import IMAGE_Z_DIM, VID_Z_DIM  # noqa: E402
from gifgan.z_model_lib import VID_DCGAN  # noqa: E402


def write_gif(video, filename, fps=25):
    """[-1, 1] float frames [T, H, W, C] -> animated GIF at 25 fps (PIL is the writer available in this image)."""
    from PIL import Image
    frames = [Image.fromarray(f if f.shape[-1] == 3 else f[..., 0]) for f in ((video + 1) * 127.5).astype(np.uint8)]
    frames[0].save(filename, save_all=True, append_images=frames[1:], duration=int(1000 / fps), loop=0)


def sample_batch(gan):
    """One batch of clips as [clips, T, H, W, C] numpy (inference-mode generator path)."""
    z = np.random.uniform(-1, 1, size=(gan.batch_size, VID_Z_DIM)).astype(np.float32)
    frames = gan.sample(torch.as_tensor(z).to(gan.store.device), is_training=False).float().cpu().numpy()
    return frames.reshape(gan.batch_size, gan.vid_length, gan.output_image_size, gan.output_image_size, gan.c_dim)


def main(argv=None):
    opts = flags.parse("sampler", argv)
    utils.pp.pprint(vars(opts))
    np.random.seed(opts.random_seed)
    os.makedirs(opts.output_directory, exist_ok=True)
    ops.set_precision(opts.precision)
    ops.reset_default_store()
    with ops.variable_scope("video_gan"):
        gan = VID_DCGAN(None, opts.vid_batch_size, VID_Z_DIM, IMAGE_Z_DIM, opts.vid_length, opts.image_size, opts.output_size,
                        c_dim=opts.c_dim, image_noise_std=opts.image_noise, activation_noise_std=opts.activation_noise)
    gan.load_checkpoint(None, opts.checkpoint_dir)
    passes = 0
    while True:
        for first in range(0, opts.num_samples, gan.batch_size):
            clips = sample_batch(gan)
            for n, clip in enumerate(clips[:opts.num_samples - first]):
                tmp = os.path.join(opts.output_directory, "tmp.gif")       # write-then-rename: readers never see a partial file
                write_gif(clip, tmp)
                os.rename(tmp, os.path.join(opts.output_directory, "%d.gif" % (first + n)))
        passes += 1
        if not opts.continuous:
            return gan
        print("finished pass %d" % passes)


if __name__ == "__main__":
    main()
