"""Sample CLI -- drop-in for /root/reference/models/recurrent_z/model_sampler.py (flags model_sampler.py:9-24):
loads a VID_DCGAN checkpoint and writes one GIF per generated clip."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gifgan import ops, utils  # noqa: E402
from gifgan.flags import Flags  # noqa: E402
from gifgan.z_model_lib import VID_DCGAN  # noqa: E402

flags = Flags()
flags.DEFINE_integer("vid_batch_size", 64, "The size of batch videos [64]")
flags.DEFINE_integer("vid_length", 16, "The length of the videos [16]")
flags.DEFINE_integer("image_size", 64, "The size of images used [64]")
flags.DEFINE_integer("output_size", 64, "The size of the output images to produce [64]")
flags.DEFINE_integer("c_dim", 3, "Dimension of image color. [3]")
flags.DEFINE_float("image_noise", 0.0, "Std of noise to add to images")
flags.DEFINE_float("activation_noise", 0.0, "Std of noise to add to D activations")
flags.DEFINE_string("checkpoint_dir", "checkpoint", "Directory to load checkpoint from")
flags.DEFINE_integer("num_samples", 200, "Number of sample gifs to generate [1000]")
flags.DEFINE_string("output_directory", "samples_nested", "Directory to write output gifs")
flags.DEFINE_integer("random_seed", 0, "Random numpy seed to use [0]")
flags.DEFINE_boolean("continuous", False, "Enable infinite video generation")
flags.DEFINE_string("precision", "bf16", "bf16 or fp32")


def write_gif(video, filename, fps=25):
    """model_sampler.py:26-28 (imageio.mimsave): uint8 frames -> animated GIF (PIL is the writer available here)."""
    video = ((video + 1) / 2 * 255).astype(np.uint8)
    from PIL import Image
    frames = [Image.fromarray(f if f.shape[-1] == 3 else f[..., 0]) for f in video]
    frames[0].save(filename, save_all=True, append_images=frames[1:], duration=int(1000 / fps), loop=0)


def main(argv=None):
    FLAGS = flags.parse(argv)
    utils.pp.pprint(vars(FLAGS))
    np.random.seed(FLAGS.random_seed)
    os.makedirs(FLAGS.output_directory, exist_ok=True)
    ops.set_precision(FLAGS.precision)
    ops.reset_default_store()
    with ops.variable_scope('video_gan'):
        vid_dcgan = VID_DCGAN(None, FLAGS.vid_batch_size, 120, 100, FLAGS.vid_length, FLAGS.image_size, FLAGS.output_size, c_dim=FLAGS.c_dim,
                              image_noise_std=FLAGS.image_noise, activation_noise_std=FLAGS.activation_noise)
        vid_dcgan.load_checkpoint(None, FLAGS.checkpoint_dir)
        cntr = 0
        while True:
            for i in range(0, FLAGS.num_samples, vid_dcgan.batch_size):
                sample_z = np.random.uniform(-1, 1, size=(vid_dcgan.batch_size, 120)).astype(np.float32)
                samples = vid_dcgan.sample(torch.as_tensor(sample_z).to(vid_dcgan.store.device), is_training=False).float().cpu().numpy()
                videos = np.reshape(samples, [vid_dcgan.batch_size, vid_dcgan.vid_length, vid_dcgan.output_image_size,
                                              vid_dcgan.output_image_size, vid_dcgan.c_dim])
                for j in range(i, min(FLAGS.num_samples, i + vid_dcgan.batch_size)):
                    tmp = os.path.join(FLAGS.output_directory, "tmp.gif")
                    write_gif(videos[j - i], tmp)
                    os.rename(tmp, os.path.join(FLAGS.output_directory, "%d.gif" % j))
            if not FLAGS.continuous:
                break
            print("***Finished iteration %d ***" % cntr)
            cntr += 1
    return vid_dcgan


if __name__ == '__main__':
    main()
