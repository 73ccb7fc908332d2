"""CLI for the video GAN -- drop-in for /root/reference/models/recurrent_z/z_model.py (same flags and defaults,
z_model.py:22-56).  `--video_list synthetic` (or no list) trains on seeded random clips."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gifgan import ops  # noqa: E402
from gifgan.flags import Flags  # noqa: E402
from gifgan.utils import pp  # noqa: E402
from gifgan.z_model_lib import VID_DCGAN  # noqa: E402

flags = Flags()
flags.DEFINE_integer("epoch", 25, "Epoch to train [25]")
flags.DEFINE_float("learning_rate", 0.0002, "Learning rate of for adam [0.0002]")
flags.DEFINE_float("beta1", 0.5, "Momentum term of adam [0.5]")
flags.DEFINE_integer("image_batch_size", 64, "The size of batch images [64]")
flags.DEFINE_integer("vid_batch_size", 64, "The size of batch images [64]")
flags.DEFINE_integer("vid_length", 16, "The length of the videos [16]")
flags.DEFINE_integer("image_size", 64, "The size of images used [64]")
flags.DEFINE_integer("output_size", 64, "The size of the output images to produce [64]")
flags.DEFINE_integer("c_dim", 3, "Dimension of image color. [3]")
flags.DEFINE_string("image_model_dir", "checkpoint", "Directory name to load the image checkpoints [checkpoint]")
flags.DEFINE_string("video_checkpoint_dir", "checkpoint", "Directory name to save the video checkpoints [checkpoint]")
flags.DEFINE_string("video_sample_dir", "samples", "Directory name to save the video samples [samples]")
flags.DEFINE_string("video_data_dir", "./data", "Directory to read dataset from")
flags.DEFINE_string("video_dataset", "", "Name of video dataset to use")
flags.add_argument("--video_list", required=False, nargs='*', default=[], help="List(s) of videos to use")
flags.DEFINE_string("log_dir", "./logs", "Directory to write log files")
flags.DEFINE_boolean("is_train", False, "True for training, False for <not implemented yet> [False]")
flags.DEFINE_boolean("video_shuffle", True, "True to shuffle the dataset, False otherwise [False]")
flags.DEFINE_boolean("train_img_gen", False, "True to make the image generator params trainable [False]")
flags.DEFINE_boolean("train_img_disc", False, "True to make the image discriminator params trainable [False]")
flags.DEFINE_integer("disc_updates", 1, "Number of discriminator updates per batch [1]")
flags.DEFINE_integer("gen_updates", 2, "Number of generator updates per batch [1]")
flags.DEFINE_float("image_noise", 0.0, "Std of noise to add to images")
flags.DEFINE_float("activation_noise", 0.0, "Std of noise to add to D activations")
flags.DEFINE_float("first_frame_loss_scalar", 0.0, "first_frame_loss_scalar")
flags.DEFINE_integer("sample_frequency", 10, "How often to save checkpoints & samples")
flags.DEFINE_integer("max_checkpoints_to_keep", 5, "Max number of checkpoints to keep")
flags.DEFINE_string("precision", "bf16", "bf16 (tensor cores) or fp32 (parity mode)")
flags.DEFINE_integer("synthetic_batches", 4, "batches per epoch when training on synthetic clips")


def main(argv=None):
    FLAGS = flags.parse(argv)
    pp.pprint(vars(FLAGS))
    ops.set_precision(FLAGS.precision)
    ops.reset_default_store()
    with ops.variable_scope('video_gan'):
        vid_z_dim = 120
        image_z_dim = 100
        vid_dcgan = VID_DCGAN(None, FLAGS.vid_batch_size, vid_z_dim, image_z_dim, FLAGS.vid_length, FLAGS.image_size, FLAGS.output_size,
                              c_dim=FLAGS.c_dim, image_noise_std=FLAGS.image_noise, activation_noise_std=FLAGS.activation_noise,
                              first_frame_loss_scalar=FLAGS.first_frame_loss_scalar)
        print("DONE")
        # Load image model weights (after init, so that they are not overwritten): z_model.py:86
        vid_dcgan.load_image_gan(None, FLAGS.image_model_dir)
        # smoke prints of z_model.py:88-101
        dev = vid_dcgan.store.device
        sample_z = torch.as_tensor(np.random.uniform(-1, 1, size=(FLAGS.vid_batch_size, vid_z_dim)).astype(np.float32)).to(dev)
        with torch.no_grad():
            out_val = vid_dcgan.generator(sample_z)[0]
            print(tuple(out_val.shape))
            imgs = vid_dcgan.img_dcgan.sampler(out_val)
            print(tuple(imgs.shape), (FLAGS.vid_batch_size, FLAGS.vid_length, FLAGS.output_size, FLAGS.output_size, FLAGS.c_dim))
        if FLAGS.is_train:
            vid_dcgan.train(None, FLAGS)
    return vid_dcgan


if __name__ == '__main__':
    main()
