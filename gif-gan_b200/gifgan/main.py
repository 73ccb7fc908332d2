"""CLI for the image DCGAN -- drop-in for /root/reference/models/recurrent_z/main.py (same flags and defaults,
main.py:10-29).  `--dataset synthetic` trains on seeded random frames (no dataset ships with the repo);
extra flags: --precision {bf16,fp32}."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gifgan import ops  # noqa: E402
from gifgan.flags import Flags  # noqa: E402
from gifgan.model import DCGAN  # noqa: E402
from gifgan.utils import pp  # noqa: E402

flags = Flags()
flags.DEFINE_integer("epoch", 25, "Epoch to train [25]")
flags.DEFINE_float("learning_rate", 0.0002, "Learning rate of for adam [0.0002]")
flags.DEFINE_float("beta1", 0.5, "Momentum term of adam [0.5]")
flags.DEFINE_integer("train_size", np.inf if False else 1 << 62, "The size of train images [np.inf]")
flags.DEFINE_integer("batch_size", 64, "The size of batch images [64]")
flags.DEFINE_integer("image_size", 108, "The size of image to use (will be center cropped) [108]")
flags.DEFINE_integer("output_size", 64, "The size of the output images to produce [64]")
flags.DEFINE_integer("c_dim", 3, "Dimension of image color. [3]")
flags.DEFINE_string("dataset", "celebA", "The name of dataset [celebA, mnist, lsun, synthetic]")
flags.DEFINE_string("checkpoint_dir", "checkpoint", "Directory name to save the checkpoints [checkpoint]")
flags.DEFINE_string("sample_dir", "samples", "Directory name to save the image samples [samples]")
flags.DEFINE_string("data_dir", "./data", "Directory to read dataset from")
flags.DEFINE_string("log_dir", "./logs", "Directory to write log files")
flags.DEFINE_string("image_glob", "*.jpg", "Glob to use to find images in the dataset directory")
flags.DEFINE_boolean("is_train", False, "True for training, False for testing [False]")
flags.DEFINE_boolean("is_crop", False, "True for training, False for testing [False]")
flags.DEFINE_boolean("visualize", False, "True for visualizing, False for nothing [False]")
flags.DEFINE_boolean("shuffle", False, "True to shuffle the dataset, False otherwise [False]")
flags.DEFINE_string("precision", "bf16", "bf16 (tensor cores) or fp32 (parity mode)")


def main(argv=None):
    FLAGS = flags.parse(argv)
    pp.pprint(vars(FLAGS))
    os.makedirs(FLAGS.checkpoint_dir, exist_ok=True)
    os.makedirs(FLAGS.sample_dir, exist_ok=True)
    ops.set_precision(FLAGS.precision)
    ops.reset_default_store()
    if FLAGS.dataset == 'mnist':
        dcgan = DCGAN(None, image_size=FLAGS.image_size, batch_size=FLAGS.batch_size, y_dim=10, output_size=28, c_dim=1,
                      dataset_name=FLAGS.dataset, is_crop=FLAGS.is_crop, checkpoint_dir=FLAGS.checkpoint_dir, sample_dir=FLAGS.sample_dir,
                      data_dir=FLAGS.data_dir, log_dir=FLAGS.log_dir, image_glob=FLAGS.image_glob, shuffle=FLAGS.shuffle)
    else:
        dcgan = DCGAN(None, image_size=FLAGS.image_size, batch_size=FLAGS.batch_size, output_size=FLAGS.output_size, c_dim=FLAGS.c_dim,
                      dataset_name=FLAGS.dataset, is_crop=FLAGS.is_crop, checkpoint_dir=FLAGS.checkpoint_dir, sample_dir=FLAGS.sample_dir,
                      data_dir=FLAGS.data_dir, log_dir=FLAGS.log_dir, image_glob=FLAGS.image_glob, shuffle=FLAGS.shuffle)
    if FLAGS.is_train:
        dcgan.train(FLAGS)
    else:
        dcgan.load(FLAGS.checkpoint_dir)
    return dcgan


if __name__ == '__main__':
    main()
