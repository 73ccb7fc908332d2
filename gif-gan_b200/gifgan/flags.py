"""Command-line flags of the entry points, as data.

The reference declares its options with tf.app.flags (main.py:10-29, z_model.py:22-56, model_sampler.py:9-24).  The
drop-in contract is the option NAMES, TYPES and DEFAULTS; here each program's options are one table
(name, kind, default, help) turned into an argparse parser, plus the options this implementation adds
(--precision, synthetic data).  Kinds: i = int, f = float, s = str, b = bool ("--flag", "--flag true/false").
"""
import argparse

_TRUE = ("1", "true", "t", "yes", "y")

# Options shared by every program that builds the video GAN.
_VIDEO_SHAPE = [
    ("vid_batch_size", "i", 64, "clips per batch"),
    ("vid_length", "i", 16, "frames per clip"),
    ("image_size", "i", 64, "edge of the frames fed to the image discriminator"),
    ("output_size", "i", 64, "edge of the generated frames"),
    ("c_dim", "i", 3, "colour channels"),
    ("image_noise", "f", 0.0, "std of the gaussian noise added to frames before the image D"),
    ("activation_noise", "f", 0.0, "std of the gaussian noise added to the image-D activations"),
]

# Options shared by the latent-search programs (z_space_finder.py:31-41, discriminator_activation_optimizer.py:33-45).
_LATENT_WEIGHTS = [
    ("pixel_L2_weight", "f", 0.0, "weight of the L2 distance between generated and target pixels"),
    ("pixel_L1_weight", "f", 0.0, "weight of the L1 distance between generated and target pixels"),
    ("activations_L2_weight", "f", 1.0, "weight of the L2 distance between discriminator h2 activations"),
    ("activations_L1_weight", "f", 0.0, "weight of the L1 distance between discriminator h2 activations"),
    ("generator_loss_weight", "f", 0.0, "weight of the generator's adversarial loss"),
]
_LATENT_DCGAN = [
    ("checkpoint_directory", "s", "", "image-GAN checkpoint to load"),
    ("image_size", "i", 64, "edge of the target frames"),
    ("output_size", "i", 64, "edge of the generated frames"),
    ("c_dim", "i", 3, "colour channels"),
    ("synthetic", "i", 0, "n > 0: search n seeded random targets instead of reading files (no dataset ships here)"),
    ("cuda_graph", "b", True, "replay one captured CUDA graph per search step instead of eager launches"),
]

_OURS = [("precision", "s", "bf16", "bf16 = tcgen05 tensor-core path, fp32 = parity mode")]

TABLES = {
    # models/recurrent_z/main.py
    "image_gan": [
        ("epoch", "i", 25, "passes over the dataset"),
        ("learning_rate", "f", 0.0002, "Adam step size"),
        ("beta1", "f", 0.5, "Adam first-moment decay"),
        ("train_size", "i", 1 << 62, "cap on the number of training images (reference default: unbounded)"),
        ("batch_size", "i", 64, "images per batch"),
        ("image_size", "i", 108, "centre-crop edge applied to the input images"),
        ("output_size", "i", 64, "edge of the generated images"),
        ("c_dim", "i", 3, "colour channels"),
        ("dataset", "s", "celebA", "celebA | mnist | lsun | synthetic (seeded random frames; no dataset ships here)"),
        ("checkpoint_dir", "s", "checkpoint", "where checkpoints are written / read"),
        ("sample_dir", "s", "samples", "where sample grids are written"),
        ("data_dir", "s", "./data", "dataset root"),
        ("log_dir", "s", "./logs", "log directory"),
        ("image_glob", "s", "*.jpg", "pattern of the image files under data_dir/dataset"),
        ("is_train", "b", False, "train (true) or just load the checkpoint (false)"),
        ("is_crop", "b", False, "centre-crop the inputs"),
        ("visualize", "b", False, "write visualisations after loading"),
        ("shuffle", "b", False, "shuffle the file list every epoch"),
    ] + _OURS,
    # models/recurrent_z/z_model.py
    "video_gan": [
        ("epoch", "i", 25, "passes over the clip list"),
        ("learning_rate", "f", 0.0002, "Adam step size"),
        ("beta1", "f", 0.5, "Adam first-moment decay"),
        ("image_batch_size", "i", 64, "batch of the image-GAN checkpoint being loaded"),
    ] + _VIDEO_SHAPE[:5] + [
        ("image_model_dir", "s", "checkpoint", "image-GAN checkpoint to start from"),
        ("video_checkpoint_dir", "s", "checkpoint", "where video-GAN checkpoints go"),
        ("video_sample_dir", "s", "samples", "where sample clips go"),
        ("video_data_dir", "s", "./data", "clip dataset root"),
        ("video_dataset", "s", "", "name of the clip dataset"),
        ("log_dir", "s", "./logs", "log directory"),
        ("is_train", "b", False, "run the training loop"),
        ("video_shuffle", "b", True, "shuffle the clip list every epoch"),
        ("train_img_gen", "b", False, "also train the image generator (default: frozen)"),
        ("train_img_disc", "b", False, "also train the image discriminator (default: frozen)"),
        ("disc_updates", "i", 1, "discriminator updates per batch"),
        ("gen_updates", "i", 2, "generator updates per batch"),
    ] + _VIDEO_SHAPE[5:] + [
        ("first_frame_loss_scalar", "f", 0.0, "weight of the first-frame latent reconstruction loss"),
        ("sample_frequency", "i", 10, "batches between checkpoints / samples"),
        ("max_checkpoints_to_keep", "i", 5, "checkpoint rotation depth"),
        ("synthetic_batches", "i", 4, "batches per epoch when training on synthetic clips"),
    ] + _OURS,
    # models/recurrent_z/model_sampler.py  (its two path defaults are site-specific in the reference; local ones here)
    "sampler": _VIDEO_SHAPE + [
        ("checkpoint_dir", "s", "checkpoint", "video-GAN checkpoint to sample from"),
        ("num_samples", "i", 200, "GIFs to write per pass"),
        ("output_directory", "s", "samples_nested", "where the GIFs go"),
        ("random_seed", "i", 0, "numpy seed of the latents"),
        ("continuous", "b", False, "keep regenerating until interrupted"),
    ] + _OURS,
    # models/recurrent_z/z_space_finder.py:11-41 (its required options default to "" here and are checked by the program)
    "z_space_finder": [
        ("video_dataset_dir", "s", "", "directory the listed clips are read from"),
        ("output_z_folder", "s", "", "where the per-clip latents (.npy, [vid_length, 100]) are written"),
        ("output_comparison_folder", "s", "", "optional: side-by-side target | result clips"),
        ("output_image_folder", "s", "", "optional: one strip of the final frames per clip"),
        ("output_frame_folder", "s", "", "optional: the final frames as single images"),
        ("video_batch_size", "i", 8, "clips searched at once"),
        ("stop_after", "i", 0, "debug: stop after n batches"),
        ("random_seed", "i", 0, "seed of the initial latents"),
        ("num_initial_steps", "i", 500, "steps spent on the first frame before tracking"),
        ("num_steps_per_frame", "i", 100, "steps per tracked frame"),
        ("learning_rate", "f", 0.05, "Adam step size of the search"),
        ("lr_decay_amount", "f", 0.5, "factor applied once, after the initial steps"),
        ("beta1", "f", 0.5, "Adam first-moment decay"),
        ("discriminator_mode", "s", "", "train | inference: batch statistics or moving averages in the batch norms"),
        ("vid_length", "i", 16, "frames used per clip"),
        ("frame_skip", "i", 2, "take every n-th frame of the file"),
    ] + _LATENT_WEIGHTS + _LATENT_DCGAN + _OURS,
    # models/recurrent_z/discriminator_activation_optimizer.py:16-55 (GUI, progress-video and path options are not carried over)
    "activation_optimizer": [
        ("random_seed", "i", 0, "seed of the initial latents"),
        ("num_rows", "i", 8, "grid rows (batch = rows x cols)"),
        ("num_cols", "i", 8, "grid columns"),
        ("num_steps", "i", 1000, "search steps"),
        ("learning_rate", "f", 0.0002, "Adam step size of the search"),
        ("beta1", "f", 0.5, "Adam first-moment decay"),
        ("discriminator_mode", "s", "", "train | inference"),
        ("sample_dir", "s", "", "where target.png, train_<i>.png and final.png go"),
        ("sample_frequency", "i", 100, "steps between sample grids (0: none)"),
        ("lr_decay_frequency", "i", 0, "steps between learning-rate decays (0: never)"),
        ("lr_decay_amount", "f", 0.9, "decay factor"),
        ("vid_length", "i", 0, "> 0: discriminator_activation_optimizer_video.py -- search every `frame_skip`-th frame of the input clips (batch = clips x vid_length, one grid row per clip)"),
        ("frame_skip", "i", 2, "frame step when vid_length > 0"),
        ("nested", "b", False, "discriminator_activation_optimizer_nested.py: search the VIDEO latent [batch, 120] of a VID_DCGAN checkpoint so "
                               "that the first frame of every generated clip matches its target; writes the clips (vid_length default 16)"),
        ("iterative", "b", False, "with vid_length > 0: discriminator_activation_optimizer_video_iterative.py -- batch = clips; frame 0 for "
                                  "num_initial_steps, then every frame in turn for num_steps_per_frame steps, warm-started from the previous frame"),
        ("num_initial_steps", "i", 500, "iterative: steps on frame 0 before tracking"),
        ("num_steps_per_frame", "i", 100, "iterative: steps per tracked frame"),
        ("tween_frames", "i", 2, "iterative: frames interpolated in latent space between consecutive tracked frames (tween_frames/ output)"),
    ] + _LATENT_WEIGHTS + _LATENT_DCGAN + _OURS,
}


def _add(parser, name, kind, default, text):
    if kind == "b":
        parser.add_argument("--" + name, type=lambda v: str(v).lower() in _TRUE, nargs="?", const=True, default=default, help=text)
    else:
        conv = {"i": lambda s: int(float(s)), "f": float, "s": str}[kind]
        parser.add_argument("--" + name, type=conv, default=default, help=text)


def parser_for(program):
    p = argparse.ArgumentParser(prog=program)
    for spec in TABLES[program]:
        _add(p, *spec)
    if program == "video_gan":
        p.add_argument("--video_list", nargs="*", default=[], help="file(s) listing the training clips; empty or 'synthetic' = seeded random clips")
    if program == "z_space_finder":
        p.add_argument("--video_list", nargs="*", default=[], help="file(s) listing the clips to invert, one name per line")
    if program == "activation_optimizer":
        p.add_argument("--input_videos", nargs="*", default=[], help="search for the first frame of these clips")
        p.add_argument("--input_images", nargs="*", default=[], help="search for these images")
    return p


def parse(program, argv=None):
    """-> argparse.Namespace named like the reference's FLAGS."""
    return parser_for(program).parse_args(argv)


class Flags:
    """tf.app.flags-style incremental definition, kept for user scripts that extend a table."""

    def __init__(self, program=None):
        self._p = parser_for(program) if program else argparse.ArgumentParser()

    def DEFINE_integer(self, name, default, help=""):
        _add(self._p, name, "i", default, help)

    def DEFINE_float(self, name, default, help=""):
        _add(self._p, name, "f", default, help)

    def DEFINE_string(self, name, default, help=""):
        _add(self._p, name, "s", default, help)

    def DEFINE_boolean(self, name, default, help=""):
        _add(self._p, name, "b", default, help)

    def parse(self, argv=None):
        return self._p.parse_args(argv)
