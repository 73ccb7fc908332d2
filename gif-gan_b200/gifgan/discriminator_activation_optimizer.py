"""Image inversion entry point (the role of models/recurrent_z/discriminator_activation_optimizer.py; options:
flags.TABLES["activation_optimizer"]): search the latents of a num_rows x num_cols grid of target images (given images,
or the first frames of given clips, repeated to fill the grid) and write target.png, train_<i>.png and final.png.
With --vid_length N it is discriminator_activation_optimizer_video.py: all N frames of every input clip are searched at
once (batch = clips x N, no warm start between frames; the grid has one row per clip).
The interactive GUI, the progress video and the latent-path playback of the reference are display features outside
the compute path and are not carried over."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gifgan import flags, utils  # noqa: E402
from gifgan.latent_search import load_dcgan, read_video_frames, search_from_options  # noqa: E402


def load_clip_targets(opts):
    """discriminator_activation_optimizer_video.py:64-103: [clips * vid_length, s, s, c], clip-major."""
    if opts.synthetic:
        return np.random.RandomState(109).uniform(-1, 1, (opts.synthetic * opts.vid_length, opts.image_size, opts.image_size, opts.c_dim)).astype(np.float32)
    clips = []
    for v in opts.input_videos:
        frames = read_video_frames(v, opts.image_size, opts.vid_length, opts.frame_skip)
        if frames is None:
            raise SystemExit("Video %s not long enough!" % v)
        clips.append(frames)
    if not clips:
        raise SystemExit("no targets: give --input_videos or --synthetic n")
    return np.array(clips, dtype=np.float32).reshape(-1, opts.image_size, opts.image_size, opts.c_dim)


def load_targets(opts, batch):
    if opts.synthetic:
        return np.random.RandomState(108).uniform(-1, 1, (batch, opts.image_size, opts.image_size, opts.c_dim)).astype(np.float32)
    targets = [utils.get_image(p, opts.image_size, is_crop=False, resize_w=opts.image_size) for p in opts.input_images]
    for v in opts.input_videos:
        frames = read_video_frames(v, opts.image_size, 1, 1)
        if frames:
            targets.append(frames[0])
    if not targets:
        raise SystemExit("no targets: give --input_images, --input_videos or --synthetic n")
    import cv2
    targets = [t if t.shape[0] == opts.image_size else cv2.resize(t, (opts.image_size, opts.image_size)) for t in targets]
    return np.array([targets[i % len(targets)] for i in range(batch)], dtype=np.float32)


def main(argv=None):
    opts = flags.parse("activation_optimizer", argv)
    if not opts.sample_dir:
        raise SystemExit("--sample_dir is required")
    os.makedirs(opts.sample_dir, exist_ok=True)
    if opts.vid_length > 0:
        targets = load_clip_targets(opts)
        batch = len(targets)
        grid = [batch // opts.vid_length, opts.vid_length]
        search = search_from_options(load_dcgan(opts, batch), opts)
    else:
        batch = opts.num_rows * opts.num_cols
        search = search_from_options(load_dcgan(opts, batch), opts)
        targets = load_targets(opts, batch)
        grid = [opts.num_rows, opts.num_cols]
    utils.save_images(targets, grid, os.path.join(opts.sample_dir, "target.png"))

    def on_step(i, loss, srch):
        if opts.sample_frequency > 0 and i % opts.sample_frequency == 0:
            utils.save_images(srch.images().float().cpu().numpy(), grid, os.path.join(opts.sample_dir, "train_%d.png" % i))
            print("Saved sample")
        print("Step %d/%d: loss %f" % (i, opts.num_steps, loss))

    final = search.optimise(targets, opts.num_steps, opts.learning_rate, opts.lr_decay_frequency, opts.lr_decay_amount, on_step)
    utils.save_images(final, grid, os.path.join(opts.sample_dir, "final.png"))
    print("Saved final images")
    np.save(os.path.join(opts.sample_dir, "final_z.npy"), search.z.detach().cpu().numpy())
    return search


if __name__ == "__main__":
    main()
