"""Image inversion entry point (the role of models/recurrent_z/discriminator_activation_optimizer.py; options:
flags.TABLES["activation_optimizer"]): search the latents of a num_rows x num_cols grid of target images (given images,
or the first frames of given clips, repeated to fill the grid) and write target.png, train_<i>.png and final.png.
With --vid_length N it is discriminator_activation_optimizer_video.py: all N frames of every input clip are searched at
once (batch = clips x N, no warm start between frames; the grid has one row per clip).
With --nested it is discriminator_activation_optimizer_nested.py: the variable is the VIDEO latent [batch, 120] of a VID_DCGAN
checkpoint, searched through the video generator so that the first frame of every generated clip matches its target
(latent_search.NestedLatentSearch); the clips are written as final.mp4.
With --vid_length N --iterative it is discriminator_activation_optimizer_video_iterative.py: batch = clips, frame 0 is searched for
--num_initial_steps, the learning rate is multiplied by --lr_decay_amount once, then every frame is tracked in turn for
--num_steps_per_frame steps from the previous frame's latents (LatentSearch.fit_video: the schedule z_space_finder.py shares);
outputs final.png (one row per clip), final_frames/, final_z.npy [clips, N, z_dim] and tween_frames/ (latents interpolated
between consecutive frames, `tween_frames` per gap -- reference lines 243-262).  The reference's defaults for that program are
--learning_rate 0.05 --lr_decay_amount 0.5.
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


def tween_latents(zs, tween_frames):
    """discriminator_activation_optimizer_video_iterative.py:243-262: the frame sequence with `tween_frames` latents linearly
    interpolated between consecutive tracked frames.  zs [B, T, z_dim] -> list of (index, latents [B, z_dim], is_tracked_frame)."""
    B, T, _ = zs.shape
    out = []
    for i in range(T):
        cnt = i * (tween_frames + 1)
        out.append((cnt, zs[:, i], True))
        if i + 1 < T:
            for j in range(1, tween_frames + 1):
                delta = j / float(tween_frames + 1)
                out.append((cnt + j, zs[:, i + 1] * delta + zs[:, i] * (1 - delta), False))
    return out


def run_iterative(opts):
    """discriminator_activation_optimizer_video_iterative.py:86-262."""
    targets = load_clip_targets(opts)
    T = opts.vid_length
    clips = len(targets) // T
    s, c = opts.image_size, opts.c_dim
    search = search_from_options(load_dcgan(opts, clips), opts)
    utils.save_images(targets, [clips, T], os.path.join(opts.sample_dir, "target.png"))
    results, zs = search.fit_video(targets.reshape(clips, T, s, s, c), opts.num_initial_steps, opts.num_steps_per_frame, opts.learning_rate,
                                   opts.lr_decay_amount, log=print)
    utils.save_images(results.reshape(-1, s, s, c), [clips, T], os.path.join(opts.sample_dir, "final.png"))
    print("Saved final images")
    np.save(os.path.join(opts.sample_dir, "final_z.npy"), zs)
    ff, tf_ = os.path.join(opts.sample_dir, "final_frames"), os.path.join(opts.sample_dir, "tween_frames")
    os.makedirs(ff, exist_ok=True)
    os.makedirs(tf_, exist_ok=True)
    for i in range(T):
        utils.save_images(results[:, i], [1, clips], os.path.join(ff, "final_frame_%03d.png" % i))
    print("Generating tween results")
    for cnt, z, tracked in tween_latents(zs, opts.tween_frames):
        if tracked:
            img = results[:, cnt // (opts.tween_frames + 1)]
        else:
            search.assign(z)
            img = search.images().float().cpu().numpy()
        utils.save_images(img, [1, clips], os.path.join(tf_, "tween_frame_%03d.png" % cnt))
    return search, results, zs


def run_nested(opts):
    """discriminator_activation_optimizer_nested.py:128-324: a num_rows x num_cols grid of VIDEO latents is searched so that the first
    frame of every generated clip matches its target; outputs target.png, train_<i>.png / final.png (first frames), final.mp4 (the
    clips as a grid, 25 fps), final_z.npy [batch, 120]."""
    import torch
    from gifgan import ops
    from gifgan.latent_search import NestedLatentSearch, WEIGHT_NAMES, normalised_weights
    from gifgan.z_model_lib import VID_DCGAN
    if opts.discriminator_mode not in ("train", "inference"):
        raise SystemExit("--discriminator_mode must be train or inference")
    batch, T = opts.num_rows * opts.num_cols, (opts.vid_length or 16)
    ops.set_precision(opts.precision)
    ops.reset_default_store()
    with ops.variable_scope('video_gan'):                       # reference line 130
        vid = VID_DCGAN(None, batch_size=batch, z_input_size=120, z_output_size=100, vid_length=T, input_image_size=opts.image_size,
                        output_image_size=opts.output_size, c_dim=opts.c_dim, sample_cols=opts.num_cols)
    if opts.checkpoint_directory:
        if not vid.load_checkpoint(None, opts.checkpoint_directory):
            raise SystemExit("no VID_DCGAN checkpoint in %s" % opts.checkpoint_directory)
    elif not opts.synthetic:
        raise SystemExit("--checkpoint_directory is required (or --synthetic n to run on random weights and targets)")
    w = normalised_weights(opts)
    print("Normalized loss weights:")
    for k in WEIGHT_NAMES:
        print(k, w[k])
    search = NestedLatentSearch(vid, discriminator_mode=opts.discriminator_mode, beta1=opts.beta1, random_seed=opts.random_seed,
                                use_graph=opts.cuda_graph, **w)
    targets = load_targets(opts, batch)
    grid = [opts.num_rows, opts.num_cols]
    utils.save_images(targets, grid, os.path.join(opts.sample_dir, "target.png"))

    def on_step(i, loss, srch):
        if opts.sample_frequency > 0 and i % opts.sample_frequency == 0:
            utils.save_images(srch.images().float().cpu().numpy()[::T], grid, os.path.join(opts.sample_dir, "train_%d.png" % i))
            print("Saved sample")
        print("Step %d/%d: loss %f" % (i, opts.num_steps, loss))

    frames = search.optimise(targets, opts.num_steps, opts.learning_rate, opts.lr_decay_frequency, opts.lr_decay_amount, on_step)
    utils.save_images(frames[::T], grid, os.path.join(opts.sample_dir, "final.png"))
    print("Saved final images")
    np.save(os.path.join(opts.sample_dir, "final_z.npy"), search.z.detach().cpu().numpy())
    write_clip_grid(frames.reshape(opts.num_rows, opts.num_cols, T, opts.output_size, opts.output_size, opts.c_dim),
                    os.path.join(opts.sample_dir, "final.mp4"))
    return search, frames


def write_clip_grid(videos, filename, fps=25.0):
    """Reference lines 305-324: [rows, cols, T, s, s, c] in (-1, 1) -> an mp4 whose frame t is the rows x cols grid of frame t."""
    import cv2
    rows, cols, T, sz, _, c = videos.shape
    print("Writing samples to", filename)
    wr = utils.open_video_writer(filename, fps, (cols * sz, rows * sz))
    for t in range(T):
        frame = np.zeros((rows * sz, cols * sz, c), dtype=np.uint8)
        for r in range(rows):
            for k in range(cols):
                im = np.around(utils.inverse_transform(videos[r, k, t]) * 255).astype('uint8')
                frame[r * sz:(r + 1) * sz, k * sz:(k + 1) * sz, :] = cv2.cvtColor(im, cv2.COLOR_RGB2BGR) if c == 3 else im
        wr.write(frame)
    wr.release()


def main(argv=None):
    opts = flags.parse("activation_optimizer", argv)
    if not opts.sample_dir:
        raise SystemExit("--sample_dir is required")
    os.makedirs(opts.sample_dir, exist_ok=True)
    if opts.nested:
        return run_nested(opts)[0]
    if opts.iterative:
        if opts.vid_length <= 0:
            raise SystemExit("--iterative needs --vid_length N")
        return run_iterative(opts)[0]
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
