"""Video inversion entry point (the role of models/recurrent_z/z_space_finder.py; options: flags.TABLES["z_space_finder"]):
for every listed clip, search the image GAN's latent space for the z of each frame -- frame 0 from random latents, each
later frame warm-started from the previous one -- and write the [vid_length, 100] latents as <clip>.npy (the training
data of the video generator).  Optional outputs: a strip of the matched frames, single frames, a side-by-side clip."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gifgan import flags, utils  # noqa: E402
from gifgan.latent_search import load_dcgan, read_video_frames, search_from_options  # noqa: E402


def out_path(fname, base, ext=".npy"):
    return os.path.join(base, os.path.splitext(os.path.basename(fname))[0] + ext)


def write_outputs(opts, names, targets, results, latents):
    """z_space_finder.py:164-201, per clip of the batch (padding clips have no name and are dropped by zip)."""
    for name, target, result, z in zip(names, targets, results, latents):
        if opts.output_image_folder:
            utils.save_images(result, [1, opts.vid_length], out_path(name, opts.output_image_folder, ".png"))
        if opts.output_frame_folder:
            folder = out_path(name, opts.output_frame_folder, "")
            os.makedirs(folder, exist_ok=True)
            for i in range(opts.vid_length):
                utils.save_images(result[i:i + 1], [1, 1], os.path.join(folder, "%03d.png" % i))
        if opts.output_comparison_folder:
            import cv2
            size = (2 * opts.image_size, 2 * opts.image_size)                  # each half is scaled up 2x
            wr = cv2.VideoWriter(out_path(name, opts.output_comparison_folder, ".mp4"), 0x20, 25.0, (2 * size[0], size[1]))
            for t in range(opts.vid_length):
                halves = [cv2.resize(cv2.cvtColor(np.around(utils.inverse_transform(im) * 255).astype('uint8'), cv2.COLOR_RGB2BGR), size,
                                     interpolation=cv2.INTER_LINEAR) for im in (target[t], result[t])]
                wr.write(np.concatenate(halves, axis=1))
            wr.release()
        np.save(out_path(name, opts.output_z_folder), z)


def clip_batches(opts):
    """Yields (names, clips) with up to video_batch_size clips each; clips already inverted are skipped (:306-313)."""
    if opts.synthetic:
        rs = np.random.RandomState(107)
        for first in range(0, opts.synthetic, opts.video_batch_size):
            n = min(opts.video_batch_size, opts.synthetic - first)
            yield (["synthetic_%04d" % (first + i) for i in range(n)],
                   [list(rs.uniform(-1, 1, (opts.vid_length, opts.image_size, opts.image_size, opts.c_dim))) for _ in range(n)])
        return
    files = [line.strip() for lst in opts.video_list for line in open(lst) if line.strip()]
    print("Total video files found:", len(files))
    names, clips = [], []
    for fname in files:
        if os.path.exists(out_path(fname, opts.output_z_folder)):
            print("Skipping %s because already processed" % fname)
            continue
        vid = read_video_frames(os.path.join(opts.video_dataset_dir, fname), opts.image_size, opts.vid_length, opts.frame_skip)
        if vid:
            names.append(fname); clips.append(vid)
        if len(clips) == opts.video_batch_size:
            yield names, clips
            names, clips = [], []
    if clips:
        yield names, clips


def main(argv=None):
    opts = flags.parse("z_space_finder", argv)
    if not opts.output_z_folder or not (opts.video_list or opts.synthetic):
        raise SystemExit("--output_z_folder and --video_list (or --synthetic n) are required")
    for d in (opts.output_z_folder, opts.output_comparison_folder, opts.output_image_folder, opts.output_frame_folder):
        if d:
            os.makedirs(d, exist_ok=True)
    search = search_from_options(load_dcgan(opts, opts.video_batch_size), opts)
    done = 0
    for names, clips in clip_batches(opts):
        if opts.stop_after > 0 and done >= opts.stop_after:
            break
        done += 1
        empty = [np.zeros((opts.image_size, opts.image_size, opts.c_dim))] * opts.vid_length          # pad_batch, :106-110
        targets = np.array(clips + [empty] * (opts.video_batch_size - len(clips)), dtype=np.float32)
        results, latents = search.fit_video(targets, opts.num_initial_steps, opts.num_steps_per_frame, opts.learning_rate,
                                            opts.lr_decay_amount, log=print)
        print("Writing output for this batch ...")
        write_outputs(opts, names, targets, results, latents)
        print("Done output for this batch")
    return search


if __name__ == "__main__":
    main()
