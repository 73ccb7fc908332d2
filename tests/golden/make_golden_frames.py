#!/usr/bin/env python
"""Generates tests/golden/frames_resize.npz with the REAL dependency of the reference's frame decode -- OpenCV
(`cv2.resize(..., interpolation=cv2.INTER_LINEAR)`, /root/reference/models/recurrent_z/z_model_lib.py:339-346) -- plus the
reference's own normalisation formula (utils.py:57-63, numpy float64).  Run here (opencv-python 4.13.0):
    python tests/golden/make_golden_frames.py
The fixture pins oracle/image_ops.py (CPU test) and the CUDA kernel gg_frames_to_input (GPU test) to cv2's bytes."""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# (source H, W) -> destination size: reductions (the data set's clips), the exact 2x case, a mild enlargement, odd sizes
CASES = [((48, 64), 32), ((64, 64), 32), ((24, 20), 32), ((37, 65), 32), ((32, 32), 32), ((60, 80), 16), ((33, 17), 48)]


def main():
    rs = np.random.RandomState(20261019)
    out = {}
    for k, ((H, W), S) in enumerate(CASES):
        frames = rs.randint(0, 256, (2, H, W, 3)).astype(np.uint8)
        yy, xx = np.mgrid[0:H, 0:W]
        frames[1] = np.stack([yy * 255 // max(H - 1, 1), xx * 255 // max(W - 1, 1), (yy + 2 * xx) % 256], -1).astype(np.uint8)   # smooth ramps
        res = np.stack([cv2.cvtColor(cv2.resize(f, (S, S), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2RGB) for f in frames])
        out[f"src_{k}"] = frames
        out[f"size_{k}"] = np.int64(S)
        out[f"resized_rgb_{k}"] = res                                           # uint8, after BGR -> RGB
        out[f"input_{k}"] = (np.array(res) / 127.5 - 1.).astype(np.float32)     # utils.py:63, fed to a float32 placeholder
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "frames_resize.npz"), **out)
    print("wrote frames_resize.npz", len(CASES), "cases, cv2", cv2.__version__)


if __name__ == "__main__":
    main()
