"""CPU tests of the frame-decode tail (reference: models/recurrent_z/z_model_lib.py:339-346, utils.py:57-63).
The oracle restatement (oracle/image_ops.py) is PINNED here: bit for bit against cv2.resize itself over a sweep of sizes (when
OpenCV is importable, as in this image) and against the committed cv2-generated fixture tests/golden/frames_resize.npz."""
import os

import numpy as np
import pytest

from oracle import image_ops as I

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frames_resize.npz")


def _cases():
    g = np.load(GOLD)
    k = 0
    while f"src_{k}" in g:
        yield g[f"src_{k}"], int(g[f"size_{k}"]), g[f"resized_rgb_{k}"], g[f"input_{k}"]
        k += 1


def test_restatement_matches_the_cv2_fixture_bit_for_bit():
    n = 0
    for src, S, res_rgb, inp in _cases():
        for f, want in zip(src, res_rgb):
            got = I.resize_linear_u8(f, S, S)[:, :, ::-1]
            assert np.array_equal(got, want)
        got = I.frames_to_input(src, S)
        assert got.dtype == np.float32 and np.array_equal(got, inp)         # floats compared bit for bit
        n += 1
    assert n >= 7


def test_restatement_matches_cv2_over_a_size_sweep():
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(5)
    sizes = [(1, 1, 1, 1), (1, 7, 3, 3), (2, 2, 1, 1), (128, 128, 64, 64), (64, 64, 128, 128), (63, 200, 64, 64), (240, 320, 64, 64),
             (360, 480, 128, 128), (17, 500, 64, 32)]
    sizes += [tuple(int(v) for v in rs.randint(1, 300, 4)) for _ in range(120)]
    for H, W, h, w in sizes:
        img = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
        want = cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(I.resize_linear_u8(img, h, w), want.reshape(h, w, 3)), (H, W, h, w)


def test_normalisation_is_the_reference_formula():
    x = np.arange(256, dtype=np.uint8)
    want = (np.array(x) / 127.5 - 1.).astype(np.float32)                    # utils.py:63
    assert np.array_equal(I.NORMALIZE_LUT, want)
    assert I.NORMALIZE_LUT[0] == -1.0 and I.NORMALIZE_LUT[255] == 1.0


def test_identity_size_and_channel_order():
    rs = np.random.RandomState(1)
    fr = rs.randint(0, 256, (2, 64, 64, 3)).astype(np.uint8)
    got = I.frames_to_input(fr, 64)
    assert np.array_equal(got, I.NORMALIZE_LUT[fr[..., ::-1]])              # same size: a copy; BGR -> RGB
    assert np.array_equal(I.frames_to_input(fr, 64, swap_rb=False), I.NORMALIZE_LUT[fr])
    assert np.array_equal(I.resize_linear_u8(fr[0], 64, 64), fr[0])         # the fixed-point pipeline is exact at scale 1
