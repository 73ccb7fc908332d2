"""GPU parity of the frame-decode tail (gg_frames_to_input through gifgan.ops.frames_to_input) -- byte / integer work, so the bar
is BIT-EXACT: against the oracle restatement (oracle/image_ops.py, pinned to cv2.resize) on seeded frames, against the committed
cv2-generated fixture, and at the full config-3 batch (32 clips x 16 frames) through size-independent properties."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import image_ops as I  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frames_resize.npz")


def _gpu(frames, size, swap_rb=True):
    from gifgan import ops
    return ops.frames_to_input(torch.from_numpy(frames).cuda(), size, swap_rb=swap_rb).cpu().numpy()


def test_frames_match_the_cv2_fixture_bit_for_bit():
    g = np.load(GOLD)
    k = 0
    while f"src_{k}" in g:
        got = _gpu(g[f"src_{k}"], int(g[f"size_{k}"]))
        assert got.dtype == np.float32 and np.array_equal(got, g[f"input_{k}"]), k
        k += 1
    assert k >= 7


@pytest.mark.parametrize("H,W,S", [(96, 128, 64), (128, 128, 64), (64, 64, 64), (48, 40, 64), (75, 131, 64), (240, 320, 64), (360, 480, 128),
                                   (1, 1, 4), (2, 3, 1), (63, 200, 64), (17, 500, 32), (12, 12, 6), (8, 8, 4)])
def test_frames_match_the_oracle(H, W, S):
    rs = np.random.RandomState(H * 7 + W)
    fr = rs.randint(0, 256, (5, H, W, 3)).astype(np.uint8)
    assert np.array_equal(_gpu(fr, S), I.frames_to_input(fr, S))
    assert np.array_equal(_gpu(fr, S, swap_rb=False), I.frames_to_input(fr, S, swap_rb=False))


def test_frames_full_batch_properties_and_out_buffer():
    """Config 3's batch (32 clips x 16 frames at 128 x 128 -> 64 x 64, the 2x shortcut; and at 96 x 128 -> 64 x 64, the fixed-point
    path): (i) every frame equals the same frame converted alone (no cross-frame indexing error at scale); (ii) constant frames
    map to the constant's table value; (iii) channel swap commutes with the conversion; (iv) `out=` writes in place."""
    from gifgan import ops
    rs = np.random.RandomState(3)
    for H, W in ((128, 128), (96, 128)):
        fr = rs.randint(0, 256, (512, H, W, 3)).astype(np.uint8)
        fr[7] = 200
        fr[8, :, :, :] = np.array([10, 20, 30], dtype=np.uint8)
        d = torch.from_numpy(fr).cuda()
        out = torch.full((512, 64, 64, 3), 7.0, device="cuda")
        got = ops.frames_to_input(d, 64, out=out)
        assert got.data_ptr() == out.data_ptr()
        g = got.cpu().numpy()
        for i in (0, 1, 255, 511):
            assert np.array_equal(g[i], I.frames_to_input(fr[i:i + 1], 64)[0])
        assert np.all(g[7] == I.NORMALIZE_LUT[200])
        assert np.all(g[8] == I.NORMALIZE_LUT[np.array([30, 20, 10])])          # BGR (10, 20, 30) -> RGB
        assert np.array_equal(ops.frames_to_input(d.flip(-1).contiguous(), 64, swap_rb=False).cpu().numpy(), g)
        assert g.min() >= -1.0 and g.max() <= 1.0


def test_frames_rejects_bad_input():
    from gifgan import ops
    with pytest.raises(ValueError):
        ops.frames_to_input(torch.zeros(2, 8, 8, 3, device="cuda"), 4)               # float frames
    with pytest.raises(ValueError):
        ops.frames_to_input(torch.zeros(2, 8, 8, 4, dtype=torch.uint8, device="cuda"), 4)
    with pytest.raises(RuntimeError):
        ops.frames_to_input(torch.zeros(2, 8, 8, 3, dtype=torch.uint8), 4)           # host tensor: no CPU fallback
