"""TEST INFRASTRUCTURE ONLY (oracle): the tail of the reference's frame decode, restated in numpy integer arithmetic.

The reference turns every decoded frame into a network input on the host, per frame, inside the step loop
(`/root/reference/models/recurrent_z/z_model_lib.py:339-346`):

    im = cv2.resize(im, (S, S), interpolation=cv2.INTER_LINEAR)      # uint8 BGR [H0, W0, 3] -> [S, S, 3]
    im = cv2.cvtColor(im, cv2.COLOR_BGR2RGB)
    im = transform(im, is_crop=False)                                # utils.py:57-63: np.array(im) / 127.5 - 1.

The algorithm of `cv2.resize` lives in a third-party dependency that is not under /root/reference: OpenCV (the reference pins
no version; this image has opencv-python 4.13.0).  Its published algorithm for 8-bit INTER_LINEAR (modules/imgproc/src/resize.cpp:
`resizeGeneric_` with `HResizeLinear` / `VResizeLinear<uchar, int, short, FixedPtCast<int, uchar, INTER_RESIZE_COEF_BITS*2>>`)
is fixed-point:

  * source coordinate of destination index d along an axis of scale = src/dst:  f = (d + 0.5) * scale - 0.5;  s = floor(f);
    f -= s;  horizontally  s < 0 -> (s, f) = (0, 0),  s >= src - 1 -> (s, f) = (src - 1, 0);  vertically (s, f) stay as they are
    and the two ROW indices s, s + 1 are clamped into the image when the rows are fetched;
  * 11-bit coefficients  c0 = saturate_short(rint((1 - f) * 2048)),  c1 = saturate_short(rint(f * 2048))  (float32 arithmetic);
  * horizontal pass in int32:  r[x] = S[s] * c0 + S[s + 1] * c1                      (S[s + 1] is not read when c1 applies to s = src - 1)
  * vertical pass:             dst = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
  * and one shortcut taken before any of that: an exact 2x reduction in both axes is computed as INTER_AREA, i.e.
    dst = (s00 + s01 + s10 + s11 + 2) >> 2.

PINNED: `tests/test_oracle_image_ops.py` checks this restatement bit for bit against `cv2.resize` itself (installed here) over a sweep
of source / destination sizes, and `tests/golden/frames_resize.npz` holds cv2-generated vectors (generator: tests/golden/make_golden_frames.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module."""
import numpy as np

COEF_BITS = 11
COEF_ONE = 1 << COEF_BITS


def _axis_table(src: int, dst: int, clamp: bool):
    """(index s, coefficients c0, c1) for every destination index along one axis (resize.cpp: the xofs / ialpha, yofs / ibeta tables).
    clamp=True is the HORIZONTAL rule (index and fraction are clamped at the borders: (0, 0) / (src - 1, 0)); the VERTICAL tables
    keep the unclamped fraction and the row INDICES are clamped when the rows are fetched (both taps then read the border row, each
    with its own coefficient -- a different rounding from the horizontal rule, visible when an axis is enlarged)."""
    scale = np.float64(1.0) / (np.float64(dst) / np.float64(src))      # cv::resize: inv_scale = (double)dst / src; scale = 1. / inv_scale
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)          # OpenCV: fx = (float)((dx + 0.5) * scale_x - 0.5)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp:
        lo = s < 0
        s[lo] = 0
        f[lo] = 0.0
        hi = s >= src - 1
        s[hi] = src - 1
        f[hi] = 0.0
    c0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_ONE)).astype(np.int64)     # cvRound: round half to even, like rint
    c1 = np.rint(f * np.float32(COEF_ONE)).astype(np.int64)
    return s, np.clip(c0, -32768, 32767), np.clip(c1, -32768, 32767)


def resize_linear_u8(img: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) for uint8 [H, W, C] images, bit for bit."""
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W, _ = img.shape
    if H == 2 * dst_h and W == 2 * dst_w:                      # exact 2x reduction: the INTER_AREA fast path
        a = img.astype(np.int64)
        return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx, a0, a1 = _axis_table(W, dst_w, True)
    sy, b0, b1 = _axis_table(H, dst_h, False)
    src = img.astype(np.int64)
    sx1 = np.minimum(sx + 1, W - 1)                            # its coefficient is 0 wherever the clamp applies
    rows = src[:, sx, :] * a0[None, :, None] + src[:, sx1, :] * a1[None, :, None]           # horizontal pass, [H, dst_w, C]
    r0, r1 = rows[np.clip(sy, 0, H - 1)], rows[np.clip(sy + 1, 0, H - 1)]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


# utils.py:63  np.array(im) / 127.5 - 1.  evaluated in float64, then fed to a float32 placeholder
NORMALIZE_LUT = (np.arange(256, dtype=np.float64) / 127.5 - 1.0).astype(np.float32)


def frames_to_input(frames_bgr_u8: np.ndarray, size: int, swap_rb: bool = True) -> np.ndarray:
    """[N, H0, W0, 3] uint8 decoded frames (OpenCV order: BGR) -> [N, size, size, 3] float32 network input in [-1, 1]
    (z_model_lib.py:339-346: resize, BGR -> RGB, / 127.5 - 1)."""
    out = np.empty((frames_bgr_u8.shape[0], size, size, 3), dtype=np.float32)
    for n, fr in enumerate(frames_bgr_u8):
        r = resize_linear_u8(fr, size, size) if fr.shape[:2] != (size, size) else fr
        if swap_rb:
            r = r[:, :, ::-1]
        out[n] = NORMALIZE_LUT[r]
    return out
