"""CPU restatement of the trainer-side image adapters of examples/train_cnn.py:140-143 (TEST INFRASTRUCTURE ONLY):

    env = gym.wrappers.ResizeObservation(env, (84, 84))      # cv2.resize(obs, (84, 84), interpolation=cv2.INTER_AREA)
    env = gym.wrappers.GrayscaleObservation(env)             # sum(obs * [0.2125, 0.7154, 0.0721], -1).astype(uint8)
    env = gym.wrappers.FrameStackObservation(env, 4)         # last 4 frames, reset frame repeated at reset

Both wrappers are third-party code that is absent from /root/reference (gymnasium, pyproject pin 1.1.1; its
ResizeObservation calls OpenCV, lock pin opencv-python 4.11).  The arithmetic is restated from the published algorithms:

* cv2.resize, 8-bit, INTER_AREA with at least one axis enlarged (always the case here: H_pad <= 64 < 84): OpenCV emulates
  area interpolation by a bilinear kernel with `area_mode` coordinates (modules/imgproc/src/resize.cpp, cv::hal::resize ->
  resizeGeneric_ / HResizeLinear / VResizeLinear): fixed point, 11-bit coefficients,
      h(y, dx)  = S[y, sx] * a0 + S[y, sx + 1] * a1                               (int32)
      D[dy, dx] = (((b0 * (h(sy, dx) >> 4)) >> 16) + ((b1 * (h(sy + 1, dx) >> 4)) >> 16) + 2) >> 2
  PINNED against cv2 itself (4.13 in this image) on random images by tests/test_oracle_cnn_obs.py.
* GrayscaleObservation: float64 products summed left to right, truncated -- pinned against the numpy expression.
"""
import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def _area_coeffs(ssize, dsize):
    """offsets and 11-bit coefficient pairs of one axis (cv::hal::resize, linear branch with area_mode = true)."""
    inv_scale = np.float64(dsize) / np.float64(ssize)
    scale = np.float64(1.0) / inv_scale
    ofs = np.zeros(dsize, np.int32)
    coef = np.zeros((dsize, 2), np.int32)
    dmax = dsize
    for d in range(dsize):
        s = int(np.floor(d * scale))
        f = np.float32((d + 1) - (s + 1) * inv_scale)
        f = np.float32(0.0) if f <= 0 else np.float32(f - np.float32(np.floor(f)))
        if s < 0:
            f, s = np.float32(0.0), 0
        if s + 1 >= ssize:
            dmax = min(dmax, d)
            if s >= ssize - 1:
                f, s = np.float32(0.0), ssize - 1
        ofs[d] = s
        c0, c1 = np.float32(1.0) - f, f
        coef[d, 0] = int(np.clip(np.rint(np.float32(c0 * np.float32(COEF_SCALE))), -32768, 32767))
        coef[d, 1] = int(np.clip(np.rint(np.float32(c1 * np.float32(COEF_SCALE))), -32768, 32767))
    return ofs, coef, dmax


def resize_area_u8(img, dsize=(84, 84)):
    """cv2.resize(img, dsize, interpolation=cv2.INTER_AREA) for uint8 [H, W, C] when not both axes shrink."""
    img = np.asarray(img, np.uint8)
    H, W = img.shape[:2]
    dw, dh = dsize
    assert not (W >= dw and H >= dh), "true area interpolation (both axes shrinking) is not restated"
    xofs, alpha, xmax = _area_coeffs(W, dw)
    yofs, beta, _ = _area_coeffs(H, dh)
    S = img.astype(np.int32)
    # horizontal pass for every source row
    x1 = np.minimum(xofs + 1, W - 1)
    hbuf = S[:, xofs] * alpha[None, :, 0, None] + S[:, x1] * alpha[None, :, 1, None]
    if xmax < dw:
        hbuf[:, xmax:] = S[:, xofs[xmax:]] * COEF_SCALE
    out = np.empty((dh, dw) + img.shape[2:], np.uint8)
    for dy in range(dh):
        sy0 = int(np.clip(yofs[dy], 0, H - 1)); sy1 = int(np.clip(yofs[dy] + 1, 0, H - 1))
        b0, b1 = int(beta[dy, 0]), int(beta[dy, 1])
        v = (((b0 * (hbuf[sy0] >> 4)) >> 16) + ((b1 * (hbuf[sy1] >> 4)) >> 16) + 2) >> 2
        out[dy] = np.clip(v, 0, 255).astype(np.uint8)
    return out


GRAY_W = np.array([0.2125, 0.7154, 0.0721])


def grayscale(img):
    """gymnasium GrayscaleObservation (1.x): np.sum(np.multiply(obs, [0.2125, 0.7154, 0.0721]), axis=-1).astype(np.uint8)."""
    return np.sum(np.multiply(img, GRAY_W), axis=-1).astype(np.uint8)


def cnn_frame(rgb, dsize=(84, 84)):
    """RgbObservation frame -> ResizeObservation -> GrayscaleObservation: u8[84, 84]."""
    return grayscale(resize_area_u8(rgb, dsize))
