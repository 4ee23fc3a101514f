"""Pins oracle/cnn_obs_oracle.py (the ResizeObservation / GrayscaleObservation restatement) against OpenCV itself."""
import numpy as np
import pytest

from oracle.cnn_obs_oracle import cnn_frame, grayscale, resize_area_u8

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("shape", [(24, 34), (24, 46), (44, 48), (14, 23), (64, 96), (34, 33), (32, 40)], ids=str)
def test_resize_area_matches_cv2(shape):
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    palette = np.array([[0, 0, 0], [128, 128, 128], [0, 240, 240], [240, 240, 0], [160, 0, 240], [0, 240, 0], [240, 0, 0],
                        [0, 0, 240], [240, 160, 0]], np.uint8)
    for k in range(12):
        img = rng.integers(0, 256, size=shape + (3,), dtype=np.uint8) if k % 2 else palette[rng.integers(0, 9, size=shape)]
        want = cv2.resize(img, (84, 84), interpolation=cv2.INTER_AREA)
        got = resize_area_u8(img, (84, 84))
        assert got.shape == want.shape == (84, 84, 3)
        assert np.array_equal(got, want), (shape, k, np.abs(got.astype(int) - want.astype(int)).max())
    g = rng.integers(0, 256, size=(84, 84, 3), dtype=np.uint8)
    assert np.array_equal(grayscale(g), np.sum(np.multiply(g, np.array([0.2125, 0.7154, 0.0721])), axis=-1).astype(np.uint8))
    assert cnn_frame(img).shape == (84, 84) and cnn_frame(img).dtype == np.uint8
