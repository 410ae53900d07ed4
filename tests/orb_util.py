"""Shared inputs of the ORB tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def real_gray():
    import cv2
    g = cv2.imread(os.path.join(GOLDEN, "orb_crazyhorse_gray0.png"), cv2.IMREAD_GRAYSCALE)
    assert g is not None and g.shape == (768, 1024)
    return g


def blobs(h, w, seed, n=400, blur=False):
    """Rectangles and discs of random grey levels: plenty of FAST corners, many exactly equal scores / Harris responses (ties)."""
    import cv2
    r = np.random.RandomState(seed); img = np.full((h, w), 120, np.uint8)
    for _ in range(n):
        x, y = int(r.randint(0, w)), int(r.randint(0, h)); s = int(r.randint(3, 40)); c = int(r.randint(0, 256))
        if r.rand() < 0.5:
            cv2.rectangle(img, (x, y), (x + s, y + s), c, -1)
        else:
            cv2.circle(img, (x, y), s // 2, c, -1)
    return cv2.GaussianBlur(img, (5, 5), 1.2) if blur else img


def noise(h, w, seed, channels=1):
    shape = (h, w) if channels == 1 else (h, w, channels)
    return np.random.RandomState(seed).randint(0, 256, shape, dtype=np.uint8)


def textured(h, w, seed):
    """Smooth random texture + blobs: a 'natural-looking' image with sub-pixel structure on every pyramid level."""
    import cv2
    r = np.random.RandomState(seed)
    base = cv2.GaussianBlur(r.randint(0, 256, (h, w)).astype(np.float32), (0, 0), 2.5)
    base = (base - base.min()) / (base.max() - base.min()) * 255
    img = 0.6 * base + 0.4 * blobs(h, w, seed + 1, 600).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


# (name, image factory, nfeatures): the parity cases shared by the oracle (CPU) and the GPU tests
CASES = [
    ("noise_gray_640x480", lambda: noise(480, 640, 1), 5000),
    ("noise_bgr_640x480", lambda: noise(480, 640, 2, 3), 5000),
    ("blobs_1024x768", lambda: blobs(768, 1024, 3), 5000),
    ("blobs_blurred_800x600", lambda: blobs(600, 800, 4, 1500, True), 5000),
    ("odd_517x333_nf500", lambda: blobs(333, 517, 5), 500),
    ("small_150x200", lambda: blobs(200, 150, 6), 5000),
    ("tiny_100x90", lambda: blobs(90, 100, 7), 100),
    ("tiny_64x64", lambda: blobs(64, 64, 8), 100),
    ("flat_400x300", lambda: np.full((300, 400), 77, np.uint8), 5000),
    ("nf20000_1024x768", lambda: blobs(768, 1024, 9, 3000), 20000),
    ("textured_1024x768", lambda: textured(768, 1024, 10), 5000),
    ("textured_bgr_700x500", lambda: np.stack([textured(500, 700, 11 + c) for c in range(3)], 2), 2000),
    ("large_2000x1500_nf8000", lambda: textured(1500, 2000, 14), 8000),
    ("wide_1900x120", lambda: textured(120, 1900, 15), 3000),
    ("tall_90x1300", lambda: blobs(1300, 90, 16, 900), 3000),
]
