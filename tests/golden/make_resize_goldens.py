#!/usr/bin/env python
"""Golden vectors for the OpenCV-free image initialisation (kaminogpu_b200/host/ImageIO.cpp).

Run in the build container (needs cv2, the library family the reference links against):
    python tests/golden/make_resize_goldens.py
Writes small test images under tests/golden/images/ and tests/golden/image_init.npz holding, for
every (image, target size) pair, what the reference's three OpenCV calls produce:
    cv2.imread(path, cv2.IMREAD_COLOR) -> cv2.flip(., 1) -> cv2.resize(., (nPhi, nTheta))
(kernel/KaminoSolver.cu:249-262, kernel/KaminoParticles.cu:7-18), plus the density field the
reference derives from it ((B + G + R) / 3 with fReal = float, kernel/KaminoSolver.cu:264-274).
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
IMG = os.path.join(HERE, "images")


def main():
    os.makedirs(IMG, exist_ok=True)
    rng = np.random.default_rng(20261017)
    smooth = np.zeros((48, 64, 3), np.uint8)
    yy, xx = np.mgrid[0:48, 0:64]
    smooth[..., 0] = (127 + 120 * np.sin(xx / 7.0) * np.cos(yy / 5.0)).astype(np.uint8)
    smooth[..., 1] = (xx * 4) % 256
    smooth[..., 2] = (yy * 5 + xx) % 256
    noise = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    grey = rng.integers(0, 256, (20, 31), dtype=np.uint8)
    big = rng.integers(0, 256, (64, 128, 3), dtype=np.uint8)        # exactly 2x of 32 x 64: the INTER_AREA shortcut
    files = {}
    cv2.imwrite(os.path.join(IMG, "smooth.png"), smooth); files["smooth.png"] = None
    cv2.imwrite(os.path.join(IMG, "noise.ppm"), noise); files["noise.ppm"] = None
    cv2.imwrite(os.path.join(IMG, "grey.pgm"), grey); files["grey.pgm"] = None
    cv2.imwrite(os.path.join(IMG, "grey.png"), grey); files["grey.png"] = None
    cv2.imwrite(os.path.join(IMG, "big.ppm"), big); files["big.ppm"] = None
    rgba = np.dstack([noise[:16, :24], rng.integers(0, 256, (16, 24, 1), dtype=np.uint8)])
    cv2.imwrite(os.path.join(IMG, "rgba.png"), rgba); files["rgba.png"] = None
    with open(os.path.join(IMG, "ascii.ppm"), "w") as f:               # P3 with a comment and maxval 255
        f.write("P3\n# ascii sample\n3 2\n255\n255 0 0  0 255 0  0 0 255\n10 20 30  40 50 60  70 80 90\n")
    files["ascii.ppm"] = None
    out = {}
    sizes = [(16, 32), (32, 64), (64, 128)]                           # (nTheta, nPhi)
    for name in files:
        img = cv2.imread(os.path.join(IMG, name), cv2.IMREAD_COLOR)
        assert img is not None, name
        out[name + ".read"] = img
        flipped = cv2.flip(img, 1)
        for nT, nP in sizes:
            resized = cv2.resize(flipped, (nP, nT))
            out["%s.%dx%d" % (name, nT, nP)] = resized
            b = (resized[..., 0] / 255.0).astype(np.float32)
            g = (resized[..., 1] / 255.0).astype(np.float32)
            r = (resized[..., 2] / 255.0).astype(np.float32)
            out["%s.%dx%d.density" % (name, nT, nP)] = (((b + g).astype(np.float32) + r).astype(np.float32) / 3.0).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "image_init.npz"), **out)
    print("wrote", len(out), "arrays; cv2", cv2.__version__)


if __name__ == "__main__":
    main()
