"""Image-driven initialisation without OpenCV (kaminogpu_b200/host/ImageIO.cpp) against vectors
produced by the OpenCV calls the reference makes -- cv2.imread(IMREAD_COLOR), cv2.flip(., 1),
cv2.resize(., (nPhi, nTheta)) -- committed in tests/golden/image_init.npz by
tests/golden/make_resize_goldens.py (kernel/KaminoSolver.cu:243-277). Bar: bit-exact pixels and
bit-exact density floats. CPU only."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
IMAGES = ["smooth.png", "noise.ppm", "grey.pgm", "grey.png", "big.ppm", "rgba.png", "ascii.ppm"]
SIZES = [16, 32, 64]


@pytest.fixture(scope="module")
def tool(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("image") / "image_check")
    build = subprocess.run(["g++", "-std=c++17", "-O2", "-DKAMINO_HAVE_ZLIB", "-o", exe,
                            os.path.join(ROOT, "tests", "native", "image_check.cpp"),
                            os.path.join(ROOT, "kaminogpu_b200", "host", "ImageIO.cpp"), "-lz"],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    return exe


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "image_init.npz"))


@pytest.mark.parametrize("name", IMAGES)
@pytest.mark.parametrize("nTheta", SIZES)
def test_read_flip_resize_density_match_opencv(tool, golden, tmp_path, name, nTheta):
    out = str(tmp_path / "dump.bin")
    run = subprocess.run([tool, os.path.join(GOLDEN, "images", name), str(nTheta), out], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stderr)
    raw = np.fromfile(out, dtype=np.uint8)
    w, h = np.frombuffer(raw[:8].tobytes(), dtype=np.int32)
    read = golden[name + ".read"]
    assert (h, w) == read.shape[:2]
    p = 8
    got_read = raw[p:p + w * h * 3].reshape(h, w, 3); p += w * h * 3
    nPhi = 2 * nTheta
    got_resized = raw[p:p + nTheta * nPhi * 3].reshape(nTheta, nPhi, 3); p += nTheta * nPhi * 3
    got_density = np.frombuffer(raw[p:].tobytes(), dtype=np.float32).reshape(nTheta, nPhi)
    assert np.array_equal(got_read, read), "decoded pixels differ from cv2.imread"
    assert np.array_equal(got_resized, golden["%s.%dx%d" % (name, nTheta, nPhi)]), "resize differs from cv2.resize"
    want = golden["%s.%dx%d.density" % (name, nTheta, nPhi)]
    assert np.array_equal(got_density.view(np.uint32), want.view(np.uint32)), "density differs"


def test_missing_or_undecodable_image_is_reported(tool, tmp_path):
    out = str(tmp_path / "dump.bin")
    assert subprocess.run([tool, str(tmp_path / "nope.png"), "16", out]).returncode == 3
    junk = tmp_path / "junk.png"
    junk.write_bytes(b"not an image")
    assert subprocess.run([tool, str(junk), "16", out]).returncode == 3
