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


def test_particle_colours_follow_the_image(built, golden, tmp_path):
    """KaminoParticles(path, ...): colorBGR = (G, R, B) / 255 of the mirrored, resized image at the
    particle's cell (kernel/KaminoParticles.cu:56-72); positions from the rand()-driven lattice."""
    exe = str(tmp_path / "particle_colour_check")
    host = os.path.join(ROOT, "kaminogpu_b200", "host")
    build = subprocess.run(["g++", "-std=c++17", "-O2", "-DKAMINO_HAVE_ZLIB", "-I" + os.path.join(ROOT, "include"), "-I" + host,
                            "-o", exe, os.path.join(ROOT, "tests", "native", "particle_colour_check.cpp"),
                            os.path.join(host, "KaminoParticles.cpp"), os.path.join(host, "KaminoQuantity.cpp"), os.path.join(host, "ImageIO.cpp"),
                            "-L" + os.path.join(ROOT, "kaminogpu_b200"), "-lkamino_b200", "-lz",
                            "-Wl,-rpath," + os.path.join(ROOT, "kaminogpu_b200")], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    nTheta, density = 32, 4.0
    out = str(tmp_path / "particles.bin")
    run = subprocess.run([exe, os.path.join(GOLDEN, "images", "smooth.png"), str(nTheta), str(density), out],
                         capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    raw = open(out, "rb").read()
    n = int(np.frombuffer(raw[:8], dtype=np.int64)[0])
    assert n == 64 * 128                                      # numTheta = sqrt(4) * 32, numPhi = 2 numTheta
    coords = np.frombuffer(raw[8:8 + 8 * n], dtype=np.float32).reshape(n, 2)
    colours = np.frombuffer(raw[8 + 8 * n:], dtype=np.float32).reshape(n, 3)
    image = golden["smooth.png.32x64"]                        # [theta][phi][B, G, R]
    h = np.float32(np.pi / nTheta)
    x = np.minimum(np.floor(coords[:, 0] / h).astype(np.int64), 2 * nTheta - 1)
    y = np.minimum(np.floor(coords[:, 1] / h).astype(np.int64), nTheta - 1)
    px = image[y, x].astype(np.float64) / 255.0
    want = np.stack([px[:, 1], px[:, 2], px[:, 0]], axis=1).astype(np.float32)
    assert np.array_equal(colours.view(np.uint32), want.view(np.uint32))
