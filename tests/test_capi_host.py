"""CPU suite, part 2: the C-ABI library loads, exports every symbol include/kamino_b200.h
declares, its host-side initialisers reproduce the reference's initial state bit for bit,
and it fails loudly (no CPU fallback) when no CUDA device is present. No compute calls."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_api as oa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(built):
    from kaminogpu_b200 import capi
    return capi.load()


def test_every_declared_symbol_is_exported(lib):
    from kaminogpu_b200 import capi
    header = open(os.path.join(ROOT, "include", "kamino_b200.h")).read()
    declared = set(re.findall(r"\b(kamino_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)


def test_header_is_plain_c(built, tmp_path):
    """The drop-in boundary is a C ABI: include/kamino_b200.h must compile as C99 with nothing but the standard headers
    (no CUDA, torch or C++ types in any signature), and a C program must link against the library by symbol name."""
    src = tmp_path / "cabi.c"
    src.write_text('#include "kamino_b200.h"\n'
                   'int main(void) {\n'
                   '    kamino_ctx* c = 0; kamino_dist* d = 0; unsigned char id[128];\n'
                   '    int rc = kamino_create(&c, 0, 7, 5.0f, 0.005f, 1, 0);          /* nTheta = 7: rejected before any device work */\n'
                   '    if (rc != KAMINO_ERR_INVALID || c != 0) return 1;\n'
                   '    rc = kamino_dist_create(&d, 0, 100, 5.0f, 0.005f, 0, 1, id);\n'
                   '    if (rc != KAMINO_ERR_INVALID || d != 0) return 2;\n'
                   '    return kamino_version()[0] == \'k\' ? 0 : 3;\n'
                   '}\n')
    exe = tmp_path / "cabi"
    pkg = os.path.join(ROOT, "kaminogpu_b200")
    build = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src),
                            "-o", str(exe), "-L" + pkg, "-lkamino_b200", "-Wl,-rpath," + pkg], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stderr)


def test_library_is_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_version_string(lib):
    assert b"sm_100a" in lib.kamino_version()


@pytest.mark.parametrize("case", ["t16", "t32", "t64", "t128"])
def test_initial_velocity_matches_reference_dump(lib, case):
    g = oa.golden(case)
    nT = int(g["meta.nTheta"])
    u = np.zeros(nT * 2 * nT, np.float32)
    v = np.zeros((nT - 1) * 2 * nT, np.float32)
    assert lib.kamino_init_velocity_host(nT, ctypes.c_float(float(g["meta.radius"])), u.ctypes.data, v.ctypes.data) == 0
    assert np.array_equal(u, g["init.velPhi"])
    assert np.array_equal(v, g["init.velTheta"])


@pytest.mark.parametrize("case,density", [("t16", 4.0), ("t32", 4.0), ("t64", 1.0), ("t128", 1.0)])
def test_particle_seeding_matches_reference_dump(lib, case, density):
    g = oa.golden(case)
    nT = int(g["meta.nTheta"])
    n = lib.kamino_particle_count(nT, ctypes.c_float(density))
    assert n == int(g["meta.numParticles"])
    pc = np.zeros(2 * n, np.float32)
    assert lib.kamino_seed_particles_host(nT, ctypes.c_float(density), pc.ctypes.data) == 0
    assert np.array_equal(pc, g["init.particles"])


def test_particle_seeding_leaves_the_process_rand_state_alone(lib):
    """The library restates glibc's rand() locally (csrc/host_init.cpp, GlibcRand): same sequence as libc's
    never-seeded generator (checked against the oracle, which calls srand(1) / rand()), and no side effect
    on the host program's generator."""
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(12345)
    first = [libc.rand() for _ in range(3)]
    libc.srand(12345)
    n = lib.kamino_particle_count(16, ctypes.c_float(4.0))
    mine = np.zeros(2 * n, np.float32)
    assert lib.kamino_seed_particles_host(16, ctypes.c_float(4.0), oa.fptr(mine)) == 0
    assert [libc.rand() for _ in range(3)] == first
    assert np.array_equal(mine, oa.seed_particles(16, 4.0).ravel())


def test_particle_counts_of_the_baseline_configs(lib):
    # C1 particleDensity 200 -> 6,552,200 ; C2 particleDensity 2 -> 1,048,352
    assert lib.kamino_particle_count(128, ctypes.c_float(200.0)) == 6552200
    assert lib.kamino_particle_count(512, ctypes.c_float(2.0)) == 1048352
    assert lib.kamino_particle_count(512, ctypes.c_float(4.0)) == 2097152
    assert lib.kamino_particle_count(512, ctypes.c_float(0.0)) == 0


def test_bad_arguments_are_rejected_without_a_device(lib):
    ctx = ctypes.c_void_p()
    for nT in (0, 8, 100, 16384):
        rc = lib.kamino_create(ctypes.byref(ctx), 0, nT, ctypes.c_float(5.0), ctypes.c_float(0.005), 1, 0)
        assert rc != 0 and not ctx.value
    assert lib.kamino_create(ctypes.byref(ctx), 0, 64, ctypes.c_float(-1.0), ctypes.c_float(0.005), 1, 0) != 0
    assert lib.kamino_advect(None) != 0
    assert lib.kamino_step(None, 1) != 0
    assert lib.kamino_destroy(None) == 0


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product refuses to create a context."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    from kaminogpu_b200 import capi
    from kaminogpu_b200.solver import KaminoSolver
    with pytest.raises(capi.KaminoError) as e:
        KaminoSolver(64, 32, 5.0, 0.005)
    assert e.value.code == 10002 and "no CPU path" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under kaminogpu_b200/ may mention it."""
    pkg = os.path.join(ROOT, "kaminogpu_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath or "__pycache__" in dirpath:
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or fn == "Makefile":
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, fn)


def test_cli_usage_and_config_grammar(built, tmp_path):
    exe = os.path.join(ROOT, "kaminogpu_b200", "kamino")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode != 0 and "configKamino.txt" in out.stdout
    from kaminogpu_b200.solver import load_config
    cfg = tmp_path / "configKamino.txt"
    cfg.write_text("5.0 128 200.0 0.005 0.041666668 10 0.0 1 1 1 1 out/f out/p null null null\n")
    c = load_config(str(cfg))
    assert c["nTheta"] == 128 and c["frames"] == 10 and c["densityImage"] == "" and c["solidImage"] == "null"
    assert c["gridPath"] == "out/f" and abs(c["DT"] - 1 / 24) < 1e-6


def test_steps_per_frame_rule():
    """Kamino::run takes (iterations of `while (T < i*DT)`) + 1 steps per frame
    (kernel/KaminoCore.cu:887-895): 10 at the default dt = 0.005, DT = 1/24."""
    from kaminogpu_b200.solver import steps_per_frame
    assert steps_per_frame(0.005, 1.0 / 24.0, 10) == [10] * 10
    assert steps_per_frame(1.0 / 24.0, 1.0 / 24.0, 3) == [2, 2, 2]
