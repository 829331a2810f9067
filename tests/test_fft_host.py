"""CPU check of the FFT index logic (kaminogpu_b200/csrc/fft_core.cuh): the radix-16 Stockham passes
and the packed per-pass twiddle tables are executed on the host, "thread" by "thread", and compared
with a direct DFT in double precision (tests/native/fft_host_check.cu). Needs nvcc, no GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_fft_passes_match_direct_dft(tmp_path):
    exe = str(tmp_path / "fft_host_check")
    src = os.path.join(ROOT, "tests", "native", "fft_host_check.cu")
    build = subprocess.run(["nvcc", "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-o", exe, src],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([exe], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout
    assert run.stdout.count(" ok") == 22 and "FAIL" not in run.stdout, run.stdout
