"""Raw checkpoint serialisation (kaminogpu_b200/host/Checkpoint.cpp): bit-exact round
trip (including -0.0 and NaN payloads), size validation, checksum and truncation detection. CPU only;
the resume-equals-uninterrupted property of the CLI is in the GPU suite."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_checkpoint_round_trip_and_corruption_detection(tmp_path):
    exe = str(tmp_path / "checkpoint_check")
    build = subprocess.run(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "native", "checkpoint_check.cpp"),
                            os.path.join(ROOT, "kaminogpu_b200", "host", "Checkpoint.cpp")], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([exe, str(tmp_path / "state.ck")], capture_output=True, text=True)
    assert run.returncode == 0 and "checkpoint check ok" in run.stdout, run.stdout + run.stderr
