"""Pack raw state dumps of the reference CUDA build into the committed golden fixtures.

How the fixtures were made (round 1, on a B200 through gpurun):
  1. oracle/ref_harness/Makefile builds oracle/_ref/kamino_ref from the UNMODIFIED sources
     under /root/reference/KaminoGPU/kernel (+ a link shim for OpenCV/Partio).
  2. `bash oracle/ref_harness/make_goldens.sh` runs it on the GPU box; dumps land in
     gpurun_out/ref_dumps/<case>/<tag>.<field>.f32 (raw little-endian float32).
  3. `python tests/golden/import_ref_dumps.py` (this script) selects tags and writes
     tests/golden/ref_<case>.npz with keys "<tag>.<field>" plus the scalar metadata.
Cases: t16 (nTheta=16, particleDensity=4, 3 steps), t32 (32, 4, 12 steps),
t64 (64, 1, 3 steps), t128 (128, 1, 100 steps); dt=0.005, radius=5, the reference's FBM
initial velocity, the synthetic density of SURVEY.md section 8d, rand()-seeded particles.
Tags: init, s<k>_adv / s<k>_geo / s<k>_proj = state after that phase of step k
(s<k>_proj is the state after step k).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "..", "gpurun_out", "ref_dumps")

SELECT = {
    "t16": None,  # everything
    "t32": ["init", "s1_adv", "s1_geo", "s1_proj", "s2_adv", "s2_geo", "s2_proj", "s10_proj", "s12_proj"],
    "t64": ["init", "s1_adv", "s1_geo", "s1_proj", "s3_proj"],
    "t128": ["init", "s1_adv", "s1_geo", "s1_proj", "s10_proj", "s100_proj"],
}


def main():
    for case, tags in SELECT.items():
        d = os.path.join(SRC, case)
        if not os.path.isdir(d):
            print("skip", case, "(no dump directory)")
            continue
        meta = dict(line.split() for line in open(os.path.join(d, "meta.txt")))
        out = {"meta.nTheta": np.int64(meta["nTheta"]), "meta.numParticles": np.int64(meta["numParticles"]),
               "meta.dt": np.float32(meta["dt"]), "meta.radius": np.float32(meta["radius"]),
               "meta.nSteps": np.int64(meta["nSteps"])}
        for fn in sorted(os.listdir(d)):
            if not fn.endswith(".f32"):
                continue
            tag, field, _ = fn.split(".")
            if tags is not None and tag not in tags:
                continue
            out["%s.%s" % (tag, field)] = np.fromfile(os.path.join(d, fn), dtype=np.float32)
        path = os.path.join(HERE, "ref_%s.npz" % case)
        np.savez_compressed(path, **out)
        print(case, len(out), "arrays ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    sys.exit(main())
