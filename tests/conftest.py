import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Make sure the product library and the oracle exist (built in-tree, never JIT-cached)."""
    lib = os.path.join(ROOT, "kaminogpu_b200", "libkamino_b200.so")
    exe = os.path.join(ROOT, "kaminogpu_b200", "kamino")
    import subprocess
    if not (os.path.exists(lib) and os.path.exists(exe)):
        import __graft_entry__
        __graft_entry__.build()
    else:       # no-op when up to date
        subprocess.run(["make", "-C", os.path.join(ROOT, "kaminogpu_b200"), "all"], check=True,
                       stdout=subprocess.DEVNULL)
    import oracle_api
    oracle_api.build_oracle()
    return lib
