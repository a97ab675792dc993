import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
    # the native pieces are built in-tree and git-ignored: (re)build them when a source is newer
    # (`make` is a no-op otherwise; nvcc cross-compiles sm_100a without a GPU)
    import subprocess
    for sub in (os.path.join("panfeed_b200", "csrc"), "oracle"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, sub)], check=False,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
