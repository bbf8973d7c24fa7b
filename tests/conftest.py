import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The shared library is a build artefact (git-ignored): build it once when the tree is fresh and nvcc is here.
    lib = os.path.join(ROOT, "bayesformer_b200", "libbayesformer_b200.so")
    if not os.path.exists(lib):
        import shutil
        import subprocess

        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            subprocess.run(["make", "-C", os.path.join(ROOT, "bayesformer_b200", "csrc"), "-j", str(os.cpu_count() or 4)],
                           check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
