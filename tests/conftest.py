import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices")
    # the product library is built in-tree; building is not using the oracle
    from tfmesos_b200 import build
    try:
        build.build()
    except Exception as exc:  # no nvcc: the prebuilt .so must already be there
        if not os.path.exists(build.OUT):
            raise RuntimeError("libpsx.so missing and cannot be built: %s" % exc)


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    n = _n_gpus()
    for item in items:
        if "multigpu" in item.keywords and n < 2:
            item.add_marker(pytest.mark.skip(reason="needs >= 2 GPUs, have %d" % n))
        elif "gpu" in item.keywords and n < 1:
            # a plain `pytest` on a CPU-only box: the GPU tests are skipped, not failed
            # (the product itself still has no CPU fallback -- tests/test_abi.py)
            item.add_marker(pytest.mark.skip(reason="needs a CUDA device"))
