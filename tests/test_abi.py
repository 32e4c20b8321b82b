"""The C-ABI library loads on a box without a GPU and exports every symbol that
include/psx.h declares; compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from tfmesos_b200 import psx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "psx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(psx_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(psx.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(psx.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_abi_version_and_constants_match_header():
    text = open(os.path.join(ROOT, "include", "psx.h")).read()
    consts = dict(re.findall(r"#define (PSX_[A-Z0-9_]+) \(?(-?\d+)\)?", text))
    assert psx.lib().psx_abi_version() == int(consts["PSX_ABI_VERSION"]) == psx.ABI_VERSION
    assert int(consts["PSX_HANDLE_BYTES"]) == psx.HANDLE_BYTES
    assert int(consts["PSX_MAX_SLOTS"]) == psx.MAX_SLOTS
    assert (int(consts["PSX_OPT_SGD"]), int(consts["PSX_OPT_ADAM"])) == (psx.OPT_SGD, psx.OPT_ADAM)
    assert (int(consts["PSX_MODE_ASYNC_ORDERED"]), int(consts["PSX_MODE_SUM"]),
            int(consts["PSX_MODE_SYNC_MEAN"])) == (psx.MODE_ASYNC_ORDERED, psx.MODE_SUM, psx.MODE_SYNC_MEAN)
    assert (int(consts["PSX_F32"]), int(consts["PSX_BF16"])) == (psx.F32, psx.BF16)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tfmesos_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, fn
                assert "ps_oracle" not in src, fn


def _has_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(_has_gpu(), reason="checks behaviour on a box WITHOUT a GPU")
def test_compute_calls_fail_loudly_without_a_gpu():
    with pytest.raises(RuntimeError):
        psx.init(0)
    with pytest.raises(RuntimeError):
        psx.Shard(0, 1024)
    assert psx.last_error() != ""


def test_bad_handles_are_rejected():
    with pytest.raises(RuntimeError, match="not a psx handle"):
        psx.Client(b"\0" * psx.HANDLE_BYTES, 0, 0)
