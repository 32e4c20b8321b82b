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


def test_batch_struct_layout_matches_header():
    """struct psx_op in include/psx.h <-> ctypes mirror: same field order and size."""
    text = open(os.path.join(ROOT, "include", "psx.h")).read()
    body = re.search(r"typedef struct psx_op \{(.*?)\} psx_op;", text, re.S).group(1)
    names = re.findall(r"\b(?:\*?)(\w+)(?:,|;)", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    names = [n for n in names if n not in ("int32_t", "uint64_t", "uint32_t", "void")]
    assert names == [f[0] for f in psx.Op._fields_]
    assert ctypes.sizeof(psx.Op) == 16 + 24 + 16 + 8          # 4 x i32, 3 x u64, 2 ptr, 2 x u32
    consts = dict(re.findall(r"#define (PSX_OP_[A-Z_]+) (\d+)", text))
    for name, val in consts.items():
        assert getattr(psx, name[4:]) == int(val), name


def test_batch_reports_the_failing_op_without_touching_cuda():
    b = psx.Batch([dict(op=psx.OP_WAIT_APPLIED, id=12345, uses_seq=False, stream=0),
                   dict(op=99, stream=0)])
    with pytest.raises(RuntimeError, match="batch op 0 failed"):
        b.run(1)
    b2 = psx.Batch([dict(op=99, stream=0)])
    with pytest.raises(RuntimeError, match="unknown opcode 99"):
        b2.run(1)


def test_argument_validation_that_needs_no_device():
    lib = psx.lib()
    ids = (ctypes.c_uint64 * 1)(7)
    assert lib.psx_signal_many(ids, 0, 1, None) == -1                  # PSX_EINVAL: n out of range
    assert lib.psx_signal_many(ids, 1, 1, None) == -1                  # unknown client id
    assert b"unknown client id" in lib.psx_last_error()
    assert lib.psx_wait_mailbox(424242, 1, None) == -1
    assert lib.psx_shard_destroy(424242) == -1
    out = ctypes.c_uint64(0)
    assert lib.psx_list_create(424242, None, None, None, 1, ctypes.byref(out)) == -1
