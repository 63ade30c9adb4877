"""The C ABI without a GPU: the library loads, exports every symbol include/csa_gpu.h declares,
and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from common import ROOT

LIB = os.path.join(ROOT, "csa_b200", "csrc", "libcsa_gpu.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", os.path.dirname(LIB)], stdout=subprocess.DEVNULL)
    return C.CDLL(LIB)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "csa_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(csa_gpu_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/csa_gpu.h but not exported"
    assert lib.csa_gpu_abi_version() == 1


def test_no_device_no_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = C.c_void_p()
    lib.csa_gpu_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    rc = lib.csa_gpu_create(0, C.byref(ctx))
    assert rc == -1 and not ctx.value  # CSA_GPU_ENODEV
    lib.csa_gpu_last_error.restype = C.c_char_p
    assert b"no CPU fallback" in lib.csa_gpu_last_error()


def test_product_does_not_reference_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's cpu legs may touch oracle/"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "csa_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "csa_oracle" not in text and "oracle/" not in text, f"{f} mentions the oracle"
