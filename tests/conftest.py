"""pytest configuration: registers the `gpu` marker and provides the two libraries under comparison.

  ezb   -- the product, libezb200.so (CUDA; every `-m gpu` test calls it through its C ABI)
  ezref -- the oracle, oracle/_ref/libezref.so = the reference's own src/interp/*.c compiled in place +
           oracle/fortran_shim.c (test infrastructure only; built by oracle/Makefile when /root/reference exists)
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libezref.so")
SHIM_SO = os.path.join(ROOT, "oracle", "libezshim.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_oracle():
    if not os.path.exists(REF_SO) and os.path.isdir("/root/reference/src/interp"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])
    return os.path.exists(REF_SO)


@pytest.fixture(scope="session")
def ezref():
    from librmn_b200 import EzLib
    if not _ensure_oracle():
        pytest.skip("oracle/_ref/libezref.so is not built and /root/reference is not available to build it")
    return EzLib(REF_SO)


@pytest.fixture(scope="session")
def ezb():
    import librmn_b200
    lib = librmn_b200.load(build_if_missing=True)
    return lib


@pytest.fixture(scope="session")
def cuda_ok():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return True


@pytest.fixture
def tiles_only(ezb):
    """Calls of 1-4 fields normally take the direct kernel (c_ezb200_set_direct_max_fields); tests that are about the tile
    kernels, or that compare calls of different field counts bit for bit, switch it off for their duration."""
    prev = ezb.set_direct_max_fields(0)
    yield
    ezb.set_direct_max_fields(prev)
