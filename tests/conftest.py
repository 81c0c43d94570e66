import os
import sys

import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "webrtc-audio-processing_b200", "python"),
          os.path.join(ROOT, "tests"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The compiled, unmodified reference (oracle/_ref/libwap_ref.so)."""
    import build_ref
    build_ref.build(verbose=False)
    import ref
    ref.lib()
    return ref


@pytest.fixture(scope="session")
def emu_lib():
    """Kernel sources compiled for the CPU warp emulator (test infrastructure)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    import wap_b200
    return wap_b200.load(build_emu.build(verbose=False))


@pytest.fixture(scope="session")
def gpu_lib():
    """The product: libwap_b200.so (sm_100a).  No fallback: missing => error."""
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import wap_b200
    return wap_b200.load()
