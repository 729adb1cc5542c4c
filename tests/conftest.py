"""Shared fixtures.  `-m "not gpu"` runs here (CPU only); `-m gpu` runs on a B200 box."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "convex-2d-gpu-collision-detection_b200"


def pytest_report_header(config):
    try:
        mod = importlib.import_module(PKG)
        return f"libsatmc: {mod.load_library().satmc_version().decode()}  [{mod.LIB_PATH}]"
    except Exception as e:                                    # the ABI tests report a missing library properly
        return f"libsatmc: not loadable ({e})"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def satmc():
    """The product package (ctypes host mirror of include/satmc.h)."""
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def workloads(satmc):
    return importlib.import_module(PKG + ".workloads")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (C restatement), built on demand with gcc."""
    from oracle.binding import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def refgpu():
    """The unmodified reference compiled for sm_100a (oracle/_ref); GPU tests only."""
    from oracle.binding import RefGpu
    return RefGpu()


@pytest.fixture(scope="session")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    return torch


@pytest.fixture(scope="session")
def ctx(satmc, torch_cuda):
    c = satmc.Context(0)
    yield c
    c.close()


class Dev:
    """numpy <-> device helpers over torch (device memory only; no torch compute on the path)."""

    def __init__(self, torch):
        self.torch = torch

    def put(self, a: np.ndarray):
        a = np.ascontiguousarray(a)
        if a.dtype.fields is not None:                       # structured satmc_pair array -> raw float32 view
            a = a.view(np.float32)
        if a.dtype == np.uint64:
            return self.torch.from_numpy(a.view(np.int64)).cuda()
        if a.dtype == np.uint32:
            return self.torch.from_numpy(a.view(np.int32)).cuda()
        return self.torch.from_numpy(a).cuda()

    def zeros(self, n, dtype):
        t = {np.uint64: self.torch.int64, np.uint8: self.torch.uint8, np.float32: self.torch.float32,
             np.int32: self.torch.int32, np.uint32: self.torch.int32}[dtype]
        return self.torch.zeros(n, dtype=t, device="cuda")

    def get(self, t, dtype=None):
        a = t.cpu().numpy()
        if dtype is not None:
            a = a.view(dtype)
        return a


@pytest.fixture(scope="session")
def dev(torch_cuda):
    return Dev(torch_cuda)
