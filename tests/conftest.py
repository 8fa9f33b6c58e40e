import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def host_shim():
    """Host build (g++) of the product's device arithmetic headers, for CPU-side checks."""
    import ctypes as C
    import subprocess
    d = os.path.join(ROOT, "tests", "host_shim")
    so = os.path.join(d, "libhost_shim.so")
    src = os.path.join(d, "host_math_shim.cpp")
    deps = [src] + [os.path.join(ROOT, "dsac_b200", "csrc", f) for f in ("pose_math.cuh", "sampler.cuh", "lm_math.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    return C.CDLL(so)


@pytest.fixture(scope="session")
def engine_mod():
    from dsac_b200 import engine as E
    E.load()
    return E


GOLDEN = os.path.join(ROOT, "tests", "golden")
