import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def _oracle_path(omp: bool = False) -> str:
    name = "libsaccot_oracle_omp.so" if omp else "libsaccot_oracle.so"
    path = os.path.join(ROOT, "oracle", name)
    src = os.path.join(ROOT, "oracle", "sac_cot_oracle.cpp")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), name], check=True,
                       stdout=subprocess.DEVNULL)
    return path


@pytest.fixture(scope="session")
def oracle_lib():
    """The from-paper CPU oracle (test infrastructure), bound through the shared ctypes ABI."""
    from sac_cot_b200 import _abi
    return _abi.bind(ctypes.CDLL(_oracle_path()))


@pytest.fixture(scope="session")
def oracle_omp_lib():
    from sac_cot_b200 import _abi
    return _abi.bind(ctypes.CDLL(_oracle_path(omp=True)))


@pytest.fixture()
def oracle(oracle_lib):
    from sac_cot_b200.api import Registrar
    reg = Registrar(lib=oracle_lib)
    reg.set("keep_debug", 1)
    yield reg
    reg.close()


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA product library; GPU tests fail (not skip) if it is missing."""
    from sac_cot_b200.api import load_library
    return load_library()


@pytest.fixture()
def gpu(gpu_lib):
    from sac_cot_b200.api import Registrar
    reg = Registrar(lib=gpu_lib, device=0)
    reg.set("keep_debug", 1)
    # stage-parity tests compare every edge key and the full histogram: the POPC kernels provide them; the
    # tensor-core kernel (which prunes keys) and the automatic choice have their own tests
    reg.set("triangle_path", 0)
    yield reg
    reg.close()
