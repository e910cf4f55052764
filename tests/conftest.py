import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

PACKAGE = "massively-parallel-gpu-ode-solver_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (its directory name is not a Python identifier)."""
    return importlib.import_module(PACKAGE)


@pytest.fixture(scope="session")
def workloads(pkg):
    return pkg.workloads


@pytest.fixture(scope="session")
def oracle():
    """The checker (CPU restatement + host-compiled reference); built on demand."""
    import subprocess
    odir = os.path.join(ROOT, "oracle")
    if not os.path.exists(os.path.join(odir, "liboracle_perthread.so")):
        subprocess.run(["make", "-C", odir, "oracle"], check=True, capture_output=True)
    import oracle_py
    return oracle_py
