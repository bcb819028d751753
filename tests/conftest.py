"""pytest configuration: the `gpu` marker and shared helpers.

`-m "not gpu"` (run on the CPU-only build box): oracle vs the reference's golden vectors, host logic, C-ABI
surface. `-m gpu` (run on a B200): the parity tests proper, all through the C ABI of libkzp_b200.so.
Nothing in the GPU tests reads /root/reference: fixtures live in tests/golden/, bigger inputs are generated at
test time by oracle/kzp_port.c, and the reference itself is present as the prebuilt oracle/_ref/libkzp_ref.so.
"""
import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "slow: full BASELINE.json sizes")


@pytest.fixture(scope="session")
def oracle():
    import bn254

    return bn254


@pytest.fixture(scope="session")
def kzp():
    """The product package; building here is a no-op when the in-tree library is current."""
    import keyless_zk_proofs_b200 as k

    if not os.path.exists(k.LIB_PATH):
        k.build()
    k.lib()
    return k


@pytest.fixture(scope="session")
def gpu(kzp):
    n = kzp.device_count()
    assert n > 0, "GPU test selected but no CUDA device is visible (there is no CPU fallback to test instead)"
    return n


_REF_SRC = "/root/reference/rust-rapidsnark/rapidsnark/src"


@pytest.fixture(scope="session")
def _ref_lib():
    import refutil

    return refutil.load_ref()


@pytest.fixture
def ref(request, _ref_lib):
    """The reference prover itself (unmodified sources compiled by oracle/Makefile into oracle/_ref).
    A missing library is a FAILURE wherever it is supposed to exist: under the gpu marker (it travels prebuilt with
    the snapshot; the parity tests are void without it) and in the build container (where /root/reference is mounted
    and __graft_entry__.build() compiles it). Only a CPU checkout with neither may skip."""
    if _ref_lib is None:
        msg = "oracle/_ref/libkzp_ref.so is missing: run `make -C oracle ref` where /root/reference is mounted"
        if request.node.get_closest_marker("gpu") is not None or os.path.isdir(_REF_SRC):
            pytest.fail(msg)
        pytest.skip(msg)
    return _ref_lib


@pytest.fixture(scope="session")
def port():
    import refutil

    return refutil.load_port()


@pytest.fixture(scope="session")
def workdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("kzp"))
