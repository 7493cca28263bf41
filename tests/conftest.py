"""pytest configuration: markers, import path, and one-time builds of the checkers.

`-m "not gpu"` tests need: oracle/liboracle_port.so, harness/libqzcorpus.so, qatzip_b200/libqatzip.so
(loaded and symbol-checked only) and, when /root/reference exists, oracle/_ref.
`-m gpu` tests call through the C ABI of qatzip_b200/libqatzip.so on a real B200.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _run(cmd, cwd):
    subprocess.run(cmd, cwd=cwd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)


@pytest.fixture(scope="session", autouse=True)
def built():
    import __graft_entry__ as ge
    ge.build_checkers()
    if not os.path.exists(os.path.join(ROOT, "qatzip_b200", "libqatzip.so")):
        ge.build()
    return True


@pytest.fixture(scope="session")
def corpus(built):
    from harness.qzapi import Corpus
    return Corpus()


@pytest.fixture(scope="session")
def port(built):
    from harness.qzapi import OraclePort
    return OraclePort()


@pytest.fixture(scope="session")
def ref(built):
    from harness.qzapi import QzLib, REF_SO
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return QzLib(REF_SO)


@pytest.fixture(scope="session")
def prod(built):
    from harness.qzapi import QzLib, PRODUCT_SO
    return QzLib(PRODUCT_SO)


def has_gpu():
    try:
        import ctypes
        from harness.qzapi import PRODUCT_SO
        lib = ctypes.CDLL(PRODUCT_SO)
        return lib.qzb200DeviceCount() > 0
    except Exception:
        return False
