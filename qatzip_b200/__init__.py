"""qatzip_b200 -- B200-native chunked deflate/LZ4 codec behind the qatzip.h C ABI.

The product is the shared library `qatzip_b200/libqatzip.so` (C ABI: include/qatzip.h,
include/qatzip_b200.h), built from `qatzip_b200/csrc/` by `make` with nvcc for sm_100a.
This Python package only locates and loads it; there is no Python or CPU implementation
of the codec here, and loading fails loudly if the library has not been built.
"""
import ctypes
import os

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libqatzip.so")


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (needs nvcc) first")
    return ctypes.CDLL(LIB_PATH)
