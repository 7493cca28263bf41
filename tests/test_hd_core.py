"""The scalar host+device code of the codec (Huffman construction, dynamic header coding,
inflate core, symbol arithmetic, CRC-32 combine, xxHash32) compiled for the CPU and run against
zlib: the CUDA kernels execute these same functions."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hd_selftest(tmp_path):
    exe = str(tmp_path / "hd_selftest")
    subprocess.run(["g++", "-O2", "-Wall", "-Wno-unused-function", "-o", exe,
                    os.path.join(ROOT, "tests", "cpu", "hd_selftest.cpp"), "-lz"], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "hd_selftest ok" in out
