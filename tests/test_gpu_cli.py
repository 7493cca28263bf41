"""Drop-in proof: the reference's own command line tool -- utils/qzip.c,
qzip_main.c, qzip_7z.c compiled UNCHANGED by oracle/Makefile `qzip` -- linked against
qatzip_b200/libqatzip.so, compresses files that gzip(1) / the oracle decode, and restores them."""
import gzip
import os
import subprocess

import pytest

from harness import qzapi as q
from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
QZIP = os.path.join(ROOT, "oracle", "_ref", "qzip_b200")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_gpu(), reason="needs a CUDA device"),
              pytest.mark.skipif(not os.path.exists(QZIP), reason="reference CLI not built (needs /root/reference at build time)")]


def run(*args, cwd):
    r = subprocess.run([QZIP, *args], cwd=cwd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (args, r.stdout[-500:], r.stderr[-500:])
    return r


def test_cli_imports_only_public_symbols():
    out = subprocess.run(["nm", "-D", "--undefined-only", QZIP], capture_output=True, text=True).stdout
    used = sorted(l.split()[-1] for l in out.splitlines() if " qz" in l or "logMessage" in l)
    assert "qzCompress" in used and "qzDecompress" in used and "logMessage" in used
    lib = q.QzLib(q.PRODUCT_SO).lib
    for sym in used:
        assert hasattr(lib, sym), sym


@pytest.mark.parametrize("fmt,ext", [("gzipext", ".gz"), ("gzip", ".gz"), ("lz4", ".lz4")])
def test_cli_roundtrip(tmp_path, corpus, port, fmt, ext):
    data = corpus.make(q.Corpus.SILESIA_LIKE, 5 << 20, first_seg=3)[: (5 << 20) - 12345]
    src = tmp_path / "sample.bin"
    src.write_bytes(data)
    # the CLI copies -A into comp_algorithm and qzSetupSessionLZ4 insists on QZ_LZ4 (reference utils/qzip.c:471)
    run("-k", *(("-A", "lz4") if fmt == "lz4" else ()), "-O", fmt, str(src), cwd=tmp_path)
    comp = tmp_path / ("sample.bin" + ext)
    blob = comp.read_bytes()
    assert 0 < len(blob) < len(data)
    if fmt != "lz4":
        assert gzip.decompress(blob) == data                       # independent RFC 1952 decoder
    else:
        assert port.decompress(blob, q.FMT_LZ4, len(data) + 8) == data
    src.unlink()
    run("-d", "-k", *(("-A", "lz4") if fmt == "lz4" else ()), str(comp), cwd=tmp_path)
    assert src.read_bytes() == data


def test_cli_chunk_size_and_block_size(tmp_path, corpus):
    data = corpus.make(q.Corpus.SILESIA_LIKE, 3 << 20, first_seg=7)
    src = tmp_path / "b.bin"
    src.write_bytes(data)
    run("-k", "-C", "16384", "-b", "1048576", str(src), cwd=tmp_path)
    assert gzip.decompress((tmp_path / "b.bin.gz").read_bytes()) == data
