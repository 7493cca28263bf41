"""Generates tests/golden/*: small input files and the streams the UNMODIFIED reference
(oracle/_ref, compiled from /root/reference by oracle/Makefile) produces for them, plus CRC-32
(zlib) and XXH32 (the reference's vendored xxhash, through its LZ4 footer) of each input.
Run in the build container:  python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from harness import qzapi as q  # noqa: E402

ref, cor = q.QzLib(q.REF_SO), q.Corpus()
cases = []
inputs = {
    "text_20k.bin": cor.make(q.Corpus.SILESIA_LIKE, 1 << 20, first_seg=0)[:20000],
    "records_9k.bin": cor.make(q.Corpus.SILESIA_LIKE, 1 << 20, first_seg=2)[:9000],
    "rle_12k.bin": cor.make(q.Corpus.REF_RLE, 1 << 20)[:12000],
    "one.bin": b"Q",
}
for name, data in inputs.items():
    open(os.path.join(HERE, name), "wb").write(data)
    lz = ref.compress(data, fmt=q.FMT_LZ4, hw_buff_sz=4096)
    xxh = int.from_bytes(lz[-4:], "little")          # content checksum written by liblz4 via the reference
    for fmt in (q.QZ_DEFLATE_4B, q.QZ_DEFLATE_GZIP, q.QZ_DEFLATE_GZIP_EXT, q.QZ_DEFLATE_RAW, q.FMT_LZ4, q.FMT_ZLIB):
        hw = 4096
        blob = ref.compress(data, fmt=fmt, hw_buff_sz=hw)
        sname = f"{name[:-4]}.{q.FMT_NAMES[fmt].lower()}"
        open(os.path.join(HERE, sname), "wb").write(blob)
        cases.append({"input": name, "stream": sname, "fmt": fmt, "hw_buff_sz": hw,
                      "input_sha256": hashlib.sha256(data).hexdigest(), "crc32": zlib.crc32(data), "xxh32": xxh, "adler32": zlib.adler32(data)})
json.dump({"generator": "tests/golden/make_golden.py", "reference": "intel/QATzip 1.3.1 software path (zlib 1.3, liblz4 1.9.4)",
           "cases": cases}, open(os.path.join(HERE, "manifest.json"), "w"), indent=1)
print(len(cases), "cases")
