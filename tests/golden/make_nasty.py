"""Golden fixture for query-file ingest: a BED file full of the lines parse_bed (src/igd_base.c:53-72) has an
opinion about, searched by the UNMODIFIED reference (oracle/_ref/libigdref.so: getOverlaps / getOverlaps_v) on the
`dense` golden database.  Run in the build container (needs /root/reference via oracle/_ref); writes
tests/golden/nasty/{q.bed.gz, expected.json}.

    python tests/golden/make_nasty.py
"""
import gzip
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from tests.golden_util import GoldenCase  # noqa: E402


def nasty_lines(rng) -> bytes:
    out = []
    w = out.append
    for _ in range(3000):                                   # plain lines so that every kind below has company
        c = rng.integers(1, 4)
        s = int(rng.integers(0, 250000))
        w(f"chr{c}\t{s}\t{s + int(rng.integers(1, 30000))}\t.\t{int(rng.integers(0, 1000))}\n".encode())
    special = [
        b"chr1\t 1000\t  3000\n",                           # atol skips leading blanks
        b"chr1\t+1000\t+3000\n", b"chr1\t-1000\t3000\n", b"chr1\t1000\t-3000\n",
        b"chr1\t1000x\t3000y\tjunk\n", b"chr1\t1e3\t3e3\n", b"chr1\t0x10\t5000\n",
        b"chr1\t\t3000\n", b"chr1\t1000\t\n", b"chr1\t1000\n", b"chr1\n", b"\n", b"\t\t\n",
        b"chr1\t1000\t3000\r\n", b"chr1\t1000\t3000\t.\t5\r\n", b"chr1\r\n",
        b"chr2\t99999999999\t100000000999\n",               # long -> int32 truncation
        b"chr2\t5\t4294967396\n",                           # 2^32 + 100 -> 100
        b"chr2\t5\t2147483648\n",                           # -> INT_MIN: rejected (en > 0)
        b"chr2\t5\t9223372036854775807\n", b"chr2\t5\t9223372036854775808\n",
        b"chr2\t5\t99999999999999999999999\n", b"chr2\t-99999999999999999999999\t300\n",
        b"chr2\t3000\t1000\n", b"chr2\t2000\t2000\n",       # qs >= qe
        b"chrZ\t10\t20\n", b"chr\t10\t20\n", b"ch1\t10\t20\n", b"Chr1\t10\t20\n", b" chr1\t10\t20\n",
        b"chr1 \t10\t20\n", b"chr11\t10\t20\n",
        b"chr" + b"q" * 36 + b"\t10\t20\n", b"chr" + b"q" * 37 + b"\t10\t20\n",
        b"chr3\t100\t200\t" + b"z" * 70000 + b"\n",         # longer than any line buffer
        b"chr3\t100\t2\x0000\n",                            # NUL inside the end field: the line ends there
        b"chr3\x00\t100\t200\n", b"chr3\t1\x0000\t200\n",
        b"chr3\t\v\f\r 77\t\v 9000\n",
        b"#chr1\t10\t20\n", b"track name=x\n", b"chr1  10  20\n",
    ]
    for sp in special:
        out.insert(int(rng.integers(0, len(out))), sp)
    out.append(b"chr1\t500\t90000")                          # no newline at the end of the file
    return b"".join(out)


def main():
    assert O.have_ref(), "oracle/_ref is missing: run `make -C oracle` where /root/reference exists"
    out = os.path.join(HERE, "nasty")
    os.makedirs(out, exist_ok=True)
    text = nasty_lines(np.random.default_rng(20261017))
    with tempfile.TemporaryDirectory() as tmp:
        case = GoldenCase("dense", tmp)
        qf = os.path.join(tmp, "nasty.bed")
        open(qf, "wb").write(text)
        ref = O.RefLib(case.igd)
        exp = {}
        for v in (0, 1, 500):
            hits, ret = ref.count_file(qf, v)
            exp["hits" if v == 0 else f"hits_v{v}"] = hits.tolist()
            exp["ret" if v == 0 else f"ret_v{v}"] = int(ret)
        ref.close()
    with gzip.GzipFile(os.path.join(out, "q.bed.gz"), "wb", mtime=0) as f:
        f.write(text)
    json.dump(exp, open(os.path.join(out, "expected.json"), "w"))
    print("wrote", out, "sum of hits", sum(exp["hits"]))


if __name__ == "__main__":
    main()
