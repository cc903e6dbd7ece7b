"""The host C of the library (igd_host.c: query ingest, .igd / _index.tsv parsing, error paths) built with
-fsanitize=address,undefined against a stub device layer (tests/csrc/stub_dev.c) -- SURVEY.md §5's sanitizer
build of the host C; runs without a GPU.  Checks: no sanitizer report, the reader accepts exactly the lines the
oracle's (reference-pinned) reader accepts, truncated and corrupt databases fail with a message instead of a crash."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

from igd_b200 import api, synth
from oracle import oracle as O
from tests.conftest import SMALL_NAMES, SMALL_SIZES
from tests.golden_util import GOLDEN, GoldenCase

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENV = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:halt_on_error=1", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")


@pytest.fixture(scope="module")
def probe(workdir):
    exe = os.path.join(workdir, "host_probe_asan")
    cc = shutil.which("gcc") or "gcc"
    r = subprocess.run([cc, "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                        "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "igd_b200", "csrc"), "-o", exe,
                        os.path.join(ROOT, "tests", "csrc", "host_probe.c"), os.path.join(ROOT, "tests", "csrc", "stub_dev.c"),
                        os.path.join(ROOT, "igd_b200", "csrc", "igd_host.c"), "-lz", "-lpthread"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        # a missing sanitizer runtime is a reason to skip; a stub that no longer matches igd_dev.h is a failure
        assert "sanitize" in r.stderr or "asan" in r.stderr.lower() or "cannot find" in r.stderr, r.stderr[-1500:]
        pytest.skip("no sanitizer runtime here: " + r.stderr[-300:])
    return exe


def run(probe, *args):
    r = subprocess.run([probe, *args], capture_output=True, env=ENV)
    assert r.returncode == 0 and b"ERROR" not in r.stderr and b"runtime error" not in r.stderr, r.stderr.decode()[-3000:]
    return r.stdout.decode(errors="replace")


def checksum(a):
    h = 0
    for x in np.asarray(a, np.int64).tolist():
        h = (h * 1000003 + (x & 0xffffffff)) & 0xffffffffffffffff
    return h


def test_reader_under_sanitizers_matches_oracle_reader(probe, workdir):
    files = []
    nasty_gz = os.path.join(GOLDEN, "nasty", "q.bed.gz")
    plain = os.path.join(workdir, "nasty_plain_asan.bed")
    with gzip.open(nasty_gz, "rb") as g, open(plain, "wb") as f:
        f.write(g.read())
    files += [nasty_gz, plain]
    for name in ("quirk", "dense", "dense0", "cfg1"):
        files.append(GoldenCase(name, workdir).qfile)
    long_lines = os.path.join(workdir, "long_asan.bed")
    with open(long_lines, "w") as f:
        f.write("chr1\t10\t20\t" + "x" * 300000 + "\n" + "chr" + "y" * 36 + "\t1\t2\n\n\nchr2\t 7\t9 \r\nchr3\t5")
    files.append(long_lines)
    out = run(probe, "bed", *files, os.path.join(workdir, "does_not_exist.bed"))
    rows = [ln.split("\t") for ln in out.strip().split("\n")]
    assert rows[-1][1] == "unreadable"
    for path, row in zip(files, rows):
        q = O.Queries(path)
        assert row[0] == path and int(row[1]) == q.n and int(row[2]) == len(q.names), path
        assert (int(row[3]), int(row[4]), int(row[5])) == (checksum(q.qs), checksum(q.qe), checksum(q.name_id)), path


def test_open_error_paths_under_sanitizers(probe, workdir):
    stem = os.path.join(workdir, "asan_db")
    files = [synth.gen_intervals(93000 + i, 300, ("lognormal", 6.0, 1.2), SMALL_SIZES) for i in range(10)]
    synth.write_igd_direct(stem, files, names=SMALL_NAMES)
    raw = open(stem + ".igd", "rb").read()
    tsv = open(stem + "_index.tsv").read()

    def variant(tag, data, with_tsv=True):
        p = os.path.join(workdir, f"asan_db_{tag}.igd")
        open(p, "wb").write(data)
        if with_tsv:
            open(p[:-4] + "_index.tsv", "w").write(tsv)
        return p

    hdr = np.frombuffer(raw[:12], np.int32).copy()
    cases = {
        "whole": (variant("whole", raw), "no CUDA device"),                 # parses, then there is no device
        "gtype0": (variant("g0", np.array([hdr[0], 0, hdr[2]], np.int32).tobytes() + raw[12:]), "no CUDA device"),
        "cut_header": (variant("cut1", raw[:10]), "truncated header"),
        "cut_tables": (variant("cut2", raw[:12 + 4 * int(hdr[2]) + 20]), "truncated header"),
        "bad_gtype": (variant("gt", np.array([hdr[0], 7, hdr[2]], np.int32).tobytes() + raw[12:]), "unknown gType"),
        "bad_nbp": (variant("nbp", np.array([0, 1, hdr[2]], np.int32).tobytes() + raw[12:]), "corrupt header"),
        "neg_tiles": (variant("neg", raw[:12] + np.array([-5], np.int32).tobytes() + raw[16:]), "corrupt header"),
        "no_tsv": (variant("notsv", raw, with_tsv=False), "file not found"),
        "missing": (os.path.join(workdir, "nope.igd"), "Can't open file"),
    }
    for tag, (path, msg) in cases.items():
        out = run(probe, "open", path)
        state, err = out.rstrip("\n").split("\t", 1)
        assert state == "failed" and msg in err, (tag, out)
