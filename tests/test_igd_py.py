"""igd_py: the reference's Python module (src_py/igd_py.pyx) under its own name over the C-ABI -- a script
written against the reference, like src_py/igd_test.py, runs unchanged: igd.create(src, out, name, 16384);
igd.open(path); hits = np.zeros(igd.get_nFiles(), 'int64'); total = igd.search_n(qfile, hits)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from igd_b200 import api, synth
from oracle import oracle as O
from tests.conftest import SMALL_SIZES
from tests.golden_util import GoldenCase

BIND = os.path.join(os.path.dirname(api.LIB_PATH), "..", "bindings")


@pytest.fixture(scope="module")
def igd_py():
    r = subprocess.run(["make", "-C", BIND], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("igd_py did not build here: " + r.stderr[-300:])
    sys.path.insert(0, os.path.abspath(BIND))
    import igd_py as mod
    return mod


def test_create_and_open_error(igd_py, workdir):
    bed = os.path.join(workdir, "igdpy_bed")
    os.makedirs(bed)
    names = ["chr1", "chr2", "chr3"]
    for i in range(10):
        synth.write_bed(os.path.join(bed, f"f{i:05d}.bed"),
                        synth.gen_intervals(97000 + i, 500, ("lognormal", 6.5, 1.3), SMALL_SIZES), names)
    out = os.path.join(workdir, "igdpy_out")
    g = igd_py.igd_py()
    assert g.get_nFiles() == 0
    g.create(bed, out, "db", 16384)                     # the reference's call, src_py/igd_test.py:18
    db = O.OracleDB(os.path.join(out, "db.igd"))
    assert db.nfiles == 10 and db.nbp == 16384
    with pytest.raises(RuntimeError):
        g.open(os.path.join(out, "missing.igd"))
    if api.load_library().igd_b200_device_count() <= 0:
        with pytest.raises(RuntimeError, match="no CUDA device"):
            g.open(os.path.join(out, "db.igd"))
    with pytest.raises(RuntimeError, match="no database is open"):
        g.search_n("x.bed", np.zeros(10, np.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["quirk", "dense", "cfg1"])
def test_search_like_the_reference_script(igd_py, name, workdir):
    case = GoldenCase(name, workdir)
    igd = igd_py.igd_py()
    igd.open(case.igd)
    n = igd.get_nFiles()
    hits = np.zeros(n, dtype="int64")
    total = igd.search_n(case.qfile, hits)
    assert np.array_equal(hits, case.hits(0)) and total == int(case.hits(0).sum())
    total2 = igd.search_n(case.qfile, hits)             # hits[] is incremented; the return value is its sum
    assert np.array_equal(hits, 2 * case.hits(0)) and total2 == 2 * total
    for r in case.exp["regions"]:
        if r["v"] == 0:
            h1 = np.zeros(n, dtype="int64")
            igd.search_1(r["chrm"], r["qs"], r["qe"], h1)
            assert h1.tolist() == r["hits"]
    with pytest.raises(ValueError):
        igd.search_n(case.qfile, np.zeros(max(n - 1, 0), dtype="int64"))
