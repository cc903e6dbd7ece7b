"""The C-ABI shared library: loads, exports every symbol include/igd_b200.h declares, and its host-only entry
points (query ingest) agree with the reference's parse_bed rule.  No GPU needed; no compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from igd_b200 import api
from oracle import oracle as O
from tests.golden_util import GoldenCase


def declared_functions():
    text = open(api.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(api.LIB_PATH)
    names = declared_functions()
    assert {"igd_b200_open", "get_overlaps_n", "search", "get_overlaps", "getOverlaps", "getOverlaps_v",
            "get_overlaps_v", "getMap", "getMap_v", "get_overlaps_f1", "getOverlaps_f1", "seqOverlaps",
            "igd_search", "igd_b200_count_device", "igd_b200_full", "igd_b200_hitsmap"} <= set(names)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_cli_binary_exists_and_prints_usage():
    import subprocess
    r = subprocess.run([api.CLI_PATH, "search"], capture_output=True, text=True)
    assert r.returncode == 0 and "usage:" in r.stderr and "-q <query file>" in r.stderr
    r = subprocess.run([api.CLI_PATH, "search", "nope.txt", "-q", "x"], capture_output=True, text=True)
    assert r.stdout == "nope.txt is not an igd database"
    r = subprocess.run([api.CLI_PATH, "search", "/nonexistent/db.igd", "-q", "x"], capture_output=True, text=True)
    assert r.stdout == "/nonexistent/db.igd does not exist"


def test_open_fails_loudly_without_gpu_or_file(workdir):
    L = api.load_library()
    with pytest.raises(api.IGDError):
        api.IGD(os.path.join(workdir, "missing.igd"))
    if L.igd_b200_device_count() <= 0:
        case = GoldenCase("quirk", workdir)
        with pytest.raises(api.IGDError, match="no CUDA device"):
            api.IGD(case.igd)


@pytest.mark.parametrize("name", ["quirk", "dense", "cfg1"])
def test_query_ingest_matches_reference_rule(name, workdir):
    """igd_b200_read_bed accepts exactly the lines the reference's parse_bed accepts (the oracle's reader is
    pinned by test_oracle_golden; hits of the golden cases depend on it line for line)."""
    case = GoldenCase(name, workdir)
    mine, ref = api.QuerySet(case.qfile), O.Queries(case.qfile)
    assert mine.n == ref.n and mine.names == ref.names
    assert np.array_equal(mine.qs, ref.qs) and np.array_equal(mine.qe, ref.qe)
    assert np.array_equal(mine.name_id, ref.name_id)


def test_query_ingest_gz_and_long_lines(workdir):
    import gzip
    p = os.path.join(workdir, "long.bed.gz")
    with gzip.open(p, "wt") as f:
        f.write("chr1\t10\t20\t" + "x" * 200000 + "\n")
        f.write("chr" + "y" * 37 + "\t1\t2\n")          # 40 chars: rejected (strlen < 40)
        f.write("chr" + "y" * 36 + "\t1\t2\n")          # 39 chars: accepted
        f.write("\n\nchr2\t 7\t9 \n")
    mine, ref = api.QuerySet(p), O.Queries(p)
    assert mine.n == ref.n == 3
    assert mine.qs.tolist() == ref.qs.tolist() == [10, 1, 7]
    assert mine.names == ref.names
