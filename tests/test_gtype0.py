"""gType-0 databases (`igd create -s 0`: 12-byte gdata0_t records without a value, src/igd_base.h:48-52,
src/igd_create.c:245-332; searched by get_overlaps0 / getOverlaps0 / get_overlaps_f0, src/igd_search.c:30-250).

tests/golden/dense0 was built AND searched by the unmodified reference (tests/golden/make_golden.py dense0).
CPU tests pin the oracle's gType-0 reader to it; GPU tests hold the CUDA path (C-ABI and CLI) to it."""
import os
import subprocess

import numpy as np
import pytest

from igd_b200 import synth
from oracle import oracle as O
from tests.conftest import SMALL_NAMES, SMALL_SIZES, make_queries
from tests.golden_util import GoldenCase


@pytest.fixture(scope="module")
def case(workdir):
    return GoldenCase("dense0", workdir)


@pytest.fixture(scope="module")
def twin_dbs(workdir):
    """the same records written as a gType-1 and as a gType-0 database"""
    files = [synth.gen_intervals(31000 + i, 2500, ("lognormal", 6.0, 1.3), SMALL_SIZES) for i in range(11)]
    s1, s0 = os.path.join(workdir, "twin1"), os.path.join(workdir, "twin0")
    synth.write_igd_direct(s1, files, names=SMALL_NAMES)
    synth.write_igd_direct(s0, files, names=SMALL_NAMES, gtype=0)
    return s1 + ".igd", s0 + ".igd"


# ---- CPU: the oracle's gType-0 reader against the reference ----------------------------------------------
def test_oracle_reads_gtype0_golden(case):
    db = O.OracleDB(case.igd)
    assert db.gtype == 0 and case.exp["gtype"] == 0
    q = O.Queries(case.qfile)
    ctg = db.resolve(q)
    for v in (0, 500):                      # -v is not looked at on a gType-0 database (igd_search.c:1024)
        hits, ret = db.count(ctg, q.qs, q.qe, v)
        assert np.array_equal(hits, case.hits(0)) and ret == case.ret(0) == 0
    for r in case.exp["regions"]:
        hits, ret = db.count([db.ctg_id(r["chrm"])], [r["qs"]], [r["qe"]], 0)
        assert hits.tolist() == r["hits"] and ret == r["ret"]


def test_oracle_full_order_gtype0(case):
    db = O.OracleDB(case.igd)
    q = O.Queries(os.path.join(case.dir, "qf.bed"))
    per, oq, s, e, f = db.full(db.resolve(q), q.qs, q.qe)
    blocks, total = O.parse_full_stdout(case.stdout("f"))
    assert total == per.sum()
    want = [(s_, e_, fn) for (_, _, _, L) in blocks for (s_, e_, fn) in L]
    assert [(int(a), int(b), db.file_names[c]) for a, b, c in zip(s, e, f)] == want


def test_writer_gtype0_layout(twin_dbs):
    """the direct writer's gType-0 file = the gType-1 file minus the value column"""
    p1, p0 = twin_dbs
    d1, d0 = O.OracleDB(p1), O.OracleDB(p0)
    assert (d1.gtype, d0.gtype) == (1, 0) and d1.nrecords == d0.nrecords and d1.ntile == d0.ntile
    hdr = 12 + 44 * d1.nctg + 4 * sum(d1.ntile)
    assert os.path.getsize(p1) == hdr + 16 * d1.nrecords and os.path.getsize(p0) == hdr + 12 * d0.nrecords
    a = np.fromfile(p1, np.int32, offset=hdr).reshape(-1, 4)
    b = np.fromfile(p0, np.int32, offset=hdr).reshape(-1, 3)
    assert np.array_equal(a[:, :3], b)
    q = make_queries(5, 3000, ("lognormal", 7.0, 1.5))
    assert np.array_equal(d1.count(q.ctg, q.start, q.end)[0], d0.count(q.ctg, q.start, q.end)[0])


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_writer_gtype0_matches_reference_create(workdir):
    """`igd create -s 0` by the unmodified reference and the direct writer give the same file"""
    bed = os.path.join(workdir, "g0bed")
    os.makedirs(bed)
    names = ["chr1", "chr2", "chr3"]
    files = [synth.gen_intervals(32000 + i, 800, ("lognormal", 6.5, 1.3), SMALL_SIZES) for i in range(10)]
    for i, iv in enumerate(files):
        synth.write_bed(os.path.join(bed, f"f{i:05d}.bed"), iv, names)
    ref = O.ref_create(bed, os.path.join(workdir, "g0ref"), "db", gtype=0)
    synth.write_igd_direct(os.path.join(workdir, "g0mine"), files, names=names, gtype=0)
    a, b = open(ref, "rb").read(), open(os.path.join(workdir, "g0mine.igd"), "rb").read()
    assert len(a) == len(b)
    da, db = O.OracleDB(ref), O.OracleDB(os.path.join(workdir, "g0mine.igd"))
    q = make_queries(6, 2000, ("lognormal", 7.0, 1.5))
    assert np.array_equal(da.count(q.ctg, q.start, q.end)[0], db.count(q.ctg, q.start, q.end)[0])


# ---- GPU: the CUDA path against the reference's outputs ----------------------------------------------------
@pytest.fixture(scope="module")
def api():
    from igd_b200 import api
    return api


@pytest.mark.gpu
def test_gpu_count_gtype0_golden(case, api):
    L = api.load_library()
    with api.IGD(case.igd) as g:
        assert g.gtype == 0
        for v in (0, 500):
            hits, n = g.search_n(case.qfile, v=v)
            assert np.array_equal(hits, case.hits(0)) and n == int(case.hits(0).sum())
        q = api.QuerySet(case.qfile)
        hits, _ = g.count(g.resolve(q), q.qs, q.qe, v=500)
        assert np.array_equal(hits, case.hits(0))
        g.use()
        hits = np.zeros(g.nfiles, np.int64)
        assert L.getOverlaps0(os.fsencode(case.qfile), hits) == 0
        assert np.array_equal(hits, case.hits(0))
        for r in case.exp["regions"]:
            hits = np.zeros(g.nfiles, np.int64)
            assert L.get_overlaps0(r["chrm"].encode(), r["qs"], r["qe"], hits) == r["ret"]
            assert hits.tolist() == r["hits"]


@pytest.mark.gpu
def test_gpu_cli_gtype0_byte_exact(case, api, workdir):
    for key in case.exp["cli_args"]:
        r = subprocess.run([api.CLI_PATH, "search", case.igd] + case.cli_args(key), capture_output=True, cwd=workdir)
        assert r.returncode == 0, r.stderr.decode()
        assert case.normalise(r.stdout.decode(errors="replace")) == case.stdout(key), (key, r.stderr.decode()[-300:])


@pytest.mark.gpu
def test_gpu_gtype0_equals_gtype1_without_value(twin_dbs, api):
    """every mode on a gType-0 image = the same records with value 0: count and -f identical to the gType-1
    twin, -m identical, -m -v all zero, Seqpare identical (none of them reads `value` without -v)"""
    p1, p0 = twin_dbs
    q = make_queries(9, 4000, ("lognormal", 7.0, 1.5))
    qbed = os.path.join(os.path.dirname(p0), "twin_q.bed")
    synth.write_bed(qbed, q, SMALL_NAMES)
    with api.IGD(p1) as g1, api.IGD(p0) as g0:
        assert np.array_equal(g1.count(q.ctg, q.start, q.end)[0], g0.count(q.ctg, q.start, q.end)[0])
        for a, b in zip(g1.full(q.ctg, q.start, q.end), g0.full(q.ctg, q.start, q.end)):
            assert np.array_equal(a, b)
        assert np.array_equal(g1.hitsmap(0), g0.hitsmap(0))
        assert not g0.hitsmap(500).any()
        assert np.array_equal(g1.seqpare(qbed), g0.seqpare(qbed))
        od = O.OracleDB(p0)
        assert np.array_equal(g0.count(q.ctg, q.start, q.end, v=700)[0], od.count(q.ctg, q.start, q.end, 700)[0])
