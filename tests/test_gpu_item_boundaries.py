"""Work-item boundaries of the count kernels (an item = up to 256 entries of one tile x up to 64 segments): tiles
of every size around 8 / 256 / 512 / 768 records, and tiles whose first 300 records all start before the bin (the
later-bin dedupe rule decided in a tile's second item), with and without -v, against the oracle.

(A variant of the packed kernel that let a tile's last item carry up to 64 records past 256 -- 26 % fewer items at
BASELINE config 2 -- was built on these tests and measured SLOWER, 0.167 vs 0.159 ms: the small remainder items are
cheap, their windows prune almost every segment.  It was dropped; profiles/r1l_tail_variant.json keeps the numbers.)"""
import os

import numpy as np
import pytest

from igd_b200 import synth

pytestmark = pytest.mark.gpu

NBP = 16384


def boundary_db(stem):
    """one contig; tile k holds a chosen number of records around 256 / 320 / 512 / 576 ..., plus 300 long records
    that start in tile 0 and span tiles 0..3 (so tiles 1..3 begin with 300 records that start before the bin)"""
    sizes = [1, 2, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 255, 256, 257, 300, 319, 320, 321, 322, 383, 384, 500, 511, 512, 513, 525, 575, 576, 577,
             578, 640, 767, 768, 769, 832, 833, 1100]
    rng = np.random.default_rng(2024)
    files = []
    nfiles = 12
    per_file = [[] for _ in range(nfiles)]
    for k, n in enumerate(sizes):
        t = 4 + k                                   # tiles 4.. hold the sized sets (short records inside the bin)
        s = rng.integers(t * NBP, (t + 1) * NBP - 700, n)
        if n > 4:
            s[: n // 3] = s[0]                      # duplicate starts (SURVEY §8c's list of known-answer cases)
        ln = rng.integers(1, 600, n)
        for a, b in zip(s, ln):
            per_file[int(rng.integers(0, nfiles))].append((0, int(a), int(a + b), int(rng.integers(0, 1000))))
    for j in range(300):                            # long records over tiles 0..3
        a = int(rng.integers(100, 9000))
        per_file[j % nfiles].append((0, a, a + 3 * NBP + int(rng.integers(0, 6000)), int(rng.integers(0, 1000))))
    for t in (1, 2, 3):                             # a few in-bin records in those tiles: 300 + r records each
        for j in range((5, 20, 63)[t - 1]):
            a = int(rng.integers(t * NBP + 10, (t + 1) * NBP - 500))
            per_file[j % nfiles].append((0, a, a + int(rng.integers(1, 400)), int(rng.integers(0, 1000))))
    for rows in per_file:
        r = np.array(rows, np.int64)
        files.append(synth.Intervals(r[:, 0].astype(np.int32), r[:, 1].astype(np.int32), r[:, 2].astype(np.int32),
                                     r[:, 3].astype(np.int32)))
    synth.write_igd_direct(stem, files, names=["chrT"])
    return (4 + len(sizes)) * NBP


def queries(span, seed, n):
    rng = np.random.default_rng(seed)
    s = rng.integers(0, span, n)
    ln = np.where(rng.random(n) < 0.8, rng.integers(1, 2000, n), rng.integers(2000, 5 * NBP, n))
    return np.zeros(n, np.int32), s.astype(np.int32), (s + ln).astype(np.int32)


@pytest.fixture(scope="module")
def db(workdir):
    stem = os.path.join(workdir, "tail_db")
    span = boundary_db(stem)
    return stem + ".igd", span


@pytest.mark.parametrize("v", [0, 1, 500])
def test_item_boundaries(db, v):
    from igd_b200 import api
    from oracle import oracle as O
    path, span = db
    odb = O.OracleDB(path)
    with api.IGD(path) as g:
        for seed, n in ((1, 30000), (2, 17), (3, 1000)):
            ctg, qs, qe = queries(span, seed, n)
            want, _ = odb.count(ctg, qs, qe, v)
            got, _ = g.count(ctg, qs, qe, v)
            assert np.array_equal(got, want), (v, seed)
        items = g.last_plan()["items"]
    assert want.sum() > 0 and items > 0


@pytest.mark.parametrize("v", [0, 500])
def test_dense_tiles_long_tailed_lengths(workdir, v):
    """thousands of records per tile with a long-tailed length distribution (the END half of the count-mode copy
    is a tile sorted by end: k_pack_end), also for queries spanning many bins (the later-bin rule)"""
    from igd_b200 import api
    from oracle import oracle as O
    from tests.conftest import SMALL_NAMES, SMALL_SIZES
    stem = os.path.join(workdir, "dense_split")
    if not os.path.exists(stem + ".igd"):
        files = [synth.gen_intervals(70000 + i, 20000, ("lognormal", 6.3, 1.5), SMALL_SIZES) for i in range(15)]
        synth.write_igd_direct(stem, files, names=SMALL_NAMES)
    odb = O.OracleDB(stem + ".igd")
    assert odb.nrecords / sum(odb.ntile) > 2500
    qa = synth.gen_intervals(71, 30000, ("lognormal", 6.5, 1.2), SMALL_SIZES)
    qb = synth.gen_intervals(72, 2000, ("fixed", 90000), SMALL_SIZES)
    ctg = np.concatenate([qa.ctg, qb.ctg])
    qs, qe = np.concatenate([qa.start, qb.start]), np.concatenate([qa.end, qb.end])
    want, _ = odb.count(ctg, qs, qe, v)
    with api.IGD(stem + ".igd") as g:
        got, _ = g.count(ctg, qs, qe, v)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("v", [0, 500])
def test_very_dense_tiles_every_windows_kernel(workdir, v):
    """tiles of ~12, ~36, ~71, ~110 and ~164 groups of 256 records: the descriptor kernel's three search depths (32,
    64, 128 staged groups: k_windows_warp) and the tiles beyond them (k_windows<32>), with ~1000 segments (16 chunks)
    per tile"""
    from igd_b200 import api
    from oracle import oracle as O
    names = ["chrP", "chrQ", "chrR", "chrS", "chrT"]
    sizes = np.full(5, 4 * 16384, dtype=np.int64)
    stem = os.path.join(workdir, "very_dense")
    if not os.path.exists(stem + ".igd"):
        files = []
        for i in range(8):
            rng = np.random.default_rng(81000 + i)
            n = 50000
            ctg = rng.choice(5, size=n, p=[0.03, 0.09, 0.18, 0.28, 0.42]).astype(np.int32)
            ln = np.clip(np.exp(rng.normal(5.5, 1.2, n)), 1, 60000).astype(np.int64)
            start = np.floor(rng.random(n) * np.maximum(sizes[ctg] - ln, 1)).astype(np.int64)
            end = np.minimum(start + ln, sizes[ctg])
            files.append(synth.Intervals(ctg, start.astype(np.int32), end.astype(np.int32),
                                         rng.integers(0, 1001, size=n, dtype=np.int32)))
        synth.write_igd_direct(stem, files, names=names)
    odb = O.OracleDB(stem + ".igd")
    qa = synth.gen_intervals(82, 20000, ("lognormal", 6.0, 1.2), sizes)
    qb = synth.gen_intervals(83, 500, ("fixed", 40000), sizes)
    ctg = np.concatenate([qa.ctg, qb.ctg])
    qs, qe = np.concatenate([qa.start, qb.start]), np.concatenate([qa.end, qb.end])
    want, _ = odb.count(ctg, qs, qe, v)
    with api.IGD(stem + ".igd") as g:
        most = max(g.ncnt(c, t) for c in range(5) for t in range(g.ntile(c)))
        assert most >= 128 * 256                        # a tile beyond the staged kernel's reach is in there
        got, _ = g.count(ctg, qs, qe, v)
    assert want.sum() > 0 and np.array_equal(got, want)


def test_item_boundary_tiles_full_hitsmap_seqpare(db, workdir):
    """the same tiles (1, 2, 15, 16, 17, 31, 32, 33 ... records, duplicate starts) through the query-centric
    kernels: -f (the 32-ary ballot search and the descending scan), -m, -m -v, and Seqpare's candidate lists"""
    from igd_b200 import api
    from oracle import oracle as O
    path, span = db
    odb = O.OracleDB(path)
    ctg, qs, qe = queries(span, 9, 6000)
    with api.IGD(path) as g:
        per, _, s, e, f = odb.full(ctg, qs, qe)
        qoff, gs, ge, gf = g.full(ctg, qs, qe)
        assert np.array_equal(np.diff(qoff), per) and np.array_equal(gs, s) and np.array_equal(ge, e) and np.array_equal(gf, f)
        for v in (0, 500):
            assert np.array_equal(g.hitsmap(v), odb.hitsmap(v)), v
        off, rec = g.seq_overlaps(ctg[:400], qs[:400], qe[:400])
        for i in range(0, 400, 7):
            assert rec[off[i]:off[i + 1]].tobytes() == odb.seq_overlaps(int(ctg[i]), int(qs[i]), int(qe[i])).tobytes(), i
