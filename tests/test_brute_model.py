"""A second opinion on the oracle: an independent brute-force model of count mode, written
from the SEMANTICS rather than from the reference's tile walk.

Working the reference's control flow through (Appendix B) gives a closed form: every interval of the query's contig
with start < qe and end > qs is counted exactly once -- the first bin takes those that reach into it, a later bin
those that start in it, and nothing else can overlap -- PROVIDED the query's first bin n1 = qs / nbp exists
(0 <= n1 <= last tile of the contig) and, without -v, is not empty (the first-bin-empty rule, src/igd_search.c:468);
with -v the value filter `value >= v` applies and the empty-first-bin rule does not (:635-659).  The model below is
that sentence in numpy, all pairs, on the un-replicated intervals recovered from the database; it never looks at
tile counts, offsets or sort order.  It must reproduce the committed outputs of the unmodified reference."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.golden_util import GoldenCase


def intervals_of(igd_path):
    """(contig names, nbp, per record: ctg, file, start, end, value) with every interval once (its copy in the
    tile it starts in; an interval with a negative start lives in tile 0)"""
    raw = np.fromfile(igd_path, np.int32)
    nbp, gtype, m = (int(x) for x in raw[:3])
    ntile = raw[3:3 + m].astype(np.int64)
    T = int(ntile.sum())
    ncnt = raw[3 + m:3 + m + T].astype(np.int64)
    names = [raw[3 + m + T + 10 * c:3 + m + T + 10 * (c + 1)].tobytes().split(b"\0")[0].decode() for c in range(m)]
    W = 4 if gtype == 1 else 3
    rec = raw[3 + m + T + 10 * m:].reshape(-1, W)
    ctg = np.repeat(np.repeat(np.arange(m), ntile), ncnt)
    tile = np.repeat(np.concatenate([np.arange(k) for k in ntile]), ncnt)
    first = np.maximum(rec[:, 1], 0) // nbp == tile
    r = rec[first]
    value = r[:, 3] if W == 4 else np.zeros(len(r), np.int32)
    return names, nbp, ctg[first], r[:, 0], r[:, 1].astype(np.int64), r[:, 2].astype(np.int64), value


def brute_count(igd_path, qnames, qs, qe, v, nfiles):
    names, nbp, ctg, fidx, s, e, val = intervals_of(igd_path)
    hits = np.zeros(nfiles, np.int64)
    qs, qe = np.asarray(qs, np.int64), np.asarray(qe, np.int64)
    qc = np.array([names.index(n) if n in names else -1 for n in qnames])
    for c in range(len(names)):
        sel = ctg == c
        S, E, F, V = s[sel], e[sel], fidx[sel], val[sel]
        if len(S) == 0:
            continue
        last_tile = int(((E - 1) // nbp).max())
        occupied = np.zeros(last_tile + 2, bool)               # bins some interval reaches into
        for a, b in zip(np.maximum(S, 0) // nbp, (E - 1) // nbp):
            occupied[a:b + 1] = True
        mine = np.nonzero(qc == c)[0]
        a, b = qs[mine], qe[mine]
        n1 = np.where(a >= 0, a // nbp, -((-a) // nbp))        # C division truncates toward zero
        ok = (n1 >= 0) & (n1 <= last_tile)
        if v <= 0:
            ok &= occupied[np.clip(n1, 0, last_tile)]
        keep = V >= v if v > 0 else np.ones(len(S), bool)
        for i0 in range(0, len(mine), 512):
            aa, bb, kk = a[i0:i0 + 512], b[i0:i0 + 512], ok[i0:i0 + 512]
            M = (S[None, :] < bb[:, None]) & (E[None, :] > aa[:, None]) & kk[:, None] & keep[None, :]
            hits += np.bincount(F, weights=M.sum(0), minlength=nfiles).astype(np.int64)
    return hits


@pytest.mark.parametrize("name", ["quirk", "dense", "cfg1", "dense0"])
def test_brute_model_reproduces_the_reference(name, workdir):
    case = GoldenCase(name, workdir)
    q = O.Queries(case.qfile)                       # the reference-pinned reader (which lines are queries at all)
    qnames = [q.names[i] for i in q.name_id]
    nfiles = len(case.exp["hits"])
    vs = (0,) if name == "dense0" else (0, 1, 500, 1000)
    for v in vs:
        got = brute_count(case.igd, qnames, q.qs, q.qe, v, nfiles)
        assert np.array_equal(got, case.hits(v)), (name, v)
    for r in case.exp["regions"]:
        got = brute_count(case.igd, [r["chrm"]], [r["qs"]], [r["qe"]], r["v"], nfiles)
        assert got.tolist() == r["hits"], r


def test_brute_model_agrees_with_oracle_on_seeded_inputs(dense_db):
    from tests.conftest import SMALL_NAMES, make_queries
    odb = O.OracleDB(dense_db)
    for seed, law in ((1, ("lognormal", 6.0, 1.2)), (2, ("fixed", 40000)), (3, ("lognormal", 9.0, 1.0))):
        q = make_queries(seed, 1500, law)
        names = [SMALL_NAMES[c] for c in q.ctg]
        for v in (0, 300):
            want, _ = odb.count(q.ctg, q.start, q.end, v)
            assert np.array_equal(brute_count(dense_db, names, q.start, q.end, v, odb.nfiles), want), (seed, v)


def brute_hitsmap(igd_path, v, nfiles):
    """-m in closed form: the tile-by-tile self-join with its dedupe rule (a record that starts before the bin
    only meets partners that start inside it, src/igd_search.c:802-804) counts every ORDERED pair of overlapping
    intervals of a contig exactly once -- in the tile where the later of the two starts -- self-pairs included;
    getMap_v keeps the pairs whose two values both exceed v (:856,872)."""
    names, nbp, ctg, fidx, s, e, val = intervals_of(igd_path)
    hm = np.zeros((nfiles, nfiles), np.int64)
    for c in range(len(names)):
        sel = ctg == c
        S, E, F, V = s[sel], e[sel], fidx[sel], val[sel]
        if v > 0:
            k = V > v
            S, E, F = S[k], E[k], F[k]
        for i0 in range(0, len(S), 1024):
            M = (S[None, :] < E[i0:i0 + 1024, None]) & (E[None, :] > S[i0:i0 + 1024, None])
            rows = np.repeat(F[i0:i0 + 1024], M.sum(1))
            cols = np.broadcast_to(F[None, :], M.shape)[M]
            np.add.at(hm, (rows, cols), 1)
    return (hm & 0xffffffff).astype(np.uint32)


@pytest.mark.parametrize("name", ["quirk", "dense", "cfg1"])
def test_brute_hitsmap_reproduces_the_reference(name, workdir):
    case = GoldenCase(name, workdir)
    n = len(case.exp["hits"])
    assert np.array_equal(brute_hitsmap(case.igd, 0, n), case.hitsmap(0))
    assert np.array_equal(brute_hitsmap(case.igd, 500, n), case.hitsmap(500))
