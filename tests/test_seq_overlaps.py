"""seq_overlaps (src/igd_search.h:19, src/igd_search.c:253-353): the per-query candidate list behind Seqpare --
overlap_t {idx_t, idx_g, idx_f, float sm} records in the reference's scan order.

The oracle's restatement is pinned on the unmodified reference's own seq_overlaps (live when oracle/_ref is built,
and through the committed digests of tests/golden/seq_overlaps.json otherwise); the CUDA path (batched
igd_b200_seq_overlaps and the reference-named one-query seq_overlaps) is held to the oracle record for record,
the fp32 similarity bit for bit."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.conftest import make_queries
from tests.golden_util import GOLDEN, GoldenCase

FIX = os.path.join(GOLDEN, "seq_overlaps.json")
CASES = ["quirk", "dense", "cfg1"]


def sample(q, ctg, k=300):
    return [i for i in range(0, q.n, max(1, q.n // k)) if ctg[i] >= 0]


@pytest.mark.parametrize("name", CASES)
def test_oracle_vs_committed_reference_digest(name, workdir):
    case = GoldenCase(name, workdir)
    db, q = O.OracleDB(case.igd), O.Queries(case.qfile)
    ctg = db.resolve(q)
    h, n = hashlib.sha256(), 0
    for i in sample(q, ctg):
        r = db.seq_overlaps(int(ctg[i]), int(q.qs[i]), int(q.qe[i]))
        h.update(r.tobytes())
        n += len(r)
    exp = json.load(open(FIX))[name]
    assert (n, h.hexdigest()) == (exp["records"], exp["sha256"])


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", CASES)
def test_oracle_vs_reference_live(name, workdir):
    case = GoldenCase(name, workdir)
    db, ref, q = O.OracleDB(case.igd), O.RefLib(case.igd), O.Queries(case.qfile)
    ctg = db.resolve(q)
    try:
        for i in sample(q, ctg):
            a = db.seq_overlaps(int(ctg[i]), int(q.qs[i]), int(q.qe[i]))
            b = ref.seq_overlaps(q.names[q.name_id[i]], int(q.qs[i]), int(q.qe[i]))
            assert a.tobytes() == b.tobytes(), (name, i)
    finally:
        ref.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_seq_overlaps_golden_cases(name, workdir):
    from igd_b200 import api
    case = GoldenCase(name, workdir)
    db, q = O.OracleDB(case.igd), O.Queries(case.qfile)
    ctg = db.resolve(q)
    with api.IGD(case.igd) as g:
        off, rec = g.seq_overlaps(ctg, q.qs, q.qe)
        assert off[0] == 0 and off[-1] == len(rec)
        h, n = hashlib.sha256(), 0
        for i in sample(q, ctg):
            mine = rec[off[i]:off[i + 1]]
            assert mine.tobytes() == db.seq_overlaps(int(ctg[i]), int(q.qs[i]), int(q.qe[i])).tobytes(), (name, i)
            h.update(mine.tobytes())
            n += len(mine)
        exp = json.load(open(FIX))[name]
        assert (n, h.hexdigest()) == (exp["records"], exp["sha256"])         # = the reference's own records
        # the reference-named entry point appends to a caller-owned, growable overlaps_t
        L = api.load_library()
        g.use()
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        o = api._Overlaps(0, 0, None)
        picks = sample(q, ctg, 12)
        for i in picks:
            L.seq_overlaps(q.names[q.name_id[i]].encode(), int(q.qs[i]), int(q.qe[i]), C.byref(o))
        want = np.concatenate([rec[off[i]:off[i + 1]] for i in picks]) if picks else rec[:0]
        assert o.nn == len(want) and o.mm >= o.nn
        got = np.frombuffer(C.string_at(o.olist, 16 * o.nn), api.OVERLAP_DTYPE) if o.nn else want
        assert got.tobytes() == want.tobytes()
        if o.olist:
            libc.free(o.olist)


@pytest.mark.gpu
def test_gpu_seq_overlaps_seeded(dense_db):
    """long queries over a dense database: many bins per query, hundreds of records per list"""
    from igd_b200 import api
    db = O.OracleDB(dense_db)
    q = make_queries(77, 400, ("lognormal", 9.5, 1.0))
    with api.IGD(dense_db) as g:
        off, rec = g.seq_overlaps(q.ctg, q.start, q.end)
    assert len(rec) > 50000
    for i in range(len(q.ctg)):
        want = db.seq_overlaps(int(q.ctg[i]), int(q.start[i]), int(q.end[i]))
        assert rec[off[i]:off[i + 1]].tobytes() == want.tobytes(), i


def make_fixture():
    """python -m tests.test_seq_overlaps (needs oracle/_ref): digests of the REFERENCE's records"""
    import tempfile
    assert O.have_ref()
    exp = {}
    with tempfile.TemporaryDirectory() as d:
        for name in CASES:
            case = GoldenCase(name, d)
            db, ref, q = O.OracleDB(case.igd), O.RefLib(case.igd), O.Queries(case.qfile)
            ctg = db.resolve(q)
            h, n = hashlib.sha256(), 0
            for i in sample(q, ctg):
                r = ref.seq_overlaps(q.names[q.name_id[i]], int(q.qs[i]), int(q.qe[i]))
                h.update(r.tobytes())
                n += len(r)
            ref.close()
            exp[name] = {"records": n, "sha256": h.hexdigest()}
    json.dump(exp, open(FIX, "w"), indent=1)
    print(exp)


if __name__ == "__main__":
    make_fixture()
