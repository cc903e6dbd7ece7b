"""The pipelined count entry (igd_b200_count_submit / _wait: two batches in flight, H2D of one under the search of the
other) against the one-call entry and the oracle -- including a batch whose plan does not fit the scratch buffers
while the next one is already enqueued behind it (both are then run again)."""
import numpy as np
import pytest

from tests.conftest import make_queries

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from igd_b200 import api
    return api


def _batches():
    # sizes and shapes that change from batch to batch: short, long (many bins per query: the (query, tile) expansion
    # outgrows the 2 n + 2 T segments the scratch is sized for -> misfit), empty, abnormal (qs >= qe, second pipeline)
    b = [make_queries(201, 5000, ("lognormal", 6.0, 1.0)),
         make_queries(202, 4000, ("fixed", 150000)),
         make_queries(203, 9000, ("fixed", 1000)),
         make_queries(204, 0, ("fixed", 1000)),
         make_queries(205, 6000, ("fixed", 300000)),
         make_queries(206, 7000, ("lognormal", 5.0, 1.5))]
    odd = make_queries(207, 3000, ("fixed", 500))
    odd.start[::3] = odd.end[::3] + 5           # qs >= qe: accepted by parse_bed, left to the int32 pipeline
    b.append(odd)
    return b


@pytest.mark.parametrize("v", [0, 500])
def test_pipelined_equals_one_call_and_oracle(dense_db, api, v):
    from oracle import oracle as O
    odb = O.OracleDB(dense_db)
    g = api.IGD(dense_db)
    bs = _batches()
    arrs = [tuple(np.ascontiguousarray(x, dtype=np.int32) for x in (q.ctg, q.start, q.end)) for q in bs]
    want = [odb.count(q.ctg, q.start, q.end, v)[0] for q in bs]
    got = [np.zeros(g.nfiles, np.int64) for _ in bs]
    ret = [None] * len(bs)

    def submit(i):
        c, s, e = arrs[i]
        g.count_submit(i & 1, len(c), c.ctypes.data, s.ctypes.data, e.ctypes.data, v)

    submit(0)
    for i in range(1, len(bs)):
        submit(i)
        ret[i - 1] = g.count_wait((i - 1) & 1, got[i - 1])
    ret[-1] = g.count_wait((len(bs) - 1) & 1, got[-1])
    for i in range(len(bs)):
        assert np.array_equal(got[i], want[i]), f"batch {i}"
        assert ret[i] == int(want[i].sum())
    # the handle is back to normal: the one-call entry gives the same
    one, _ = g.count(*arrs[1], v)
    assert np.array_equal(one, want[1])
    g.close()
    odb.close()


def test_pipeline_rules(dense_db, api):
    g = api.IGD(dense_db)
    q = make_queries(208, 1000, ("fixed", 1000))
    c, s, e = (np.ascontiguousarray(x, dtype=np.int32) for x in (q.ctg, q.start, q.end))
    hits = np.zeros(g.nfiles, np.int64)
    with pytest.raises(api.IGDError):
        g.count_wait(0, hits)                                   # nothing submitted
    g.count_submit(0, len(c), c.ctypes.data, s.ctypes.data, e.ctypes.data)
    with pytest.raises(api.IGDError):
        g.count_submit(0, len(c), c.ctypes.data, s.ctypes.data, e.ctypes.data)      # slot busy
    with pytest.raises(api.IGDError):
        g.count(c, s, e)                                        # the one-call entry does not mix with a busy slot
    g.count_submit(1, len(c), c.ctypes.data, s.ctypes.data, e.ctypes.data)
    with pytest.raises(api.IGDError):
        g.count_wait(1, hits)                                   # out of submission order
    g.count_wait(0, hits)
    g.count_wait(1, hits)
    ref, _ = g.count(c, s, e)
    assert np.array_equal(hits, 2 * ref)                        # incremented, like igd_b200_count
    g.close()
