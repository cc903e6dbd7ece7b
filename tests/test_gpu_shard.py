"""K8 on the device (igd_dev_shard.cu: igd_b200_shard_cost / _cuts / _partition, and the tile-range form of the count
entry) against a numpy statement of the same plan: per-tile cost = sum of igd_b200_query_cost over the queries whose
first bin the tile is; cuts = searchsorted of its prefix sum; partition = the queries regrouped by destination."""
import numpy as np
import pytest

from igd_b200 import workloads
from tests.conftest import make_queries

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("parts", [1, 2, 3, 8])
def test_shard_plan_matches_numpy(dense_db, parts):
    import torch
    from igd_b200 import api
    from oracle import oracle as O
    hdr = workloads.read_igd_header(dense_db)
    q = make_queries(123, 20000, ("lognormal", 7.5, 1.6))
    rng = np.random.default_rng(5)
    ctg, qs, qe = q.ctg.copy(), q.start.copy(), q.end.copy()
    ctg[:50] = rng.integers(-3, 8, 50)                              # unknown contigs, starts beyond the last tile
    qs[50:90] += 5_000_000
    qe[50:90] += 5_000_000
    with api.IGD(dense_db) as g:
        T = g.ntiles
        dev = torch.device("cuda", 0)
        d = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (ctg, qs, qe)]
        cost = torch.zeros(T + 1, dtype=torch.int64, device=dev)
        cuts = torch.zeros(parts + 1, dtype=torch.int32, device=dev)
        off = torch.zeros(parts + 1, dtype=torch.int64, device=dev)
        reach = torch.zeros(parts, dtype=torch.int32, device=dev)
        out = torch.empty((3, len(ctg)), dtype=torch.int32, device=dev)
        n = len(ctg)
        torch.cuda.synchronize()
        st = torch.cuda.Stream(device=dev)              # one real stream for the library's kernels and torch's
        torch.cuda.set_stream(st)
        g.shard_cost(n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), cost.data_ptr(), st.cuda_stream)
        g.shard_cuts(cost.data_ptr(), parts, cuts.data_ptr(), st.cuda_stream)
        g.shard_partition(n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), cuts.data_ptr(), parts,
                          out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), off.data_ptr(), reach.data_ptr(),
                          st.cuda_stream)
        torch.cuda.synchronize()
        cost, cuts, off, reach, out = (x.cpu().numpy() for x in (cost, cuts, off, reach, out))
        # the numpy model
        qc = g.query_cost(ctg, qs, qe)
        ok = (ctg >= 0) & (ctg < hdr["nctg"])
        c0 = np.where(ok, ctg, 0)
        n1 = qs // hdr["nbp"]
        ok &= (qs >= 0) & (n1 < hdr["ntile"][c0])
        tile = np.where(ok, hdr["tbase"][c0] + n1, T)
        want_cost = np.bincount(tile, weights=qc.astype(np.float64), minlength=T + 1).astype(np.int64)
        assert np.array_equal(cost, want_cost)
        cum = np.cumsum(want_cost[:T])
        want_cuts = [0] + [int(np.searchsorted(cum, cum[-1] * p / parts)) for p in range(1, parts)] + [T]
        assert cuts.tolist() == want_cuts
        dest = np.where(ok, np.searchsorted(np.array(want_cuts[1:-1]), tile, side="right"), parts - 1)
        assert off.tolist() == [0] + np.cumsum(np.bincount(dest, minlength=parts)).tolist()
        key = lambda a, b, c: set(zip(a.tolist(), b.tolist(), c.tolist()))
        for p in range(parts):
            sel = dest == p
            a, b = off[p], off[p + 1]
            assert sorted(zip(out[0, a:b].tolist(), out[1, a:b].tolist(), out[2, a:b].tolist())) == \
                sorted(zip(ctg[sel].tolist(), qs[sel].tolist(), qe[sel].tolist()))
            if sel.any() and ok[sel].any():
                s2 = sel & ok
                n2 = np.minimum((qe[s2] - 1) // hdr["nbp"], hdr["ntile"][ctg[s2]] - 1)
                last = hdr["tbase"][ctg[s2]] + np.maximum(n1[s2], n2)
                assert reach[p] == last.max() + 1
        # the shards, searched one by one over their tile ranges, add up to the whole
        odb = O.OracleDB(dense_db)
        known = (ctg >= 0) & (ctg < hdr["nctg"])                   # an unknown chrom counts nothing (src/igd_search.c:456-458)
        want, _ = odb.count(ctg[known], qs[known], qe[known], 0)
        hits = torch.zeros(g.nfiles, dtype=torch.int64, device=dev)
        dout = torch.from_numpy(out).to(dev)
        for p in range(parts):
            a, b = int(off[p]), int(off[p + 1])
            if b > a:
                g.count_device_tiles(b - a, dout[0, a:].data_ptr(), dout[1, a:].data_ptr(), dout[2, a:].data_ptr(), 0,
                                     hits.data_ptr(), st.cuda_stream, want_cuts[p], max(int(reach[p]), want_cuts[p]))
                g.count_device_finish()
        torch.cuda.synchronize()
        got = hits.cpu().numpy()
        torch.cuda.set_stream(torch.cuda.default_stream(dev))
        assert np.array_equal(got, want), (int(got.sum()), int(want.sum()))
