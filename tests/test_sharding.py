"""Multi-GPU host logic on CPU: the (contig, start) sort + scan-byte prefix-sum partition of a query set and the
sum-combination of per-shard hits[] -- world_size 2 over gloo, with the CPU oracle standing in for the per-rank
search (the device path is covered by the -m gpu tests; this covers what is different at N > 1)."""
import os
import socket

import numpy as np
import pytest

from igd_b200 import workloads
from tests.conftest import make_dense_db, make_queries


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, igd_path, out_dir):
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    q = make_queries(31, 6000, ("lognormal", 7.0, 1.5))
    hdr = workloads.read_igd_header(igd_path)
    ctg, qs, qe, bounds = workloads.sort_and_shard(hdr, q.ctg, q.start, q.end, world)
    a, b = bounds[rank], bounds[rank + 1]
    db = O.OracleDB(igd_path)
    out = {}
    for v in (0, 500):
        hits, _ = db.count(ctg[a:b], qs[a:b], qe[a:b], v)
        t = torch.from_numpy(hits.copy())
        dist.all_reduce(t)                              # what the NCCL int64 allreduce does on the GPUs
        out[f"hits_v{v}"] = t.numpy()
    cost = workloads.query_cost(hdr, ctg, qs, qe)
    out["shard_cost"] = np.array([cost[a:b].sum()])
    out["bounds"] = np.array(bounds)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **out)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_counts_sum_to_the_whole(workdir, world):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    stem = os.path.join(workdir, f"shard_db{world}")
    make_dense_db(stem, n_files=11, n_per_file=2000, seed=13)
    out_dir = os.path.join(workdir, f"shard_out{world}")
    os.makedirs(out_dir, exist_ok=True)
    mp.spawn(_rank_main, args=(world, _free_port(), stem + ".igd", out_dir), nprocs=world, join=True)
    q = make_queries(31, 6000, ("lognormal", 7.0, 1.5))
    db = O.OracleDB(stem + ".igd")
    res = [np.load(os.path.join(out_dir, f"rank{r}.npz")) for r in range(world)]
    for v in (0, 500):
        want, _ = db.count(q.ctg, q.start, q.end, v)
        for r in res:
            assert np.array_equal(r[f"hits_v{v}"], want)
    # same bounds on every rank, covering every query exactly once, balanced by estimated scan bytes
    b0 = res[0]["bounds"]
    assert all(np.array_equal(r["bounds"], b0) for r in res)
    assert b0[0] == 0 and b0[-1] == 6000 and np.all(np.diff(b0) > 0)
    costs = np.array([r["shard_cost"][0] for r in res], float)
    assert costs.max() / costs.mean() < 1.05


def test_query_cost_matches_tile_table(workdir):
    """16 + sum over touched bins of (8 + 4 nCnt), C truncating division, clamped to the contig's last tile."""
    stem = os.path.join(workdir, "cost_db")
    make_dense_db(stem, n_files=10, n_per_file=500, seed=3)
    hdr = workloads.read_igd_header(stem + ".igd")
    ctg = np.array([0, 0, 1, 2, 5, -1, 0, 2], np.int32)
    qs = np.array([0, 16383, 100, 150000, 5, 5, 10_000_000, 190000], np.int32)
    qe = np.array([1, 16385, 200000, 900000, 50, 50, 10_000_100, 200000], np.int32)
    got = workloads.query_cost(hdr, ctg, qs, qe)
    for i in range(len(ctg)):
        want = 16
        c = ctg[i]
        if 0 <= c < hdr["nctg"]:
            n1, n2, mT = qs[i] // hdr["nbp"], (qe[i] - 1) // hdr["nbp"], hdr["ntile"][c] - 1
            if n1 <= mT:
                for j in range(n1, max(n1, min(n2, mT)) + 1):
                    want += 8 + 4 * int(hdr["ncnt"][hdr["tbase"][c] + j])
        assert got[i] == want, i
