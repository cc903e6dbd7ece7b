"""igd_b200.multi.ShardedSearch -- the N > 1 logic of count, -f and -m (shard, search, combine, restore order).

CPU (`not gpu`): world_size 2 and 3 over gloo with the CPU oracle standing in as the per-rank engine; what is
checked is everything that differs from the single-GPU path: the partition, the allreduce, the gather and the
restoration of the input's query order.
GPU (`gpu`): the same with the CUDA library as the engine, two ranks (on two GPUs when the box has them,
otherwise both on cuda:0 with gloo carrying the collectives)."""
import os
import socket

import numpy as np
import pytest

from igd_b200 import multi, workloads
from tests.conftest import make_dense_db, make_queries


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleEngine:
    """Adapter: the oracle with the engine interface of igd_b200.api.IGD (tests only)."""

    def __init__(self, path):
        from oracle import oracle as O
        self.db = O.OracleDB(path)
        self.nfiles = self.db.nfiles

    def count(self, ctg, qs, qe, v=0):
        hits, _ = self.db.count(ctg, qs, qe, v)
        return hits, int(hits.sum())

    def full(self, ctg, qs, qe):
        per, _, s, e, f = self.db.full(ctg, qs, qe)
        return np.concatenate([[0], np.cumsum(per)]).astype(np.int64), s, e, f

    def hitsmap(self, v=0, tiles=None):
        return self.db.hitsmap(v, tiles)


def _queries():
    q = make_queries(77, 5000, ("lognormal", 7.5, 1.6))
    rng = np.random.default_rng(3)
    p = rng.permutation(len(q))                                     # unsorted, as a query file would be
    return q.ctg[p], q.start[p], q.end[p]


def _rank_main(rank, world, port, igd_path, out_dir, engine_kind):
    import torch
    import torch.distributed as dist
    comm = "cpu"
    if engine_kind == "cuda":
        from igd_b200 import api
        ndev = torch.cuda.device_count()
        if ndev >= world:
            torch.cuda.set_device(rank)
            dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                    device_id=torch.device("cuda", rank))
            comm = f"cuda:{rank}"
            eng = api.IGD(igd_path, device=rank)
        else:
            dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
            eng = api.IGD(igd_path, device=0)
    else:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        eng = OracleEngine(igd_path)
    sh = multi.ShardedSearch(eng, workloads.read_igd_header(igd_path), comm_device=comm)
    ctg, qs, qe = _queries()
    out = {}
    for v in (0, 500):
        out[f"hits_v{v}"] = sh.count(ctg, qs, qe, v)
        out[f"map_v{v}"] = sh.hitsmap(v)
    if comm != "cpu":
        # K8 on the devices: every rank starts from a 1/world slice of the unsorted queries (multi.DeviceSharder)
        import torch
        n = len(ctg)
        a, b = n * rank // world, n * (rank + 1) // world
        dev = torch.device(comm)
        st = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(st):
            shd = multi.DeviceSharder(eng, dev)
            t = [torch.from_numpy(np.ascontiguousarray(x[a:b])).to(dev) for x in (ctg, qs, qe)]
            c, s_, e_, lo, hi = shd.shard(*t)
            hits = torch.zeros(eng.nfiles, dtype=torch.int64, device=dev)
            if c.shape[0]:
                eng.count_device_tiles(int(c.shape[0]), c.data_ptr(), s_.data_ptr(), e_.data_ptr(), 0, hits.data_ptr(),
                                       st.cuda_stream, lo, hi)
                eng.count_device_finish()
            dist.all_reduce(hits)
            got = torch.tensor([int(c.shape[0])], dtype=torch.int64, device=dev)
            dist.all_reduce(got)
            assert int(got.item()) == n
            out["hits_sharded_on_device"] = hits.cpu().numpy()
        st.synchronize()
    r = sh.full(ctg, qs, qe, dst=world - 1)
    assert (r is None) == (rank != world - 1)
    if r is not None:
        out["qoff"], out["s"], out["e"], out["f"] = r
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **out)
    dist.destroy_process_group()


def _check(workdir, world, engine_kind):
    import torch.multiprocessing as mp
    stem = os.path.join(workdir, f"multi_db_{engine_kind}{world}")
    make_dense_db(stem, n_files=13, n_per_file=1500, seed=21)
    out_dir = os.path.join(workdir, f"multi_out_{engine_kind}{world}")
    os.makedirs(out_dir, exist_ok=True)
    mp.spawn(_rank_main, args=(world, _free_port(), stem + ".igd", out_dir, engine_kind), nprocs=world, join=True)
    ref = OracleEngine(stem + ".igd")
    ctg, qs, qe = _queries()
    res = [np.load(os.path.join(out_dir, f"rank{r}.npz")) for r in range(world)]
    for v in (0, 500):
        want, _ = ref.count(ctg, qs, qe, v)
        wmap = ref.hitsmap(v)
        assert want.sum() > 0 and wmap.sum() > 0
        for r in res:
            assert np.array_equal(r[f"hits_v{v}"], want)
            assert np.array_equal(r[f"map_v{v}"], wmap)
            if v == 0 and "hits_sharded_on_device" in r.files:
                assert np.array_equal(r["hits_sharded_on_device"], want)
    qoff, s, e, f = ref.full(ctg, qs, qe)                           # single engine, input order
    last = res[world - 1]
    assert np.array_equal(last["qoff"], qoff)
    assert np.array_equal(last["s"], s) and np.array_equal(last["e"], e) and np.array_equal(last["f"], f)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_gloo_cpu(workdir, world):
    _check(workdir, world, "oracle")


@pytest.mark.gpu
def test_sharded_search_cuda_two_ranks(workdir):
    _check(workdir, 2, "cuda")


def test_tile_bounds_cover_and_balance(workdir):
    stem = os.path.join(workdir, "tb_db")
    make_dense_db(stem, n_files=10, n_per_file=800, seed=5)
    hdr = workloads.read_igd_header(stem + ".igd")
    for parts in (1, 2, 3, 8):
        b = multi.tile_bounds(hdr, parts)
        assert b[0] == 0 and b[-1] == len(hdr["ncnt"]) and all(x <= y for x, y in zip(b, b[1:]))
        n = hdr["ncnt"].astype(float)
        cost = np.array([(n[b[i]:b[i + 1]] ** 2).sum() for i in range(parts)])
        if parts <= 3:
            assert cost.max() < 1.3 * cost.mean()


@pytest.mark.gpu
def test_hitsmap_tile_ranges_add_up(dense_db):
    """igd_b200_hitsmap_tiles: maps of disjoint tile ranges add up to the whole map, each equal to the oracle's."""
    from igd_b200 import api
    from oracle import oracle as O
    g, o = api.IGD(dense_db), O.OracleDB(dense_db)
    T = g.ntiles
    cuts = [0, 1, T // 3, T // 3, T - 2, T]
    for v in (0, 300):
        acc = np.zeros((g.nfiles, g.nfiles), np.uint32)
        for lo, hi in zip(cuts, cuts[1:]):
            part = g.hitsmap(v, tiles=(lo, hi))
            assert np.array_equal(part, o.hitsmap(v, tiles=(lo, hi))), (v, lo, hi)
            acc += part
        assert np.array_equal(acc, g.hitsmap(v)) and np.array_equal(acc, o.hitsmap(v))
    g.close()


def _reducer_rank(rank, world, port, out_dir):
    import signal
    import torch
    import torch.distributed as dist
    from igd_b200 import api
    signal.alarm(150)                                   # a lost flag would spin forever: die instead
    ndev = torch.cuda.device_count()
    dev = rank if ndev >= world else 0                  # one GPU: the ranks share it (IPC works within a device too)
    torch.cuda.set_device(dev)
    if ndev >= world:
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    nfiles = 1905 if world == 2 else 10000              # 4 and 20 slices (CTAs) of the file range
    red = api.HitsReducer(dev, nfiles)
    stream = torch.cuda.Stream(device=dev)
    ok = True
    with torch.cuda.stream(stream):
        for epoch in range(7):
            g = torch.Generator().manual_seed(1000 * epoch + rank)
            mine = torch.randint(0, 2**40, (nfiles,), generator=g, dtype=torch.int64).to(f"cuda:{dev}")
            want = sum(torch.randint(0, 2**40, (nfiles,), generator=torch.Generator().manual_seed(1000 * epoch + q),
                                     dtype=torch.int64) for q in range(world))
            if epoch & 1:                               # out of place: the input is left zeroed
                out = torch.full((nfiles,), -1, dtype=torch.int64, device=f"cuda:{dev}")
                red.allreduce_into(mine.data_ptr(), out.data_ptr(), stream.cuda_stream)
                stream.synchronize()
                ok = ok and bool(torch.equal(out.cpu(), want)) and int(mine.abs().sum()) == 0
            else:
                red.allreduce(mine.data_ptr(), stream.cuda_stream)
                stream.synchronize()
                ok = ok and bool(torch.equal(mine.cpu(), want))
    ok = ok and red.status() == 0
    red.close()
    open(os.path.join(out_dir, f"reducer{rank}.txt"), "w").write("ok" if ok else "MISMATCH")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_peer_memory_hits_reducer(workdir, world):
    """igd_b200_reducer_*: hits[] summed over the ranks by one kernel per rank over peer memory, seven epochs in a
    row (both mailbox slots reused), against the sums computed on the host.  On a one-GPU box the ranks share the
    device (their kernels wait for each other across time slices)."""
    import torch.multiprocessing as mp
    out_dir = os.path.join(workdir, f"reducer_out{world}")
    os.makedirs(out_dir, exist_ok=True)
    mp.spawn(_reducer_rank, args=(world, _free_port(), out_dir), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(out_dir, f"reducer{r}.txt")).read() == "ok"


def _dispatch_rank(rank, world, port, igd_path, out_dir):
    import signal
    import torch
    import torch.distributed as dist
    from igd_b200 import api
    signal.alarm(200)                                   # a lost flag would spin (20 s, then an error): die instead
    ndev = torch.cuda.device_count()
    dev = rank if ndev >= world else 0                  # one GPU: the ranks share it
    torch.cuda.set_device(dev)
    if ndev >= world:
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = api.IGD(igd_path, device=dev)
    hdr = workloads.read_igd_header(igd_path)
    T, tdev = g.ntiles, torch.device("cuda", dev)
    ok = True
    stream = torch.cuda.Stream(device=tdev)
    with torch.cuda.stream(stream):
        st = stream.cuda_stream
        disp = api.QueryDispatcher(g, dev, 40000)
        res = {}
        for epoch, (seed, n) in enumerate([(301, 30000), (302, 7), (303, 40000), (304, 0), (305, 20000)]):
            q = make_queries(seed, n, ("lognormal", 7.0, 1.5))
            ctg, qs, qe = q.ctg.copy(), q.start.copy(), q.end.copy()
            if n > 100:
                ctg[:20] = -1                           # unknown contigs go to the last rank and count nothing
            a, b = n * rank // world, n * (rank + 1) // world
            full = [torch.from_numpy(np.ascontiguousarray(x)).to(tdev) for x in (ctg, qs, qe)]
            mine = [x[a:b].contiguous() for x in full]
            cost = torch.zeros(T + 1, dtype=torch.int64, device=tdev)
            cuts = torch.zeros(world + 1, dtype=torch.int32, device=tdev)
            if n:
                g.shard_cost(n, full[0].data_ptr(), full[1].data_ptr(), full[2].data_ptr(), cost.data_ptr(), st)
            g.shard_cuts(cost.data_ptr(), world, cuts.data_ptr(), st)
            m, pc, ps, pe, lo, hi = disp.dispatch(b - a, mine[0].data_ptr(), mine[1].data_ptr(), mine[2].data_ptr(),
                                                  cuts.data_ptr(), st)
            got = [torch.as_tensor(api.device_view(p, m, "<i4"), device=tdev).cpu().numpy() if m else np.zeros(0, np.int32)
                   for p in (pc, ps, pe)]
            # the numpy model of the plan (tests/test_gpu_shard.py): destination = the cut range of the first tile
            cu = cuts.cpu().numpy()
            okq = (ctg >= 0) & (ctg < hdr["nctg"])
            c0 = np.where(okq, ctg, 0)
            n1 = qs // hdr["nbp"]
            okq &= (qs >= 0) & (n1 < hdr["ntile"][c0])
            tile = np.where(okq, hdr["tbase"][c0] + n1, T)
            dest = np.where(okq, np.searchsorted(cu[1:-1], tile, side="right"), world - 1)
            src = np.minimum(np.searchsorted(np.array([n * r // world for r in range(1, world)]), np.arange(n), side="right"),
                             world - 1) if n else np.zeros(0, np.int64)
            pos = 0
            for s_ in range(world):                     # rank-major by source; inside a source any order
                sel = (dest == rank) & (src == s_)
                k = int(sel.sum())
                want = sorted(zip(ctg[sel].tolist(), qs[sel].tolist(), qe[sel].tolist()))
                have = sorted(zip(got[0][pos:pos + k].tolist(), got[1][pos:pos + k].tolist(), got[2][pos:pos + k].tolist()))
                ok = ok and want == have
                pos += k
            ok = ok and pos == m and lo == int(cu[rank])
            hits = torch.zeros(g.nfiles, dtype=torch.int64, device=tdev)
            if m:
                g.count_device_tiles(m, pc, ps, pe, 0, hits.data_ptr(), st, lo, hi)
                g.count_device_finish()
            stream.synchronize()
            # the same exchange with the host out of the loop: the shape of the shard stays on the device
            qc, qs_, qe_, shape = disp.dispatch_async(b - a, mine[0].data_ptr(), mine[1].data_ptr(), mine[2].data_ptr(),
                                                      cuts.data_ptr(), st)
            hits2 = torch.zeros(g.nfiles, dtype=torch.int64, device=tdev)
            g.count_device_shaped(40000, shape, qc, qs_, qe_, 0, hits2.data_ptr(), st)
            g.count_device_finish()
            stream.synchronize()
            sh = torch.as_tensor(api.device_view(shape, 3, "<i8"), device=tdev).cpu().numpy()
            ok = ok and (qc, qs_, qe_) == (pc, ps, pe) and [int(sh[0]), int(sh[1]), max(int(sh[1]), int(sh[2]))] == [m, lo, hi] \
                and bool(torch.equal(hits, hits2))
            h = hits if ndev >= world else hits.cpu()   # (NCCL reduces device tensors, gloo host tensors)
            dist.all_reduce(h)
            res[f"hits{epoch}"] = h.cpu().numpy()
        disp.status()
        disp.close()
        # the whole step as ONE C call (igd_b200_sharded_*): cost, its sum over the ranks, cuts, exchange, search, sum
        sc = api.ShardedCounter(g, dev, 40000)
        for epoch, (seed, n) in enumerate([(311, 25000), (312, 0), (313, 40000), (314, 3)]):
            q = make_queries(seed, n, ("lognormal", 7.0, 1.5))
            a, b = n * rank // world, n * (rank + 1) // world
            mine = [torch.from_numpy(np.ascontiguousarray(x[a:b])).to(tdev) for x in (q.ctg, q.start, q.end)]
            out = torch.full((g.nfiles,), -7, dtype=torch.int64, device=tdev)
            sc.count(b - a, mine[0].data_ptr(), mine[1].data_ptr(), mine[2].data_ptr(), 0, out.data_ptr(), st)
            stream.synchronize()
            res[f"step{epoch}"] = out.cpu().numpy()
        sc.status()
        sc.close()
    g.close()
    np.savez(os.path.join(out_dir, f"dispatch{rank}.npz"), ok=np.array([ok]), **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_peer_memory_query_dispatch(workdir, world):
    """igd_b200_dispatcher_*: the partition by destination rank fused with its all-to-all over peer memory, five
    epochs in a row (both matrix slots and the one receive buffer reused; a tiny batch, an empty one), against the
    numpy model of the plan, and the shards searched over their tile ranges against the oracle.  On a one-GPU box
    the ranks share the device."""
    import torch.multiprocessing as mp
    stem = os.path.join(workdir, f"dispatch_db{world}")
    make_dense_db(stem, n_files=11, n_per_file=1500, seed=33)
    out_dir = os.path.join(workdir, f"dispatch_out{world}")
    os.makedirs(out_dir, exist_ok=True)
    mp.spawn(_dispatch_rank, args=(world, _free_port(), stem + ".igd", out_dir), nprocs=world, join=True)
    ref = OracleEngine(stem + ".igd")
    for r in range(world):
        z = np.load(os.path.join(out_dir, f"dispatch{r}.npz"))
        assert bool(z["ok"][0])
        for epoch, (seed, n) in enumerate([(301, 30000), (302, 7), (303, 40000), (304, 0), (305, 20000)]):
            q = make_queries(seed, n, ("lognormal", 7.0, 1.5))
            ctg = q.ctg.copy()
            if n > 100:
                ctg[:20] = -1
            known = ctg >= 0
            want, _ = ref.count(ctg[known], q.start[known], q.end[known], 0)
            assert np.array_equal(z[f"hits{epoch}"], want), (r, epoch)
        for epoch, (seed, n) in enumerate([(311, 25000), (312, 0), (313, 40000), (314, 3)]):
            q = make_queries(seed, n, ("lognormal", 7.0, 1.5))
            want, _ = ref.count(q.ctg, q.start, q.end, 0)
            assert np.array_equal(z[f"step{epoch}"], want), (r, "one-call step", epoch)
