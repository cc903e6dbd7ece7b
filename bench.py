#!/usr/bin/env python
"""Benchmark of the IGD batch overlap search hot path (count mode) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3]

One "step" = one pass of the hot path over the whole query batch of the workload (BASELINE.json configs[2] by
default, the one the metric and the target are quoted on: 10,000 files x 100k intervals = 1e9, 10M x 1 kb queries,
count mode; it fits one GPU).  Prints ONE JSON line.  With N GPUs the database is replicated and the SAME 10M
queries (strong scaling), sorted by (contig, start), are cut into N contiguous cost-balanced shards; hits[] is
combined by one exchange kernel over NVLink peer memory (NCCL allreduce as the fallback).

* value  : query intervals/s, whole job, queries already resident in HBM (CUDA events, max over ranks)
* e2e    : the same through the C-ABI call with HOST (pinned) buffers: H2D + kernels + D2H inside
* roofline : the bytes the search kernel reads per launch (descriptors + entries + segments, counted by the
             planner on the device; equal to ncu's dram__bytes within 10 %) / its CUDA-event duration, against
             MEASURED_PEAKS.json's HBM copy bandwidth.  The reference-scan-range figure of SURVEY.md 8(d) is
             reported beside it as `algorithmic_speedup` -- this kernel does not do that work.
* cpu_baseline : the oracle port timed on one host core on a bounded sample of the same workload (reported, not a
             target); parity_vs_oracle compares hits[] of the WHOLE query set (oracle on all host cores)

--impl reference times the unmodified reference binary (oracle/_ref/igd, all host cores, one process per
core over line-splits of the query file) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from igd_b200 import synth, workloads  # noqa: E402

METRIC = "query_intervals_per_s"
UNIT = "queries/s"


def env_int(k, d):
    return int(os.environ.get(k, d))


# stdout carries exactly ONE line, the JSON: whatever libraries print there meanwhile (NCCL announces its
# version on stdout when NCCL_DEBUG is set) goes to stderr instead
_JSON_FD = None


def claim_stdout():
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                     "hw_power_brake": 0x80, "sw_power_cap": 0x4}
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.002)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max,
                "samples": len(self.sm), "reasons": sorted(self.reasons)}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML names as closest to the GPU, before any pinned host buffer is
    allocated: on a two-socket box the host<->device copies of the end-to-end path otherwise cross the
    socket interconnect on some runs (12 MB H2D measured at 21 GB/s instead of 52 GB/s)."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, x in enumerate(words) for b in range(64) if x >> b & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} CPUs near GPU {index}"
    except Exception as e:  # pragma: no cover
        return f"not bound ({type(e).__name__})"
    return "not bound"


def workload_setup(args, rank, world, barrier):
    """DB on disk (built once by rank 0) + the query set mapped to database contig ids."""
    workdir = args.workdir
    if rank == 0:
        t0 = time.time()
        path = workloads.build_db(args.workload, workdir)
        build_s = time.time() - t0
    barrier()
    path = os.path.join(workdir, args.workload + ".igd")
    # strong scaling: the query set is the workload's, whatever the number of GPUs
    q = workloads.queries(args.workload, args.queries or workloads.CONFIGS[args.workload]["n_queries"])
    return path, q, (build_s if rank == 0 else 0.0)


def config_of(args, n_total, world):
    """the `config` object of the JSON line -- the same for both arms"""
    c = workloads.CONFIGS[args.workload]
    return {"workload": f"{args.workload}: {c['n_files']} files x {c['per_file']} intervals (hg38, -b 14), "
                        f"{n_total} queries, count mode",
            "n_files": c["n_files"], "intervals_per_file": c["per_file"], "n_queries": n_total,
            "l2": "256 MiB flush write between steps (untimed); DB image > 126 MB L2",
            "parallelism": (f"db replicated on {world} GPUs; the {n_total} queries cut into {world} contiguous tile ranges of "
                            f"the (contig, start) order by scan-byte prefix sum, on the devices; hits[] combined by one exchange"
                            if world > 1 else "single GPU")}


def oracle_leg(path, ctg, qs, qe, sample, cache_file, full=True):
    """cpu_baseline: the oracle port on ONE host core over a strided sample (also the algorithmic-byte counters);
    then hits[] of the whole query set on all host cores (the parity check of the run)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    n = len(ctg)
    if os.path.exists(cache_file):
        c = json.load(open(cache_file))
        if c.get("n") == n and c.get("sample") == sample and (c.get("hits_full") is not None or not full):
            return c
    odb = O.OracleDB(path)
    m = min(n, sample)
    # a strided sample keeps the (ctg, start) distribution of the whole set
    sel = np.arange(m) * (n // m) if m < n else np.arange(n)
    t0 = time.perf_counter()
    hits, _, st = odb.count(ctg[sel], qs[sel], qe[sel], 0, stats=True)
    dt = time.perf_counter() - t0
    scale = n / m
    out = {"n": n, "sample": sample, "sample_n": int(m), "seconds": dt, "qps": m / dt,
           "stats": st, "scale": scale, "bytes_total": st["bytes"] * scale, "hits_total": st["hits"] * scale,
           "tile_bytes_sample": st["tile_bytes"], "hits_full": None}
    if full:
        cores = os.cpu_count() or 1
        cuts = [n * k // (4 * cores) for k in range(4 * cores + 1)]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:       # orc_count releases the GIL (ctypes) and only reads the database
            parts = list(ex.map(lambda ab: odb.count(np.ascontiguousarray(ctg[ab[0]:ab[1]]),
                                                     np.ascontiguousarray(qs[ab[0]:ab[1]]),
                                                     np.ascontiguousarray(qe[ab[0]:ab[1]]), 0)[0],
                                zip(cuts[:-1], cuts[1:])))
        out["hits_full"] = np.sum(parts, axis=0).astype(np.int64).tolist()
        out["full_seconds"], out["full_cores"] = time.perf_counter() - t0, cores
    odb.close()
    with open(cache_file, "w") as f:
        json.dump(out, f)
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from igd_b200 import api

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the igd_b200 search path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)
    dev = torch.device("cuda", local)

    path, q, build_s = workload_setup(args, rank, world, barrier)
    numa = bind_to_gpu_numa_node(local)          # after the multi-process database build, before any pinned buffer
    t0 = time.time()
    igd = api.IGD(path, device=local)
    open_s = time.time() - t0
    ctg_all = igd.resolve_names(synth.HG38_NAMES)[q.ctg].astype(np.int32)
    qs_all, qe_all = q.start, q.end
    n_total = len(ctg_all)

    # a real (non-default) stream: the library enqueues its kernels on the stream it is handed, and the
    # CUDA events / NCCL collectives below are recorded on the same one
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    stream = side.cuda_stream

    # multi-GPU: DB replicated; every rank starts from ITS 1/N slice of the query set as generated (any order) and
    # the shards -- contiguous tile ranges of the (contig, start) order, balanced by the prefix sum of estimated scan
    # bytes -- are made on the devices (igd_b200_shard_*: per-tile cost, all-reduce, cuts, partition, all-to-all)
    tile_lo, tile_hi = 0, int(igd.ntiles)
    if world > 1:
        from igd_b200.multi import DeviceSharder
        a, b = n_total * rank // world, n_total * (rank + 1) // world
        ctg_l, qs_l, qe_l = ctg_all[a:b], qs_all[a:b], qe_all[a:b]
        # the query exchange by the library's own kernels over NVLink peer memory (igd_b200_dispatcher_*: partition
        # fused with its all-to-all); IGD_B200_NCCL_SHARD=1 keeps the partition + NCCL all-to-all
        sharder = DeviceSharder(igd, dev, peer_cap=0 if os.environ.get("IGD_B200_NCCL_SHARD") else n_total)
        u_ctg, u_qs, u_qe = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (ctg_l, qs_l, qe_l))
        d_ctg, d_qs, d_qe, tile_lo, tile_hi = sharder.shard(u_ctg, u_qs, u_qe)
        d_ctg, d_qs, d_qe = d_ctg.clone(), d_qs.clone(), d_qe.clone()
        torch.cuda.synchronize()
    else:
        ctg_l, qs_l, qe_l = ctg_all, qs_all, qe_all
        if args.shard_of > 1:
            # profiling aid: the work of ONE rank of an N-GPU run on a single GPU (the shard cut on the host like
            # tests/test_gpu_shard.py's reference) -- what ncu can look at, since it must not run multi-rank commands
            hdr = workloads.read_igd_header(path)
            s_ctg, s_qs, s_qe, bounds = workloads.sort_and_shard(hdr, ctg_all, qs_all, qe_all, args.shard_of)
            a, b = bounds[args.shard_rank], bounds[args.shard_rank + 1]
            ctg_l, qs_l, qe_l = s_ctg[a:b], s_qs[a:b], s_qe[a:b]
            nbp, tb = hdr["nbp"], hdr["tbase"]
            first = tb[ctg_l] + qs_l // nbp
            last = tb[ctg_l] + np.minimum((qe_l.astype(np.int64) - 1) // nbp, hdr["ntile"][ctg_l] - 1)
            tile_lo, tile_hi = int(first.min()), int(max(last.max(), first.max())) + 1
            args.no_parity = True
        d_ctg = torch.from_numpy(np.ascontiguousarray(ctg_l)).to(dev)
        d_qs = torch.from_numpy(np.ascontiguousarray(qs_l)).to(dev)
        d_qe = torch.from_numpy(np.ascontiguousarray(qe_l)).to(dev)
    n_local = int(d_ctg.shape[0])                   # queries of this rank's shard
    n_slice = len(ctg_l)                            # queries this rank uploads in the end-to-end leg

    d_hits = torch.zeros(igd.nfiles, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    # the one exchange of the path: hits[] summed over the ranks by the library's own kernel over NVLink peer
    # memory (igd_b200_reducer_*); NCCL's all-reduce if that cannot be set up on every rank (or on request)
    reducer = None
    if world > 1:
        ok = 0
        if not os.environ.get("IGD_B200_NCCL_REDUCE"):
            try:
                reducer = api.HitsReducer(local, igd.nfiles)
                ok = 1
            except Exception as e:  # pragma: no cover
                print(f"[bench] rank {rank}: peer-memory reducer unavailable ({e}); NCCL all-reduce", file=sys.stderr)
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0 and reducer is not None:
            reducer.close()
            reducer = None

    # N > 1: a rank counts into d_part, and the exchange kernel writes the sum to d_hits and leaves d_part zeroed for
    # the next step (igd_b200_reducer_allreduce_into) -- no separate launch to clear the counters
    d_part = torch.zeros(igd.nfiles, dtype=torch.int64, device=dev) if world > 1 else d_hits

    def combine():
        if reducer is not None:
            reducer.allreduce_into(d_part.data_ptr(), d_hits.data_ptr(), stream)
        else:
            d_hits.copy_(d_part)
            d_part.zero_()
            dist.all_reduce(d_hits)

    def step_device():
        if world == 1:
            d_hits.zero_()
        if n_local:
            igd.count_device_tiles(n_local, d_ctg.data_ptr(), d_qs.data_ptr(), d_qe.data_ptr(), 0, d_part.data_ptr(),
                                   stream, tile_lo, tile_hi)
            igd.count_device_finish()
        if world > 1:
            combine()

    # pinned host copies for the end-to-end path
    h_ctg = torch.from_numpy(np.ascontiguousarray(ctg_l)).pin_memory()
    h_qs = torch.from_numpy(np.ascontiguousarray(qs_l)).pin_memory()
    h_qe = torch.from_numpy(np.ascontiguousarray(qe_l)).pin_memory()
    h_hits = torch.zeros(igd.nfiles, dtype=torch.int64).pin_memory()
    e_ctg, e_qs, e_qe = (torch.empty(n_slice, dtype=torch.int32, device=dev) for _ in range(3))

    def step_e2e():
        if world == 1:
            # the reference-facing C-ABI call: host buffers in, hits[] out (igd_b200_count)
            hits = np.zeros(igd.nfiles, np.int64)
            igd.count_ptr(n_local, h_ctg.data_ptr(), h_qs.data_ptr(), h_qe.data_ptr(), 0, hits)
            return hits
        # this rank's slice of the unsorted query set: H2D, shard plan + exchange on the devices, search, combine
        e_ctg.copy_(h_ctg, non_blocking=True)
        e_qs.copy_(h_qs, non_blocking=True)
        e_qe.copy_(h_qe, non_blocking=True)
        s_ctg, s_qs, s_qe, lo, hi = sharder.shard(e_ctg, e_qs, e_qe)
        if s_ctg.shape[0]:
            igd.count_device_tiles(int(s_ctg.shape[0]), s_ctg.data_ptr(), s_qs.data_ptr(), s_qe.data_ptr(), 0,
                                   d_part.data_ptr(), stream, lo, hi)
            igd.count_device_finish()
        combine()
        h_hits.copy_(d_hits, non_blocking=True)
        torch.cuda.synchronize()
        return h_hits.numpy()

    # ---- warm-up ---------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        step_device()
    torch.cuda.synchronize()
    if world > 1:
        # did every exchange meet all its peers?  if not (anywhere), everybody goes back to NCCL and warms up again
        bad = torch.tensor([reducer.status() if reducer is not None else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if reducer is not None and int(bad.item()) != 0:
            print(f"[bench] rank {rank}: peer-memory exchange failed its check; NCCL all-reduce", file=sys.stderr)
            reducer.close()
            reducer = None
            for _ in range(3):
                step_device()
            torch.cuda.synchronize()
    hits_dev = d_hits.cpu().numpy().copy()

    # ---- timed: device-resident inputs -----------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = igd.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    barrier()
    torch.cuda.synchronize()
    for i in range(args.steps):
        flush.fill_(i & 0xff)               # evict the previous step's tiles from L2 (not timed)
        ev[i][0].record()
        step_device()
        ev[i][1].record()
        torch.cuda.synchronize()
        kernel_ms.append(igd.last_search_kernel_ms())
    barrier()
    torch.cuda.synchronize()
    launches = igd.launch_count() - launches0
    plan = igd.last_plan()
    # one more (untimed) step with CUDA events around each of the four launches
    igd.stage_timing(True)
    flush.fill_(3)
    step_device()
    torch.cuda.synchronize()
    stages = igd.last_stage_ms()
    igd.stage_timing(False)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))

    # ---- timed: end to end from host buffers --------------------------------------------------------
    # one call at a time: H2D, then the search, then D2H, each step on its own (N = 1: also reported as
    # e2e.single_call, the latency of one igd_b200_count)
    for _ in range(3):
        hits_e2e = step_e2e()
    barrier()
    torch.cuda.synchronize()
    e2e_s = 0.0
    for i in range(args.steps):
        flush.fill_(i & 0xff)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        hits_e2e = step_e2e()
        e2e_s += time.perf_counter() - t0
    e2e_split = igd.last_timing() if world == 1 else None
    # N = 1: the same K batches as a STREAM through igd_b200_count_submit / _wait -- two in flight, the H2D of batch
    # k + 1 under the search of batch k; every step still copies its 12 B per query in and its hits[] out inside the
    # timed region (no L2 flush between the steps: each one reads several GB of the database, far beyond the L2)
    e2e_single, e2e_phases, h2d_under = None, None, None
    if world > 1:
        # N > 1: the same K steps as a stream -- the H2D of step k + 1 (this rank's 1/N of the queries, on a copy stream
        # into the other of two input buffers) under the shard plan, exchange and search of step k
        copy_st = torch.cuda.Stream(device=dev)
        step_c = None
        if reducer is not None and sharder.peer is not None and not os.environ.get("IGD_B200_PY_STEP"):
            step_c = api.ShardedCounter(igd, local, n_total)
        e_in = [[torch.empty(n_slice, dtype=torch.int32, device=dev) for _ in range(3)] for _ in range(2)]
        h_out = [torch.zeros(igd.nfiles, dtype=torch.int64).pin_memory() for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]

        h2d_ev = []                                     # (begin, end) of every upload, on the copy stream

        d_out2 = [torch.zeros(igd.nfiles, dtype=torch.int64, device=dev) for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        done_valid = [False, False]

        def upload(k):
            with torch.cuda.stream(copy_st):
                if step_c is not None and k >= 2 and done_valid[k & 1]:
                    copy_st.wait_event(ev_done[k & 1])  # step k - 2 (the last reader of these buffers) has finished
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record(copy_st)
                for dst, src in zip(e_in[k & 1], (h_ctg, h_qs, h_qe)):
                    dst.copy_(src, non_blocking=True)
                b_.record(copy_st)
                h2d_ev.append((a_, b_))
                ev_in[k & 1].record(copy_st)

        def stream_steps(k_steps):
            upload(0)
            for k in range(k_steps):
                if k + 1 < k_steps:
                    upload(k + 1)                       # (its buffers were last read by step k - 1)
                side.wait_event(ev_in[k & 1])
                if step_c is not None:
                    # the whole step as ONE C call (igd_b200_sharded_count): cost, its sum over the ranks, cuts, query
                    # exchange, search, sum of hits[] -- the library's kernels over peer memory, no host round trip.
                    # The host reads a step's result one step late (two result buffers), so step k + 1 is enqueued
                    # while step k still runs and the GPU goes from one step straight into the next.
                    x = e_in[k & 1]
                    step_c.count(n_slice, x[0].data_ptr(), x[1].data_ptr(), x[2].data_ptr(), 0, d_out2[k & 1].data_ptr(), stream)
                    h_out[k & 1].copy_(d_out2[k & 1], non_blocking=True)
                    ev_done[k & 1].record(side)
                    done_valid[k & 1] = True
                    if k >= 1:
                        ev_done[(k - 1) & 1].synchronize()      # the result of step k - 1 is on the host
                    if k + 1 == k_steps:
                        ev_done[k & 1].synchronize()
                    continue
                if sharder.peer is not None:
                    # no host round trip between the exchange and the search: the shard's shape stays on the device
                    p_c, p_s, p_e, p_shape = sharder.shard_async(*e_in[k & 1])
                    igd.count_device_shaped(n_total, p_shape, p_c, p_s, p_e, 0, d_part.data_ptr(), stream)
                    igd.count_device_finish()
                else:
                    s_ctg, s_qs, s_qe, lo, hi = sharder.shard(*e_in[k & 1])
                    if s_ctg.shape[0]:
                        igd.count_device_tiles(int(s_ctg.shape[0]), s_ctg.data_ptr(), s_qs.data_ptr(), s_qe.data_ptr(), 0,
                                               d_part.data_ptr(), stream, lo, hi)
                        igd.count_device_finish()
                combine()
                h_out[k & 1].copy_(d_hits, non_blocking=True)
                side.synchronize()                      # the step's result is on the host
            return h_out[(k_steps - 1) & 1].numpy()

        stream_steps(3)
        # one more (untimed) step, one at a time, with CUDA events at the phase boundaries
        sharder.marks = []
        for _ in range(2):
            sharder.marks.clear()
            torch.cuda.synchronize()
            barrier()
            sharder._mark("h2d_begin")
            step_e2e()
            sharder._mark("end")
        torch.cuda.synchronize()
        mk = sharder.marks
        e2e_phases = {mk[i + 1][0]: round(mk[i][1].elapsed_time(mk[i + 1][1]), 4) for i in range(len(mk) - 1)}
        sharder.marks = None
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h2d_ev.clear()
        out_last = stream_steps(args.steps)
        e2e_stream_s = time.perf_counter() - t0
        torch.cuda.synchronize()
        h2d_under = float(np.median([a_.elapsed_time(b_) for a_, b_ in h2d_ev[1:]])) if len(h2d_ev) > 1 else None
        assert np.array_equal(out_last, np.asarray(hits_e2e)), "pipelined and step-by-step end-to-end paths disagree"
        hh = torch.tensor([h2d_under or 0.0], dtype=torch.float64, device=dev)
        hall = [torch.empty_like(hh) for _ in range(world)]
        dist.all_gather(hall, hh)
        h2d_under = [round(float(x.item()), 4) for x in hall]     # every rank's: the slowest upload paces the stream
        if sharder.peer is not None:
            sharder.peer.status()                       # every dispatch met all its peers, no receive buffer overflowed
        if step_c is not None:
            step_c.status()
            barrier()
            step_c.close()
        e2e_single = {"value": None, "unit": UNIT, "ms_per_step": e2e_s / args.steps * 1e3,
                      "api": "the same steps one at a time (H2D, then shard + search + exchange, then D2H)"}
        e2e_s = e2e_stream_s
    if world == 1:
        def stream_steps(k):
            outs = [np.zeros(igd.nfiles, np.int64) for _ in range(k)]
            ptrs = (n_local, h_ctg.data_ptr(), h_qs.data_ptr(), h_qe.data_ptr(), 0)
            igd.count_submit(0, *ptrs)
            for j in range(1, k):
                igd.count_submit(j & 1, *ptrs)
                igd.count_wait((j - 1) & 1, outs[j - 1])
            igd.count_wait((k - 1) & 1, outs[k - 1])
            return outs
        stream_steps(3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        outs = stream_steps(args.steps)
        e2e_stream_s = time.perf_counter() - t0
        for o in outs:
            assert np.array_equal(o, hits_e2e), "pipelined and one-call end-to-end paths disagree"
        e2e_single = {"value": n_total / (e2e_s / args.steps), "unit": UNIT, "ms_per_step": e2e_s / args.steps * 1e3,
                      "api": "igd_b200_count (one synchronous call per step)"}
        e2e_s = e2e_stream_s
    sampler.stop_flag = True
    sampler.join()

    # ---- query-file ingest: search() on a BED file of this rank's queries (text parsed on the GPU) ------------
    file_s = None
    if world == 1:
        nf = min(n_total, 1_000_000)                        # (a 1M-line BED file: 29 MB of text)
        qbed = os.path.join(args.workdir, f"{args.workload}_q{nf}.bed")
        if not os.path.exists(qbed):
            synth.write_bed(qbed, synth.Intervals(q.ctg[:nf], q.start[:nf], q.end[:nf], q.value[:nf]))
        hits_file, _ = igd.search_n(qbed)
        hits_arr, _ = igd.count(ctg_all[:nf], qs_all[:nf], qe_all[:nf], 0)
        assert np.array_equal(hits_file, hits_arr), "file ingest and array paths disagree"
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            igd.search_n(qbed)
            ts.append(time.perf_counter() - t0)
        file_s = float(np.median(ts))
        file_bytes = os.path.getsize(qbed)
        file_n = nf

    if world > 1:
        t = torch.tensor([total_ms, e2e_s, float(np.mean(kernel_ms)), e2e_single["ms_per_step"]], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, kmean, e2e_single["ms_per_step"] = (float(x) for x in t.cpu())
        e2e_single["value"] = n_total / (e2e_single["ms_per_step"] * 1e-3)
        ln = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(ln)
        launches = int(ln.item())
    else:
        kmean = float(np.mean(kernel_ms))
    assert np.array_equal(np.asarray(hits_e2e), hits_dev), "e2e and device-resident paths disagree"

    per_rank = None
    if world > 1:
        mine = torch.tensor([float(np.mean(kernel_ms)), float(np.mean(step_ms)), float(n_local), float(plan["items"])]
                            + [float(stages[k]) for k in ("seg_count", "plan", "seg_scatter", "search")],
                            dtype=torch.float64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {k: [round(float(t[i]), 4) for t in allr] for i, k in enumerate(
            ("kernel_ms", "step_ms", "queries", "items", "seg_count_ms", "plan_ms", "seg_scatter_windows_ms", "search_ms"))}
    if world > 1:
        pl = torch.tensor([plan["own_bytes"]], dtype=torch.float64, device=dev)
        dist.all_reduce(pl, op=dist.ReduceOp.MAX)
        own_bytes = float(pl.item())                        # of the rank with the longest launch
    else:
        own_bytes = float(plan["own_bytes"])
    if rank == 0:
        ms_per_step = total_ms / args.steps
        overlaps = int(hits_dev.sum())
        peak, peak_src = peaks()
        cache = os.path.join(args.workdir, f"{args.workload}.oracle_{n_total}_{args.cpu_sample}.json")
        orc = oracle_leg(path, ctg_all, qs_all, qe_all, args.cpu_sample, cache, full=not args.no_parity)
        parity = None
        if orc.get("hits_full") is not None:
            parity = bool(np.array_equal(np.array(orc["hits_full"], np.int64), hits_dev))
        if args.shard_of > 1:
            n_total = n_local
        achieved = own_bytes / (kmean * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            tr = json.load(open(tp)).get(args.workload)
            traffic = tr.get(str(world)) if isinstance(tr, dict) else None
        cfg = config_of(args, n_total, world)
        if args.shard_of > 1:
            cfg["parallelism"] = (f"PROFILING AID: the shard of rank {args.shard_rank} of {args.shard_of} on one GPU "
                                  f"({n_local} queries, tiles [{tile_lo}, {tile_hi})); not a benchmark line")
        if world > 1:
            cfg["parallelism"] += (" kernel per rank over NVLink peer memory (igd_b200_reducer)" if reducer is not None
                                   else " NCCL int64 allreduce")
        line = {
            "metric": METRIC, "value": n_total / (ms_per_step * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u16",
            "data": "synthetic",
            "overlaps_per_s": overlaps / (ms_per_step * 1e-3), "overlaps_per_step": overlaps,
            "config": cfg,
            "run": {"n_records": int(igd.nrecords), "n_tiles": int(igd.ntiles), "db_hbm_bytes": int(igd.device_bytes()),
                    "queries_rank0": n_local, "db_build_s": round(build_s, 1), "db_open_s": round(open_s, 2),
                    "parity_seconds": orc.get("full_seconds"), "parity_cores": orc.get("full_cores")},
            "e2e": {"value": n_total / (e2e_s / args.steps), "unit": UNIT, "ms_per_step": e2e_s / args.steps * 1e3,
                    "h2d_bytes_per_step": 12 * n_total, "d2h_bytes_per_step": 8 * igd.nfiles * world,
                    "single_call": e2e_single, "phases_ms_rank0_one_step": e2e_phases,
                    "h2d_ms_per_rank_under_the_previous_step": h2d_under,
                    "device_split_ms_last_step": e2e_split,
                    # the H2D leg runs at whatever the box's PCIe path gives at that moment: 16 to 54 GB/s were
                    # seen on this pool's boxes for the same pinned buffers (profiles/microbench/h2d_*.py)
                    "h2d_gbs_last_step": (12 * n_total / (e2e_split["h2d_ms"] * 1e-3) / 1e9
                                          if e2e_split and e2e_split["h2d_ms"] > 0 else None),
                    "cpu_binding": numa,
                    "api": "igd_b200_count_submit / igd_b200_count_wait (C-ABI, pinned host buffers; the K steps as a stream, "
                           "two batches in flight: H2D of step k + 1 under the search of step k)" if world == 1 else
                           "pinned H2D of this rank's 1/N of the queries + igd_b200_shard_* (cost, all-reduce, cuts) + "
                           "igd_b200_dispatch (partition fused with its all-to-all over NVLink peer memory) + "
                           "igd_b200_count_device_shaped + hits[] exchange -- as ONE C call per step, igd_b200_sharded_count -- "
                           "+ D2H; the K steps as a stream: "
                           "the H2D of step k + 1 on a copy stream under step k"},
            "file_ingest": None if file_s is None else {
                "value": file_n / file_s, "unit": UNIT, "ms": file_s * 1e3, "file_bytes": file_bytes, "n_queries": file_n,
                "api": "search(igd, qFile) -- BED text (warm page cache) -> pinned chunks -> H2D -> parsed on the "
                       "GPU -> count"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         # the counted bytes include re-reads that the L2 serves (the second chunk of a tile with more than
                         # 64 segments reads its entries again), so `frac` can pass 1; by DRAM traffic (ncu, same kernel):
                         "frac_by_dram_traffic": (traffic / (kmean * 1e-3) / 1e9 / peak) if traffic else None,
                         "kernel": "k_count_rank (count-mode search, 16-bit START / END entries)",
                         "kernel_ms": kmean, "stages_ms_rank0": stages, "per_rank": per_rank,
                         "bytes_per_launch": own_bytes,
                         "bytes_definition": "what the search kernel reads per launch: 32 B per item descriptor + 4 B per "
                                             "(entry, item) + 4 B per segment, counted by the planner on the device "
                                             "(longest-running rank)",
                         "plan_rank0": plan, "peak_source": peak_src,
                         # SURVEY.md 8(d)'s figure -- the bytes the REFERENCE's algorithm would move for these queries
                         # (oracle-instrumented scan ranges) -- is not what this kernel does; as a ratio only:
                         "algorithmic_bytes_reference_definition": orc["bytes_total"],
                         "algorithmic_speedup": orc["bytes_total"] / (own_bytes * world) if own_bytes else None},
            "cpu_baseline": {"value": orc["qps"], "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": f"{orc['sample_n']} of {n_total} queries of the same workload (strided), "
                                       f"oracle/igd_oracle.c, {orc['seconds']:.1f} s"},
            "parity_vs_oracle": parity,
        }
        emit(line)
    if reducer is not None:
        dist.barrier()
        reducer.close()
    if world > 1:
        dist.barrier()
        sharder.close()
    igd.close()
    if world > 1:
        dist.destroy_process_group()


def run_modes(args):
    """--mode f | m: the other two sharded modes of the path as bench lines (cfg-4 / cfg-5 shapes on the cfg-2
    database by default).  f: `value` = queries/s with queries and lists resident in HBM (igd_b200_full_device; N > 1:
    shards, gather and order restoration on the devices), `e2e` = host arrays in, host lists out.  m: `value` =
    overlapping record pairs/s of the whole self-join (N > 1: tile ranges + one uint32 all-reduce), matrix to the host
    inside `e2e`."""
    import torch
    import torch.distributed as dist
    from igd_b200 import api, multi
    world, rank, local = env_int("WORLD_SIZE", 1), env_int("RANK", 0), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the igd_b200 search path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)
    wl = args.workload if args.workload != "cfg3" else "cfg2"
    if rank == 0:
        workloads.build_db(wl, args.workdir)
    barrier()
    path = os.path.join(args.workdir, wl + ".igd")
    igd = api.IGD(path, device=local)
    q = workloads.queries(wl, args.queries or workloads.CONFIGS[wl]["n_queries"])
    ctg = igd.resolve_names(synth.HG38_NAMES)[q.ctg].astype(np.int32)
    qs, qe, n = q.start, q.end, len(q)
    sh = multi.ShardedSearch(igd, workloads.read_igd_header(path), comm_device=str(dev)) if world > 1 else None
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = igd.launch_count()
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        barrier()
        torch.cuda.synchronize()
        ts = []
        for _ in range(steps):
            barrier()
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return float(np.mean(ts)), out

    steps, warm = args.steps, max(args.warmup, 3)
    if args.mode == "f":
        d = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (ctg, qs, qe)]
        if world == 1:
            dev_s, tot = timed(lambda: igd.full_device(n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), side.cuda_stream)[0],
                               steps, warm)
            e2e_s, res = timed(lambda: igd.full(ctg, qs, qe, copy=False), steps, warm)
            total = int(res[0][-1])
            assert tot == total
        else:
            e2e_s, res = timed(lambda: sh.full(ctg, qs, qe, dst=0), steps, warm)
            dev_s, total = e2e_s, (int(res[0][-1]) if rank == 0 else 0)
        unit, metric, units = UNIT, METRIC, n
        extra = {"overlaps_per_step": total, "list_bytes_per_step": 12 * total}
        d2h = 12 * total + 8 * (n + 1)
    else:
        v = 0
        if world == 1:
            m = torch.zeros(igd.nfiles * igd.nfiles, dtype=torch.int32, device=dev)
            dev_s, _ = timed(lambda: igd.hitsmap_tiles_device(v, 0, igd.ntiles, m.data_ptr(), side.cuda_stream), steps, warm)
            e2e_s, hm = timed(lambda: igd.hitsmap(v), steps, warm)
        else:
            e2e_s, hm = timed(lambda: sh.hitsmap(v), steps, warm)
            dev_s = e2e_s
        pairs = int(hm.astype(np.uint64).sum())
        unit, metric, units = "record pairs/s", "hitsmap_pairs_per_s", pairs
        extra = {"pairs_per_step": pairs, "matrix_bytes": 4 * igd.nfiles * igd.nfiles}
        d2h = 4 * igd.nfiles * igd.nfiles
    if world > 1:
        t = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, e2e_s = (float(x) for x in t.cpu())
    sampler.stop_flag = True
    sampler.join()
    if rank == 0:
        c = workloads.CONFIGS[wl]
        emit({"metric": metric, "mode": "-" + args.mode, "value": units / dev_s, "unit": unit, "n_gpus": world, "steps": steps,
              "warmup": warm, "ms_per_step": dev_s * 1e3, "higher_is_better": True, "scaling": "strong",
              "vs_baseline": None, "dtype": "int32", "data": "synthetic",
              "config": {"workload": f"{wl}: {c['n_files']} files x {c['per_file']} intervals (hg38, -b 14), "
                                     + (f"{n} queries, -f" if args.mode == "f" else "-m (database self-join)"),
                         "timing": "host clock around the call + device synchronize (the calls end with the result in place)"},
              "e2e": {"value": units / e2e_s, "unit": unit, "ms_per_step": e2e_s * 1e3,
                      "h2d_bytes_per_step": 12 * n if args.mode == "f" else 0, "d2h_bytes_per_step": d2h},
              "gpu_launches": int(igd.launch_count() - launches0), "clocks": sampler.summary(), **extra})
    igd.close()
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The reference's own CPU implementation on the box's host cores (rank 0 only)."""
    if env_int("RANK", 0) != 0:
        return
    from oracle import oracle as O
    path = workloads.build_db(args.workload, args.workdir)
    q = workloads.queries(args.workload, args.queries or workloads.CONFIGS[args.workload]["n_queries"])   # our arm's set
    n_total = len(q)
    cores = os.cpu_count() or 1
    n = min(n_total, args.ref_sample)
    qdir = os.path.join(args.workdir, f"{args.workload}_refq_{n}_{cores}")
    if not os.path.isdir(qdir):
        os.makedirs(qdir)
        sel = np.arange(n) * (n_total // n)
        for c in range(cores):
            part = sel[c::cores]
            synth.write_bed(os.path.join(qdir, f"q{c:03d}.bed"),
                            synth.Intervals(q.ctg[part], q.start[part], q.end[part], q.value[part]))
    # the same sample sorted by (contig, start): the reference keeps ONE tile cached (src/igd_search.c:469-476), so
    # sorted queries re-use it -- reported beside the headline, which takes the queries as generated
    sdir = qdir + "_sorted"
    if not os.path.isdir(sdir):
        os.makedirs(sdir)
        sel = np.arange(n) * (n_total // n)
        sel = sel[np.lexsort((q.start[sel], q.ctg[sel]))]
        for c in range(cores):
            part = sel[c * len(sel) // cores:(c + 1) * len(sel) // cores]
            synth.write_bed(os.path.join(sdir, f"q{c:03d}.bed"),
                            synth.Intervals(q.ctg[part], q.start[part], q.end[part], q.value[part]))
    kind = "reference" if O.have_ref() else "port"

    def one_step(qdir=qdir):
        t0 = time.perf_counter()
        if kind == "reference":
            procs = [subprocess.Popen([O.REF_BIN, "search", path, "-q", os.path.join(qdir, f"q{c:03d}.bed")],
                                      stdout=subprocess.DEVNULL) for c in range(cores)]
            for p in procs:
                p.wait()
        else:
            code = ("import sys; sys.path.insert(0, %r); from oracle import oracle as O; "
                    "db=O.OracleDB(sys.argv[1]); q=O.Queries(sys.argv[2]); db.count(db.resolve(q), q.qs, q.qe)" % ROOT)
            procs = [subprocess.Popen([sys.executable, "-c", code, path, os.path.join(qdir, f"q{c:03d}.bed")])
                     for c in range(cores)]
            for p in procs:
                p.wait()
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        one_step()
    total = sum(one_step() for _ in range(args.steps))
    sec = total / args.steps
    val = n / sec
    one_step(sdir)
    sorted_val = n / min(one_step(sdir) for _ in range(2))
    sample = f"{n} of {n_total} queries per step, split over {cores} processes of oracle/_ref/igd search -q" \
        if kind == "reference" else f"{n} of {n_total} queries per step, {cores} processes of the oracle port"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config_of(args, n_total, max(args.gpus, 1)),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                             "value_sorted_queries": sorted_val},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def default_workdir():
    """the 17 GB database of cfg3 is written once per box: to /dev/shm when it is a roomy tmpfs (the file is read
    back through the page cache either way), else /tmp"""
    try:
        st = os.statvfs("/dev/shm")
        if st.f_bavail * st.f_frsize > 64 << 30:
            return "/dev/shm/igd_b200_bench"
    except OSError:
        pass
    return "/tmp/igd_b200_bench"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(workloads.CONFIGS))
    ap.add_argument("--queries", type=int, default=None, help="override the (total) number of queries")
    ap.add_argument("--workdir", default=os.environ.get("IGD_B200_WORKDIR", default_workdir()))
    ap.add_argument("--cpu-sample", type=int, default=1_000_000, help="queries in the one-core cpu_baseline sample")
    ap.add_argument("--no-parity", action="store_true", help="skip the whole-set oracle run (profiling runs)")
    ap.add_argument("--mode", default="count", choices=["count", "f", "m"],
                    help="count (the contract line); f / m: the other sharded modes, on cfg2 unless --workload says otherwise")
    ap.add_argument("--shard-of", type=int, default=1, help="profiling aid (N = 1 only): run the shard one rank of this many would get")
    ap.add_argument("--shard-rank", type=int, default=0)
    ap.add_argument("--ref-sample", type=int, default=1_000_000, help="queries per step of --impl reference")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode != "count":
        run_modes(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
