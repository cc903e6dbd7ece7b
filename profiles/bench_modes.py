#!/usr/bin/env python
"""Timings of the other search modes of the path (BASELINE.json configs 4 and 5) on one B200, through the C-ABI:
  -v 500 count, -f full overlaps (1M x 1 kb queries, cfg-2 database), -s Seqpare and count on 10k hotspot 1 Mb
  queries, -m hitsmap on the 200-file database.  Prints one JSON object; run under gpurun, results are kept in
  profiles/.  (bench.py stays the contract benchmark of the count path.)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from igd_b200 import api, synth, workloads  # noqa: E402


def best(f, n=3):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        r = f()
        ts.append(time.perf_counter() - t0)
    return min(ts), r


def main():
    wd = os.environ.get("IGD_B200_WORKDIR", "/tmp/igd_b200_bench")
    out = {}
    path = workloads.build_db("cfg2", wd)
    g = api.IGD(path)
    q = workloads.queries("cfg2")
    ctg = g.resolve_names(synth.HG38_NAMES)[q.ctg].astype(np.int32)
    # cfg 4: -v 500 (count with the signal filter) and -f
    g.count(ctg, q.start, q.end, 500)
    t, (hits, n) = best(lambda: g.count(ctg, q.start, q.end, 500))
    out["cfg4_v500_count"] = {"queries": len(ctg), "overlaps": int(n), "ms": t * 1e3, "queries_per_s": len(ctg) / t,
                              "kernel_ms": g.last_search_kernel_ms(), "plan": g.last_plan()}
    g.full(ctg[:1000], q.start[:1000], q.end[:1000])
    g.full(ctg, q.start, q.end, copy=False)          # first call pins the result arena (one-time, like open)
    t, r = best(lambda: g.full(ctg, q.start, q.end, copy=False), 3)
    total = int(r[0][-1])
    out["cfg4_full"] = {"queries": len(ctg), "overlaps": total, "ms": t * 1e3, "overlaps_per_s": total / t,
                        "d2h_bytes": 12 * total + 8 * (len(ctg) + 1), "split_ms": g.last_timing(),
                        "note": "host arrays out (pinned arena): includes the D2H of the lists"}
    # cfg 5: 10k hotspot queries of 1 Mb
    hq = synth.gen_hotspot_queries(synth.seed_for(5, synth.QUERY_FILE_INDEX), 10000, 1_000_000)
    hctg = g.resolve_names(synth.HG38_NAMES)[hq.ctg].astype(np.int32)
    g.count(hctg, hq.start, hq.end, 0)
    t, (hits, n) = best(lambda: g.count(hctg, hq.start, hq.end, 0))
    out["cfg5_hotspot_count"] = {"queries": len(hctg), "overlaps": int(n), "ms": t * 1e3, "overlaps_per_s": n / t,
                                 "kernel_ms": g.last_search_kernel_ms(), "plan": g.last_plan()}
    g.close()
    # -s and -m on the 200-file database (the Seqpare candidate list of cfg-2 x 1 Mb queries is 1e10 pairs)
    mpath = workloads.build_db("mini", wd)
    m = api.IGD(mpath)
    sq = synth.gen_hotspot_queries(synth.seed_for(5, synth.QUERY_FILE_INDEX), 2000, 1_000_000)
    qbed = os.path.join(wd, "cfg5_q2000.bed")
    synth.write_bed(qbed, sq)
    t, sm = best(lambda: m.seqpare(qbed), 2)
    out["cfg5_seqpare_mini"] = {"queries": len(sq), "files": m.nfiles, "ms": t * 1e3, "sm_max": float(sm.max())}
    t, hm = best(lambda: m.hitsmap(0), 2)
    out["cfg5_hitsmap_mini"] = {"records": int(m.nrecords), "files": m.nfiles, "pairs": int(hm.sum(dtype=np.uint64)),
                                "ms": t * 1e3, "pairs_per_s": float(hm.sum(dtype=np.uint64)) / t}
    m.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
