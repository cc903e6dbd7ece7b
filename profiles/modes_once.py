"""one call of each of the other modes on the cfg-2 database (for profiling runs): -f, -m, -s, file ingest"""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from igd_b200 import api, synth, workloads
wd = os.environ.get("IGD_B200_WORKDIR", "/dev/shm/igd_b200_bench" if os.path.isdir("/dev/shm") else "/tmp/igd_b200_bench")
path = workloads.build_db("cfg2", wd)
g = api.IGD(path)
q = workloads.queries("cfg2")
ctg = g.resolve_names(synth.HG38_NAMES)[q.ctg].astype(np.int32)
for _ in range(2):
    g.full(ctg, q.start, q.end, copy=False)
    g.count(ctg, q.start, q.end, 500)
hq = synth.gen_hotspot_queries(synth.seed_for(5, synth.QUERY_FILE_INDEX), 10000, 1_000_000)
qbed = os.path.join(wd, "cfg5_q10000.bed")
synth.write_bed(qbed, hq)
for _ in range(2):
    g.seqpare(qbed)
    g.hitsmap(0)
qb = os.path.join(wd, "cfg2_q1000000.bed")
if not os.path.exists(qb):
    synth.write_bed(qb, q)
for _ in range(2):
    g.search_n(qb)
print("done")
