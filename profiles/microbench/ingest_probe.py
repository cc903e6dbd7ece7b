import os, sys, time
sys.path.insert(0, ".")
import numpy as np
from igd_b200 import api, synth, workloads
wd = os.environ.get("IGD_B200_WORKDIR", "/tmp/igd_b200_bench")
p = workloads.build_db("cfg2", wd)
q = workloads.queries("cfg2")
qbed = os.path.join(wd, "probe_q.bed")
synth.write_bed(qbed, q)
g = api.IGD(p)
for i in range(4):
    t = time.perf_counter(); g.search_n(qbed); print("search_n %.3f ms" % ((time.perf_counter() - t) * 1e3), file=sys.stderr)
