import json, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from igd_b200 import api, synth, workloads
wd = os.environ.get("IGD_B200_WORKDIR", "/tmp/igd_b200_bench")
path = workloads.build_db("cfg2", wd)
g = api.IGD(path)
out = {}
hq = synth.gen_hotspot_queries(synth.seed_for(5, synth.QUERY_FILE_INDEX), 10000, 1_000_000)
qbed = os.path.join(wd, "cfg5_q10000.bed")
synth.write_bed(qbed, hq)
for rep in range(2):
    t0 = time.perf_counter()
    try:
        sm = g.seqpare(qbed)
        out["seqpare_cfg2_10k_1Mb"] = {"ms": (time.perf_counter() - t0) * 1e3, "sm_max": float(sm.max()), "sm_sum": float(sm.sum())}
    except Exception as e:
        out["seqpare_cfg2_10k_1Mb"] = {"error": str(e)}
        break
for rep in range(2):
    t0 = time.perf_counter()
    hm = g.hitsmap(0)
    out["hitsmap_cfg2"] = {"ms": (time.perf_counter() - t0) * 1e3, "pairs": int(hm.sum(dtype=np.uint64)), "files": g.nfiles}
print(json.dumps(out))
