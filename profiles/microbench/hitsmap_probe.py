#!/usr/bin/env python
"""-m on the cfg-2 database (1 905 files, 99 M records): wall time of igd_b200_hitsmap, three calls."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from igd_b200 import api, workloads

p = workloads.build_db("cfg2", os.environ.get("IGD_B200_WORKDIR", "/tmp/igd_b200_bench"))
g = api.IGD(p)
out = []
for _ in range(3):
    t = time.perf_counter()
    hm = g.hitsmap(0)
    out.append(round((time.perf_counter() - t) * 1e3, 2))
print(json.dumps({"hitsmap_cfg2_ms": out, "pairs": int(hm.sum(dtype=np.uint64)), "symmetric": bool((hm == hm.T).all())}))
