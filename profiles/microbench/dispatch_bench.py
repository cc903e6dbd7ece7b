"""Kernel times of the query exchange (igd_b200_dispatcher_*) on ONE GPU: a world of one rank, every query local.
Separates what the kernels cost from what NVLink and the waiting for peers cost in a multi-GPU step.
usage: python profiles/microbench/dispatch_bench.py [n_queries]"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from igd_b200 import api, synth, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
wd = os.environ.get("IGD_B200_WORKDIR", "/tmp/igd_b200_bench")
path = workloads.build_db("cfg3s", wd)
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29577", rank=0, world_size=1)
g = api.IGD(path, device=0)
q = workloads.queries("cfg3s", n)
ctg = g.resolve_names(synth.HG38_NAMES)[q.ctg].astype(np.int32)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (ctg, q.start, q.end)]
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
cost = torch.zeros(g.ntiles + 1, dtype=torch.int64, device=dev)
cuts = torch.zeros(2, dtype=torch.int32, device=dev)
disp = api.QueryDispatcher(g, 0, n)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
for it in range(4):
    ev[0].record()
    cost.zero_()
    g.shard_cost(n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), cost.data_ptr(), st.cuda_stream)
    ev[1].record()
    g.shard_cuts(cost.data_ptr(), 1, cuts.data_ptr(), st.cuda_stream)
    ev[2].record()
    disp.dispatch_async(n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), cuts.data_ptr(), st.cuda_stream)
    ev[3].record()
    torch.cuda.synchronize()
print(f"n={n}: cost {ev[0].elapsed_time(ev[1]):.4f} ms, cuts {ev[1].elapsed_time(ev[2]):.4f} ms, dispatch (3 launches) {ev[2].elapsed_time(ev[3]):.4f} ms")
disp.status()
disp.close()
g.close()
