"""Run a few batches of one BASELINE config on cuda:0 (profiling driver: meant to be run under ncu).
usage: prof_config.py <config 2|3|4|5> [requests] [iterations] [gate 0|1]"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from pingoo_b200 import WafEngine  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
gate = (int(sys.argv[4]) if len(sys.argv) > 4 else 1) != 0
desc, rules, lists, mmdb, batches = bench.build_workload(cfg, 0, n)
eng = WafEngine(rules, lists, mmdb, device=0, candidate_gate=gate)
batch = batches[0]
t, cb = eng.to_device(batch)
out = torch.empty(batch.n, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
ev[0].record()
for i in range(iters):
    eng.evaluate_device(cb, out, st)
    ev[i + 1].record()
torch.cuda.synchronize()
print(eng.describe())
print("ms per batch:", [round(ev[i].elapsed_time(ev[i + 1]), 4) for i in range(iters)], "requests", batch.n)
