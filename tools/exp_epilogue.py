"""Tuning experiment (tools only): per-kernel-group times of one BASELINE config, for the default library or a variant
(PGW_LIB=<path to .so>).  usage: exp_epilogue.py [config] [requests] [steps]"""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from pingoo_b200 import WafEngine  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
desc, rules, lists, mmdb, batches = bench.build_workload(cfg, 0, n)
eng = WafEngine(rules, lists, mmdb, device=0)
t, cb = eng.to_device(batches[0])
out = torch.empty(batches[0].n, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    eng.evaluate_device(cb, out, st)
torch.cuda.synchronize()
eng.set_profiling(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    eng.evaluate_device(cb, out, st)
e1.record()
torch.cuda.synchronize()
ms, k = eng.profile_kernels()
print(os.environ.get("PGW_LIB", "default"), f"cfg {cfg} n {n}: total {e0.elapsed_time(e1) / steps:.3f} ms  gate {ms[0] / k:.3f}  scan {ms[1] / k:.3f}  epilogue+multi {ms[2] / k:.3f}")
