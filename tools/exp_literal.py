"""Experiment: config 3 with and without literal confirmation in the gate's resolve kernel (same verdicts; per-group times).
usage: python tools/exp_literal.py [requests] [config]"""
import json
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import torch

import bench
from pingoo_b200 import WafEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 3
desc, rules, lists, mmdb, batches = bench.build_workload(cfg, 0, n)
batch = batches[0]
ref = None
for label, opts in (("literal confirmation on", {}), ("literal confirmation off", {"literal_confirm": False})):
    eng = WafEngine(rules, lists, mmdb, device=0, **opts)
    t, cb = eng.to_device(batch)
    out = torch.empty(batch.n, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        eng.evaluate_device(cb, out, st)
    torch.cuda.synchronize()
    eng.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.evaluate_device(cb, out, st)
    e1.record()
    torch.cuda.synchronize()
    kms, kb = eng.profile_kernels()
    eng.set_profiling(False)
    v = out.cpu().numpy().view(np.uint32)
    if ref is None:
        ref = v.copy()
    print(json.dumps({"variant": label, "config": cfg, "requests": batch.n, "ms_per_batch": e0.elapsed_time(e1) / 10, "groups_ms": kms,
                      "mismatches_vs_first_variant": int(np.count_nonzero(v != ref)), "describe": eng.describe()[-400:]}), flush=True)
    del eng
