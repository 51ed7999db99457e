import sys, os
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np, torch
import synth
from pingoo_b200 import WafEngine
which = sys.argv[1] if len(sys.argv) > 1 else "one"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
rules, payloads, _ = synth.make_ruleset(128)
batch = synth.RequestStream(config_id=2, payloads=payloads).generate(0, n)
if which == "one":
    rules16, p16, _ = synth.make_ruleset(16, config_id=1)
    rules = [r for r in rules16 if r.name.startswith('sql_pair')][:1]
elif which == "16":
    rules, _, _ = synth.make_ruleset(16, config_id=1)
eng = WafEngine(rules, device=0)
t, cb = eng.to_device(batch)
out = torch.empty(batch.n, dtype=torch.int32, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for _ in range(5): eng.evaluate_device(cb, out, st)
torch.cuda.synchronize()
print(eng.describe())
