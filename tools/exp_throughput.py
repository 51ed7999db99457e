import sys, os
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np, torch
import synth
from pingoo_b200 import WafEngine
def run(rules, batch, label, steps=10):
    eng = WafEngine(rules, device=0)
    info = eng.info()
    t, cb = eng.to_device(batch)
    out = torch.empty(batch.n, dtype=torch.int32, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): eng.evaluate_device(cb, out, st)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i in range(steps):
        eng.evaluate_device(cb, out, st)
        evs[i + 1].record()
    torch.cuda.synchronize()
    per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(steps))
    ms = evs[0].elapsed_time(evs[steps])/steps
    label = f"{label} [min {per[0]:.3f} med {per[len(per)//2]:.3f} max {per[-1]:.3f}]"
    scanned = sum(batch.total[f] for i,f in enumerate(synth.FIELDS) if (info.scanned_fields_mask>>i)&1)
    print(f"{label}: {ms:.3f} ms  {batch.n/ms/1e3:.1f} M req/s  alg {scanned/ms/1e6:.0f} GB/s  units={info.n_scan_units} hot={info.hot_dfa_states}/{info.total_dfa_states} arena={info.table_arena_bytes} smem={info.smem_bytes}", flush=True)
rules, payloads, _ = synth.make_ruleset(128)
batch = synth.RequestStream(config_id=2, payloads=payloads).generate(0, 1_000_000)
run(rules, batch, "128 rules")
rules16, p16, _ = synth.make_ruleset(16, config_id=1)
run(rules16, batch, "16 rules")
one=[r for r in rules16 if r.name.startswith('sql_pair')][:1]
run(one, batch, "1 rule url only")
r512, p512, _ = synth.make_ruleset(512)
run(r512, batch, "512 rules")
