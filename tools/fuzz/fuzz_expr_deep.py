import sys, collections
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np
from helpers import Oracle, Sim
import test_fuzz_expressions as t
lo, hi = int(sys.argv[1]), int(sys.argv[2])
ref = collections.Counter(); bad = 0; ok = 0
for seed in range(lo, hi):
    rules, svcs, lists, batch = t.make_case(seed)
    want_v, want_s = Oracle(rules, lists, services=svcs).evaluate_routed(batch, threads=2)
    try:
        got_v, got_s = Sim(rules, lists, services=svcs).evaluate_routed(batch)
    except Exception as e:
        msg = str(e); ref[msg[msg.find("':")+2:][:90]] += 1; continue
    ok += 1
    d = np.nonzero((got_v != want_v) | (got_s != want_s))[0]
    if len(d):
        bad += 1
        print("MISMATCH seed", seed, len(d), [r.expression for r in rules], [s.route for s in svcs])
print(lo, hi, "ok", ok, "bad", bad, "refused", sum(ref.values()))
for k, v in ref.most_common(8): print("   ", v, k)
