"""Throughput / latency of the micro-batching queue under C++ load threads.

usage: python tools/queue_bench.py [threads] [max_batch] [max_delay_us] [requests_per_thread] [window]
window = 0: every thread blocks in pgw_queue_evaluate (in-flight requests = threads);
window > 0: every thread keeps that many requests in flight through pgw_queue_submit (an async server).
Prints one JSON line per configuration.  Not a bench.py metric: the queue is host-side plumbing around the measured path.
"""
import ctypes as C
import json
import os
import subprocess
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np  # noqa: E402

import synth  # noqa: E402
from pingoo_b200 import RequestQueue, WafEngine, _ffi  # noqa: E402

so = os.path.join(R, "tools", "libqueue_load.so")
if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(R, "tools", "queue_load.cpp")):
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", so, os.path.join(R, "tools", "queue_load.cpp"),
                           "-L" + os.path.join(R, "pingoo_b200"), "-l:libpingoo_waf.so", "-Wl,-rpath," + os.path.join(R, "pingoo_b200")])
lib = C.CDLL(so)
lib.queue_load_run.argtypes = [C.c_void_p, C.POINTER(_ffi.Batch), C.c_uint32, C.c_uint32, C.c_void_p] + [C.POINTER(C.c_double)] * 4
lib.queue_load_run_async.argtypes = [C.c_void_p, C.POINTER(_ffi.Batch), C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p] + [C.POINTER(C.c_double)] * 4

threads = int(sys.argv[1]) if len(sys.argv) > 1 else 64
max_batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
delay = int(sys.argv[3]) if len(sys.argv) > 3 else 200
per = int(sys.argv[4]) if len(sys.argv) > 4 else 20000
window = int(sys.argv[5]) if len(sys.argv) > 5 else 0

rules, payloads, _ = synth.make_ruleset(128)
batch = synth.RequestStream(config_id=2, payloads=payloads).generate(0, 200_000)
eng = WafEngine(rules, device=0)
want = eng.evaluate_host(batch)
q = RequestQueue(eng, max_batch=max_batch, max_delay_us=delay)
cb = batch.as_ctypes()
out = np.zeros(batch.n, dtype=np.uint32)
sec, p50, p99, mx = C.c_double(), C.c_double(), C.c_double(), C.c_double()
if window:
    fails = lib.queue_load_run_async(q._q, C.byref(cb), threads, per, window, out.ctypes.data, C.byref(sec), C.byref(p50), C.byref(p99), C.byref(mx))
else:
    fails = lib.queue_load_run(q._q, C.byref(cb), threads, per, out.ctypes.data, C.byref(sec), C.byref(p50), C.byref(p99), C.byref(mx))
st = q.stats()
done = min(batch.n, threads * per)
seen = np.zeros(batch.n, dtype=bool)
seen[[(t * per + k) % batch.n for t in range(threads) for k in range(min(per, batch.n))][:done]] = True
print(json.dumps({"mode": "submit+callback" if window else "blocking", "window": window, "threads": threads, "max_batch": max_batch, "max_delay_us": delay, "requests": threads * per, "failures": fails,
                  "requests_per_s": threads * per / sec.value, "latency_us": {"p50": p50.value, "p99": p99.value, "max": mx.value},
                  "batches": st.batches, "avg_batch": st.requests / max(1, st.batches), "full_flushes": st.full_flushes,
                  "deadline_flushes": st.deadline_flushes,
                  "verdict_mismatches_vs_batch_path": int(np.count_nonzero(out[seen] != want[seen]))}))
q.close()
