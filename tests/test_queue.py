"""Request packer + micro-batching queue: the C shaping against the Python restatement of the listener's
shaping (CPU), and many threads through the queue against the oracle (GPU)."""
import os
import threading

import numpy as np
import pytest

import scenarios
from helpers import Oracle
from pingoo_b200 import RequestQueue, WafEngine, make_request, pack_requests, shape_request

RAW = [
    dict(host="example.com", url="/", path="/", method="GET", user_agent="Mozilla/5.0"),
    dict(host="  spaced.example \t", url="/a/b/?q=1", path="/a/b///", method="POST", user_agent="  curl/8.0  "),
    dict(host="h" * 256, url="/x", path="/x", method="GET", user_agent="u" * 256),
    dict(host="h" * 257, url="/x", path="/x/", method="GET", user_agent="u" * 257),          # over-long -> ""
    dict(host=" " + "h" * 256 + " ", url="/x", path="//", method="GET", user_agent="\t" + "u" * 256 + " "),  # trimmed first
    dict(host="tab\there", url="/", path="", method="HEAD", user_agent="tab\tinside"),      # tab is allowed by to_str
    dict(host="", url="", path="", method="", user_agent=""),
    dict(host="caf\u00e9.example", url="/", path="/", method="GET", user_agent="Mozilla/5.0 \u2713"),  # non-ASCII header values
]
RAW_BYTES = [
    dict(host=b"caf\xc3\xa9.example", url=b"/caf\xc3\xa9", path=b"/caf\xc3\xa9/", method=b"GET", user_agent=b"Mozilla/5.0 (\xe2\x9c\x93)"),
    dict(host=b"ctl\x01.example", url=b"/", path=b"/", method=b"GET", user_agent=b"bad\x7fagent"),
    dict(host=b"ok.example", url=b"/", path=b"/", method=b"GET", user_agent=b"line\r\nbreak"),
]


def test_shaping_matches_the_python_restatement_of_the_listener():
    for r in RAW:
        want = pack_requests([dict(r, ip="1.2.3.4", remote_port=1)])
        got = shape_request(make_request(**{k: v.encode() for k, v in r.items()}))
        for f in ("host", "url", "path", "method", "user_agent"):
            assert got[f] == want.field(f, 0), (r, f)
    # header values that are not visible ASCII: HeaderValue::to_str fails and the listener falls back to ""
    for r in RAW_BYTES:
        got = shape_request(make_request(**r))
        assert got["host"] == (r["host"] if r["host"] == b"ok.example" else b"")
        assert got["user_agent"] == b""
        assert got["url"] == r["url"] and got["path"] == r["path"].rstrip(b"/")


@pytest.mark.gpu
def test_threads_through_the_queue_match_the_oracle():
    rules, lists, svcs, batch = scenarios.services(6_000, True)
    n = batch.n
    reqs = []
    for i in range(n):
        ip = bytes(batch.ip[i])
        addr = ".".join(str(b) for b in ip[:4]) if not batch.ip_is_v6[i] else __import__("ipaddress").ip_address(ip).compressed
        reqs.append(make_request(host=batch.field("host", i), url=batch.field("url", i), path=batch.field("path", i),
                                 method=batch.field("method", i), user_agent=batch.field("user_agent", i), ip=addr,
                                 remote_port=int(batch.remote_port[i]), flags=int(batch.flags[i]) if batch.flags is not None else 0))
    want_v, want_s = Oracle(rules, lists, services=svcs).evaluate_routed(batch, threads=os.cpu_count() or 1)
    eng = WafEngine(rules, lists, device=0, services=svcs)
    q = RequestQueue(eng, max_batch=512, max_delay_us=500)
    got_v = np.zeros(n, dtype=np.uint32)
    got_s = np.zeros(n, dtype=np.uint16)
    n_threads = 16

    def work(t):
        for i in range(t, n, n_threads):
            got_v[i], got_s[i] = q.evaluate(reqs[i])

    th = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    [t.start() for t in th]
    [t.join() for t in th]
    st = q.stats()
    assert st.requests == n and st.batches >= n // 512
    assert np.array_equal(got_v, want_v)
    assert np.array_equal(got_s, want_s)
    # a single request must come back after the deadline, not wait for a full batch
    v, s = q.evaluate(reqs[0])
    assert (v, s) == (int(want_v[0]), int(want_s[0]))
    assert q.stats().deadline_flushes >= 1
    # completion callbacks (pgw_queue_submit): everything in flight at once, full batches
    import ctypes as C
    from pingoo_b200 import _ffi

    res_v = np.full(n, 0xFFFFFFFF, dtype=np.uint32)
    res_s = np.zeros(n, dtype=np.uint16)
    left = threading.Semaphore(0)

    @_ffi.DONE_FN
    def done(user, verdict, service, rc):
        i = int(user or 0)
        res_v[i] = verdict
        res_s[i] = service
        left.release()

    q.close()
    # a long deadline and a small batch: the (slow) Python submitter fills batches before the deadline fires
    q = RequestQueue(eng, max_batch=64, max_delay_us=200_000)
    for i in range(n):
        assert eng._lib.pgw_queue_submit(q._q, C.byref(reqs[i]), done, C.c_void_p(i)) == 0
    for _ in range(n):
        assert left.acquire(timeout=30)
    assert np.array_equal(res_v, want_v) and np.array_equal(res_s, want_s)
    st = q.stats()
    assert st.requests == n and st.full_flushes >= n // 64 - 1 and st.largest_batch == 64
    q.close()


def test_queue_logic_under_thread_sanitizer(tmp_path):
    """The queue's concurrency (two sides, blocking and callback submissions mixed, deadline and full flushes) is exercised
    on the CPU with the device entry points stubbed (tests/queue_stress/stress.cpp), under ThreadSanitizer when the
    toolchain has it: no lost, duplicated or mixed-up request, no data race."""
    import subprocess

    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    exe = str(tmp_path / "stress")
    src = [os.path.join(here, "queue_stress", "stress.cpp"), os.path.join(root, "pingoo_b200", "csrc", "queue.cpp")]
    base = ["g++", "-O1", "-g", "-std=c++17", "-pthread", "-o", exe] + src
    if subprocess.run(base[:1] + ["-fsanitize=thread"] + base[1:], capture_output=True).returncode != 0:
        subprocess.check_call(base)  # no TSAN runtime: plain build, the checksums still catch mix-ups
    for args in (["64", "300"], ["8", "20000"], ["1", "0"]):
        r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr, (args, r.stdout[-300:], r.stderr[-2000:])
