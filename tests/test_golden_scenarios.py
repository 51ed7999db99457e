"""Committed expectations (tests/golden/scenario_verdicts.json, written by tests/golden/make_scenario_goldens.py):
explicit rule sets and requests with the verdict and service the oracle gave when the fixture was made.  The oracle must
still give them (regression pin of the checker), and so must the compiled tables (CPU) and the CUDA path (GPU)."""
import json
import os

import numpy as np
import pytest

from helpers import Oracle, Sim
from pingoo_b200 import Action, ListType, RequestBatch, Rule, Service, WafEngine
from pingoo_b200.batch import FIELDS

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "scenario_verdicts.json")) as f:
    CASES = json.load(f)


def _load(case):
    rules = [Rule(r["name"], r["expression"], [Action(a) for a in r["actions"]]) for r in case["rules"]]
    svcs = [Service(s["name"], s["route"]) for s in case["services"]]
    lists = {k: (ListType(t), c.encode("latin-1")) for k, (t, c) in case["lists"].items()} or None
    reqs = case["requests"]
    n = len(reqs)
    cols = {}
    for f in FIELDS:
        parts = [r[f].encode("latin-1") for r in reqs]
        offs = np.zeros(n + 1, dtype=np.uint32)
        offs[1:] = np.cumsum([len(p) for p in parts], dtype=np.uint64)
        cols[f] = (np.frombuffer(b"".join(parts), dtype=np.uint8), offs)
    batch = RequestBatch(
        n, cols, np.frombuffer(b"".join(bytes.fromhex(r["ip_hex"]) for r in reqs), dtype=np.uint8).reshape(n, 16),
        np.array([r["ip_is_v6"] for r in reqs], dtype=np.uint8), np.array([r["remote_port"] for r in reqs], dtype=np.int32),
        np.array([r["asn"] for r in reqs], dtype=np.int64), np.array([r["country"] for r in reqs], dtype=np.uint16),
        np.array([r["flags"] for r in reqs], dtype=np.uint8))
    return rules, svcs, lists, batch, np.array(case["verdicts"], dtype=np.uint32), np.array(case["services_out"], dtype=np.uint16)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_and_compiled_tables_reproduce_the_fixture(case):
    rules, svcs, lists, batch, want_v, want_s = _load(case)
    v, s = Oracle(rules, lists, services=svcs, eval_gates=case["eval_gates"]).evaluate_routed(batch)
    assert np.array_equal(v, want_v) and np.array_equal(s, want_s), "the oracle no longer reproduces its committed answers"
    # every committed scenario must compile: a construct the engine refused would be a failure here, not a skip
    v, s = Sim(rules, lists, services=svcs, eval_gates=case["eval_gates"]).evaluate_routed(batch)
    assert np.array_equal(v, want_v) and np.array_equal(s, want_s)


@pytest.mark.gpu
def test_cuda_path_reproduces_the_fixture():
    ran = 0
    for case in CASES:
        rules, svcs, lists, batch, want_v, want_s = _load(case)
        eng = WafEngine(rules, lists, device=0, services=svcs, eval_gates=case["eval_gates"])
        v, s = eng.evaluate_host_routed(batch)
        assert np.array_equal(v, want_v) and np.array_equal(s, want_s), case["name"]
        ran += 1
    assert ran == len(CASES)
