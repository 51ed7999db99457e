"""Regenerates tests/golden/scenario_verdicts.json: rule sets, explicit requests and the ORACLE's verdicts / services.

The fixture pins the oracle against itself over time (a change of oracle semantics shows up as a diff of this file) and
gives the GPU tests a committed expectation that does not depend on a live oracle run.  It is generator-independent: the
requests are explicit strings (the `ragged` and `gates` scenarios and the first random rule sets of
test_fuzz_expressions.py).  Run from the repo root:  python tests/golden/make_scenario_goldens.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import scenarios  # noqa: E402
from helpers import Oracle  # noqa: E402
from pingoo_b200 import pack_requests  # noqa: E402
from test_fuzz_expressions import make_case  # noqa: E402


def dump_case(name, rules, services, lists, batch, eval_gates=True):
    v, s = Oracle(rules, lists, services=services, eval_gates=eval_gates).evaluate_routed(batch, threads=2)
    reqs = []
    for i in range(batch.n):
        ip = bytes(batch.ip[i])
        reqs.append({
            "host": batch.field("host", i).decode("latin-1"), "url": batch.field("url", i).decode("latin-1"),
            "path": batch.field("path", i).decode("latin-1"), "method": batch.field("method", i).decode("latin-1"),
            "user_agent": batch.field("user_agent", i).decode("latin-1"), "ip_hex": ip.hex(), "ip_is_v6": int(batch.ip_is_v6[i]),
            "remote_port": int(batch.remote_port[i]), "asn": int(batch.asn[i]) if batch.asn is not None else 0,
            "country": int(batch.country[i]) if batch.country is not None else 0x5858, "flags": int(batch.flags[i]) if batch.flags is not None else 0,
        })
    return {
        "name": name, "eval_gates": eval_gates,
        "rules": [{"name": r.name, "expression": r.expression, "actions": [int(a) for a in r.actions]} for r in rules],
        "services": [{"name": s_.name, "route": s_.route} for s_ in services],
        "lists": {k: [int(t), c.decode("latin-1")] for k, (t, c) in (lists or {}).items()},
        "requests": reqs, "verdicts": [int(x) for x in v], "services_out": [int(x) for x in s],
    }


def main():
    cases = []
    rules, reqs = scenarios.ragged()
    cases.append(dump_case("ragged", rules, [], None, pack_requests(reqs)))
    rules, batch = scenarios.gates()
    cases.append(dump_case("gates", rules, [], None, batch))
    cases.append(dump_case("gates_off", rules, [], None, batch, eval_gates=False))
    for seed in range(6):
        rules, svcs, lists, batch = make_case(seed, n_requests=60)
        cases.append(dump_case(f"fuzz_{seed}", rules, svcs, lists, batch))
    with open(os.path.join(HERE, "scenario_verdicts.json"), "w") as f:
        json.dump(cases, f, indent=0, sort_keys=True)
    print(len(cases), "cases,", sum(len(c["requests"]) for c in cases), "requests")


if __name__ == "__main__":
    main()
