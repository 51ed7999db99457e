"""Exports a slice of a BASELINE workload in the fixture format bench/rust_ref replays (same schema as
scenario_verdicts.json): rules of synth.make_ruleset, list CSV text, explicit requests of synth.RequestStream with the
GeoIP columns resolved (the reference resolves asn / country before it builds the rule context,
http_listener.rs:143-157), and the ORACLE's verdicts as the expectation to check or overwrite.

usage (repo root):  python tests/golden/export_stream_fixture.py <config 1..5> <requests> <out.json>
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from helpers import Oracle  # noqa: E402


def main():
    cfg, n, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    if cfg == 1:
        import synth

        rules, payloads, _ = synth.make_ruleset(16, config_id=1)
        batch = synth.RequestStream(config_id=1, payloads=payloads, get_only=True).generate(0, n)
        lists, mmdb = {}, None
    else:
        _, rules, lists, mmdb, batches = bench.build_workload(cfg, 0, n)
        batch = batches[0]
    orc = Oracle(rules, lists, mmdb)
    v = orc.evaluate(batch, threads=os.cpu_count() or 1)
    reqs = []
    for i in range(batch.n):
        ip = bytes(batch.ip[i])
        asn, cc = (orc.geoip_lookup(ip, int(batch.ip_is_v6[i])) if mmdb is not None else (0, "XX"))
        reqs.append({
            "host": batch.field("host", i).decode("latin-1"), "url": batch.field("url", i).decode("latin-1"),
            "path": batch.field("path", i).decode("latin-1"), "method": batch.field("method", i).decode("latin-1"),
            "user_agent": batch.field("user_agent", i).decode("latin-1"), "ip_hex": ip.hex(), "ip_is_v6": int(batch.ip_is_v6[i]),
            "remote_port": int(batch.remote_port[i]), "asn": int(asn), "country": ord(cc[0]) | (ord(cc[1]) << 8),
            "flags": int(batch.flags[i]) if batch.flags is not None else 0,
        })
    case = {
        "name": f"config_{cfg}_first_{batch.n}", "eval_gates": True,
        "rules": [{"name": r.name, "expression": r.expression, "actions": [int(a) for a in r.actions]} for r in rules],
        "services": [], "lists": {k: [int(t), c.decode("latin-1")] for k, (t, c) in (lists or {}).items()},
        "requests": reqs, "verdicts": [int(x) for x in v], "services_out": [0xFFFF] * batch.n,
    }
    with open(out, "w") as f:
        json.dump([case], f)
    hist = np.bincount(v & 3, minlength=4).tolist()
    print(f"wrote {out}: {len(rules)} rules, {batch.n} requests, oracle verdict histogram allow/block/captcha/bypass = {hist}")


if __name__ == "__main__":
    main()
