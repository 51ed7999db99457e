"""GPU parity: CUDA path (through the C ABI) vs the oracle on the same seeded inputs."""
import os

import numpy as np
import pytest

import scenarios
import synth
from helpers import Oracle, Sim, fmt_verdict
from pingoo_b200 import Action, Rule, WafEngine, pack_requests

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


def _explain(batch, rules, want, got, limit=5):
    bad = np.nonzero(want != got)[0]
    lines = [f"{len(bad)} of {batch.n} verdicts differ"]
    for i in bad[:limit]:
        lines.append(f"  req {i}: oracle {fmt_verdict(want[i])} gpu {fmt_verdict(got[i])} url={batch.field('url', i)[:120]!r} ua={batch.field('user_agent', i)[:60]!r}")
    return "\n".join(lines)


def _check(rules, batch, lists=None, mmdb=None, eval_gates=True, **opts):
    eng = WafEngine(rules, lists, mmdb, device=0, eval_gates=eval_gates, **opts)
    got = eng.evaluate_host(batch)
    want = Oracle(rules, lists, mmdb, eval_gates=eval_gates).evaluate(batch, threads=THREADS)
    assert np.array_equal(got, want), _explain(batch, rules, want, got)
    # device-pointer entry point must agree with the host-pointer one
    import torch

    t, cb = eng.to_device(batch)
    out = torch.empty(batch.n, dtype=torch.int32, device="cuda")
    eng.evaluate_device(cb, out, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want)
    assert eng.info().kernel_launches >= 2
    return eng, want


def test_config1_10k_get_16_rules():
    rules, lists, mmdb, batch, gates_on = scenarios.config1()
    eng, want = _check(rules, batch)
    assert eng.info().tables_in_smem == 1
    assert np.bincount(want & 3, minlength=4)[1] > 0  # some requests are blocked


def test_config2_sample_128_rules():
    rules, lists, mmdb, batch, gates_on = scenarios.config2_sample()
    _check(rules, batch)


def test_256_rules():
    rules, lists, mmdb, batch, gates_on = scenarios.rules256()
    _check(rules, batch)


def test_ragged_and_empty_inputs():
    rules, reqs = scenarios.ragged()
    _check(rules, pack_requests(reqs))
    _check(rules, pack_requests(reqs[:1]))
    _check(rules, pack_requests(reqs[:33]), eval_gates=False)


def test_gates_flags_and_action_lists():
    rules, batch = scenarios.gates()
    _check(rules, batch)
    _check(rules, batch, eval_gates=False)


def test_lists_ints_country_and_geoip():
    rules, lists, mmdb, batch, gates_on, records = scenarios.lists_geo()
    eng, want = _check(rules, batch, lists, mmdb)
    assert eng.info().lpm_present == 1 and eng.info().geoip_loaded == 1
    assert len(set((want >> 2).tolist())) > 10  # many different rules decide


def test_geoip_lookup_batch_matches_oracle():
    import torch

    mmdb, records = synth.make_geoip(800, config_id=9)
    eng = WafEngine([Rule("r", "client.asn == 1", [Action.BLOCK])], geoip_mmdb=mmdb, device=0)
    orc = Oracle([], geoip_mmdb=mmdb)
    ip_np, v6_np = scenarios.geo_probe_addresses(records)
    ip_t, v6_t = torch.from_numpy(ip_np).cuda(), torch.from_numpy(v6_np).cuda()
    asn_t = torch.empty(len(v6_np), dtype=torch.int32, device="cuda")
    cc_t = torch.empty(len(v6_np), dtype=torch.int16, device="cuda")
    eng.geoip_lookup_device(ip_t, v6_t, asn_t, cc_t)
    torch.cuda.synchronize()
    asn, cc = asn_t.cpu().numpy().view(np.uint32), cc_t.cpu().numpy().view(np.uint16)
    for i in range(len(v6_np)):
        a, c = orc.geoip_lookup(bytes(ip_np[i]), int(v6_np[i]))
        assert (int(asn[i]), bytes([cc[i] & 0xFF, cc[i] >> 8]).decode()) == (a, c), f"address {i}"


def test_kernel_agrees_with_compiled_tables():
    """The kernel must execute the compiled tables exactly as the test-only CPU walk does (kernel mechanics, not semantics)."""
    rules, payloads, _ = synth.make_ruleset(128, config_id=2)
    batch = synth.RequestStream(config_id=2, payloads=payloads, attack_rate=0.3).generate(100_000, 40_000)
    eng = WafEngine(rules, device=0)
    assert np.array_equal(eng.evaluate_host(batch), Sim(rules).evaluate(batch))


def test_full_size_properties_config2():
    """BASELINE config 2 at full size (1M x 128): size-independent checks + oracle parity on a strided sample."""
    rules, payloads, _ = synth.make_ruleset(128, config_id=2)
    stream = synth.RequestStream(config_id=2, payloads=payloads)
    batch = stream.generate(0, 1_000_000)
    eng = WafEngine(rules, device=0)
    v = eng.evaluate_host(batch)
    # idempotence: a second evaluation returns identical verdicts
    assert np.array_equal(v, eng.evaluate_host(batch))
    # shard invariance: evaluating halves separately and concatenating gives the same verdicts
    h = batch.n // 2
    assert np.array_equal(np.concatenate([eng.evaluate_host(batch.slice(0, h)), eng.evaluate_host(batch.slice(h, batch.n))]), v)
    # decided rule indices are in range and actions are consistent with the rule's action list
    act, rule = v & 3, v >> 2
    assert np.all((rule < len(rules)) | (rule == 0x3FFFFFFF))
    # a request no rule decided is allowed, unless a gate decided it (bypass of the captcha API; block of a bad user agent)
    assert np.all(act[rule == 0x3FFFFFFF] != 2)
    by_rule = {int(r): set(int(a) for a in rules[int(r)].actions) for r in np.unique(rule[rule != 0x3FFFFFFF])}
    for r, acts in by_rule.items():
        assert set(np.unique(act[rule == r]).tolist()) <= acts, r
    # oracle parity on every 23rd request
    idx = np.arange(0, batch.n, 23)
    want = Oracle(rules).evaluate(batch, threads=THREADS)[idx] if batch.n <= 200_000 else None
    if want is None:
        sub_reqs = [batch.slice(int(i), int(i) + 1) for i in idx[:2000]]
        orc = Oracle(rules)
        for i, sb in zip(idx[:2000], sub_reqs):
            assert orc.evaluate(sb)[0] == v[i], f"request {i}"


def test_service_routes_in_the_same_pass():
    """SURVEY.md 8f #1 (http_listener.rs:266-272, http_proxy_service.rs:84-95): verdict and first matching service from
    one scan, through both entry points."""
    import torch
    from pingoo_b200._ffi import NO_SERVICE

    for catch_all in (True, False):
        rules, lists, svcs, batch = scenarios.services(60_000, catch_all)
        want_v, want_s = Oracle(rules, lists, services=svcs).evaluate_routed(batch, threads=THREADS)
        eng = WafEngine(rules, lists, device=0, services=svcs)
        got_v, got_s = eng.evaluate_host_routed(batch)
        assert np.array_equal(got_v, want_v), _explain(batch, rules, want_v, got_v)
        bad = np.nonzero(got_s != want_s)[0]
        assert len(bad) == 0, f"{len(bad)} services differ, first: req {bad[0]} oracle {want_s[bad[0]]} gpu {got_s[bad[0]]}"
        t, cb = eng.to_device(batch)
        out = torch.empty(batch.n, dtype=torch.int32, device="cuda")
        svc = torch.empty(batch.n, dtype=torch.int16, device="cuda")
        eng.evaluate_device_routed(cb, out, svc, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want_v)
        assert np.array_equal(svc.cpu().numpy().view(np.uint16), want_s)
        # the plain entry point is unaffected by the presence of services
        assert np.array_equal(eng.evaluate_host(batch), want_v)
        assert np.all(want_s[(want_v & 3) != 0] == NO_SERVICE)
    # services only (no WAF rule): everything is allowed and routed
    rules, lists, svcs, batch = scenarios.services(5_000, True)
    want_v, want_s = Oracle([], lists, services=svcs).evaluate_routed(batch, threads=THREADS)
    got_v, got_s = WafEngine([], lists, device=0, services=svcs).evaluate_host_routed(batch)
    assert np.array_equal(got_v, want_v) and np.array_equal(got_s, want_s)


def test_tiny_batches_and_concurrent_streams():
    """Edge sizes (0, 1, 31, 32, 33 requests) and several host threads evaluating on their own streams against one shared
    ruleset (many connection tasks evaluate concurrently against the same Arc'd rules)."""
    import threading

    import torch

    rules, lists, mmdb, batch, g = scenarios.config2_sample(4_000)
    eng = WafEngine(rules, device=0)
    want = Oracle(rules).evaluate(batch, threads=THREADS)
    for n in (0, 1, 31, 32, 33, 1000):
        sub = batch.slice(0, n)
        assert np.array_equal(eng.evaluate_host(sub), want[:n]), n
    parts = [batch.slice(i * 1000, (i + 1) * 1000) for i in range(4)]
    outs = [None] * 4
    errs = []

    def work(i):
        try:
            torch.cuda.set_device(0)
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                t, cb = eng.to_device(parts[i])
                o = torch.empty(parts[i].n, dtype=torch.int32, device="cuda")
                for _ in range(20):
                    eng.evaluate_device(cb, o, st.cuda_stream)
                st.synchronize()
                outs[i] = o.cpu().numpy().view(np.uint32)
        except Exception as e:  # surfaced below
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for i in range(4):
        assert np.array_equal(outs[i], want[i * 1000:(i + 1) * 1000]), i


def _big_config(cfg_id, n):
    import bench

    desc, rules, lists, mmdb, batches = bench.build_workload(cfg_id, 0, n)
    return rules, lists, mmdb, batches[0]


def test_config3_512_rules_100k_blocklist_geoip():
    """BASELINE config 3: 512 rules + 100 000-entry IP/CIDR blocklist + GeoIP ASN / country predicates
    (pingoo/lists.rs:62-113, pingoo/geoip.rs:73-91), 120 000 requests of its stream."""
    rules, lists, mmdb, batch = _big_config(3, 120_000)
    eng, want = _check(rules, batch, lists, mmdb)
    info = eng.info()
    assert info.lpm_present == 1 and info.geoip_loaded == 1 and info.n_rules == 512
    hit_rules = set((want >> 2).tolist())
    assert 1 in hit_rules and len(hit_rules) > 60  # rule 1 = the blocklist rule


def test_config4_1024_rules():
    """BASELINE config 4 rule set (1 024 rules: several DFA units per field, unit masks on the gate candidates)."""
    rules, lists, mmdb, batch = _big_config(4, 120_000)
    eng, want = _check(rules, batch)
    assert eng.info().n_rules == 1024 and eng.info().n_scan_units >= 8
    assert len(set((want >> 2).tolist())) > 100
    # the gate is a prefilter only: same verdicts without it
    off = WafEngine(rules, device=0, candidate_gate=False)
    assert off.info().gated_fields_mask == 0
    assert np.array_equal(off.evaluate_host(batch), want)


def test_config5_long_uris_pathological_rules():
    """BASELINE config 5: 8 KB URIs x 256 backtracking-prone patterns (none of them can be gated: dozens of full-field
    units, fields far longer than a claim pool's worth of bytes)."""
    rules, lists, mmdb, batch = _big_config(5, 3_000)
    eng, want = _check(rules, batch)
    assert eng.info().n_scan_units >= 20
    assert np.count_nonzero(want & 3) > 100


def test_more_than_64_scan_units():
    """A small per-unit state cap forces more units than one scan launch carries in its parameter bank (kMaxConstUnits)."""
    rules, payloads, _ = synth.make_ruleset(512, config_id=3)
    batch = synth.RequestStream(config_id=3, payloads=payloads, attack_rate=0.2).generate(0, 30_000)
    eng, want = _check(rules, batch, max_dfa_states=96)
    assert eng.info().n_scan_units > 64, eng.describe()
    eng2, _ = _check(rules, batch.slice(0, 10_000), max_dfa_states=96, candidate_gate=False)
    assert eng2.info().n_scan_units > 64


def test_host_entry_point_is_thread_safe():
    """SURVEY.md 8(b) Threading: many workers evaluate concurrently against one shared ruleset
    (http_listener.rs:91-103,134-138); pgw_evaluate_batch_host takes per-call staging from a pool."""
    import threading

    rules, lists, mmdb, batch, g = scenarios.config2_sample(24_000)
    eng = WafEngine(rules, device=0)
    want = Oracle(rules).evaluate(batch, threads=THREADS)
    parts = [(i * 3_000, (i + 1) * 3_000) for i in range(8)]
    errs = []

    def work(lo, hi):
        try:
            sub = batch.slice(lo, hi)
            for _ in range(10):
                got = eng.evaluate_host(sub)
                if not np.array_equal(got, want[lo:hi]):
                    errs.append((lo, int(np.count_nonzero(got != want[lo:hi]))))
                    return
        except Exception as e:  # surfaced below
            errs.append(e)

    th = [threading.Thread(target=work, args=p) for p in parts]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs


def test_lowered_field_comparisons_arithmetic_and_ordering():
    """Comparisons between two request fields, integer arithmetic on request variables (overflow / division errors make
    the rule "no match"), lexicographic ordering on a field: valid in the reference's language (rules/rules.rs:45-53 accepts
    whatever bel compiles), evaluated by the per-request kernel / as start-anchored patterns."""
    from test_compiler_vs_oracle import _lowered_constructs_case

    rule_sets, batch = _lowered_constructs_case()
    for rules in rule_sets[::3] + [rule_sets[-1]]:
        _check(rules, batch)


def test_adversarial_input_for_the_gate():
    """Every 16-byte chunk of the url column is a level-1 hit (the gate's hit bitmap is all ones): nothing overflows, the
    verdicts stay exact."""
    rules, payloads, _ = synth.make_ruleset(128)
    base = synth.RequestStream(config_id=2, payloads=payloads).generate(0, 3_000)
    reqs = []
    grams = ["select", "union ", "<script", "../../", "/etc/passwd", "curl/"]
    for i in range(base.n):
        g = grams[i % len(grams)]
        url = "/" + (g * 40)[: 150 + (i % 90)]
        reqs.append(dict(host=base.field("host", i).decode(), url=url, path="/" + g.strip("/ <")[:8], method="GET",
                         user_agent=(g * 8)[:60] or "x", ip="10.1.%d.%d" % (i // 250, i % 250), remote_port=1000 + i))
    _check(rules, pack_requests(reqs))


@pytest.mark.gpu
def test_constant_receivers_dynamic_lists_and_field_ordering():
    """`"GET POST".contains(method)`, list literals holding request variables, integer expressions looked up in lists and the
    byte-wise ordering of two fields (FIELD_CMP ops 4-7 in the per-request kernel) -- tests/scenarios.py value_constructs."""
    rules, lists, batch = scenarios.value_constructs()
    _check(rules, batch, lists, eval_gates=False)
    for r in rules[8:11]:   # the ordering predicates one at a time: nothing in front of them in the first-match loop
        _check([r], batch, lists, eval_gates=False)


@pytest.mark.gpu
def test_gate_resolve_with_and_without_literal_confirmation():
    """tests/test_gate.py's rule set on its alignment / boundary inputs: the user-agent and path atoms are strings with grams of
    their own, confirmed by `waf_gate_resolve_lit_kernel` (wide slots); with literal confirmation switched off the same rules run
    through `waf_gate_resolve_kernel` (narrow slots) and the DFA units.  Both against the oracle."""
    import test_gate

    reqs = []
    for frag in ["union", "../", ".php", ";nc", "'or 1=1", "<svg>", "select from", "3.env", "curl/", "bot", ".git", ".GIT/", ".env"]:
        for pre in range(0, 9):
            for post in range(0, 5):
                url = "q" * pre + frag + "r" * post
                reqs.append(dict(host="h", url=url, path="/" + "p" * (pre % 3) + (frag if post % 2 else ""), method="GET",
                                 user_agent=("Mozilla/5.0 z" if pre % 2 else "x" * pre) + (frag if post < 2 else "") + "y" * post, ip="1.2.3.4", remote_port=1, flags=0))
    batch = pack_requests(reqs * 8)
    eng, want = _check(test_gate.RULES, batch)
    d = eng.describe()
    assert "literals=" in d and "wide-slots" in d, d
    assert len(set(want.tolist())) > 8
    eng2, _ = _check(test_gate.RULES, batch, literal_confirm=False)
    d2 = eng2.describe()
    assert "literals=" not in d2 and "wide-slots" not in d2 and "user_agent/gated" in d2, d2


@pytest.mark.gpu
def test_unreferenced_predicates_are_not_evaluated():
    """tests/scenarios.py unreferenced_predicates: the service of a request must not depend on a predicate no rule mentions."""
    rules, services, lists, batch = scenarios.unreferenced_predicates()
    want_v, want_s = Oracle(rules, lists, services=services).evaluate_routed(batch, threads=THREADS)
    eng = WafEngine(rules, lists, device=0, services=services)
    got_v, got_s = eng.evaluate_host_routed(batch)
    assert np.array_equal(got_v, want_v) and np.array_equal(got_s, want_s)
