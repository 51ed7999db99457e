"""The bit-parallel NFA unit (csrc/nfa_bits.hpp, kernel_bitset.cuh): patterns whose DFA would exceed the scan-unit caps.

Rust `regex` never refuses a pattern for the size of its DFA (it falls back to the PikeVM); here such a pattern is
simulated as a bit vector of NFA positions.  CPU tests walk the compiled tables (tests/sim) against the oracle; the
`gpu` tests run the kernel through the C ABI against the oracle.  Two ways to reach the unit: patterns that really
explode (`a[ab]{15}c` needs 2^16 DFA states), and `max_dfa_states=1`, which pushes EVERY pattern of a scenario through it.
"""
import random

import numpy as np
import pytest

import scenarios
from helpers import Oracle, Sim, fmt_verdict
from pingoo_b200 import Action, Rule, pack_requests


def _explain(batch, want, got):
    bad = np.nonzero(got != want)[0]
    return "\n".join([f"{len(bad)} of {batch.n} differ"] +
                     [f"req {i}: oracle {fmt_verdict(want[i])} engine {fmt_verdict(got[i])} url={batch.field('url', i)[:120]!r} ua={batch.field('user_agent', i)[:60]!r}"
                      for i in bad[:5]])


def _sim_check(rules, batch, lists=None, mmdb=None, eval_gates=True, **opts):
    sim = Sim(rules, lists, mmdb, eval_gates=eval_gates, **opts)
    got = sim.evaluate(batch)
    want = Oracle(rules, lists, mmdb, eval_gates=eval_gates).evaluate(batch, threads=8)
    assert np.array_equal(got, want), _explain(batch, want, got)
    return sim, want


# ---- patterns that really need more DFA states than any unit may have ----------------------------------------------------

def exploding_rules():
    return [
        Rule("k15", 'http_request.url.matches("a[ab]{15}c")', [Action.BLOCK]),                       # 17 positions, 2^16 subsets
        Rule("k15i", 'http_request.url.matches("(?i)x[XY]{15}Z")', [Action.CAPTCHA]),                 # case-insensitive
        Rule("word", 'http_request.url.matches("a[ab]{14}\\\\b[ -]")', [Action.BLOCK]),                 # \b inside
        Rule("notword", 'http_request.url.matches("b[ab]{14}\\\\B[ab_]")', [Action.CAPTCHA]),           # \B inside
        Rule("line", 'http_request.url.matches("(?m)b[ab]{15}$")', [Action.CAPTCHA, Action.BLOCK]),    # end of a line
        Rule("end", 'http_request.url.matches("a[ab]{16}$")', [Action.BLOCK]),                        # end of the field
        Rule("gap", 'http_request.url.matches("a[ab]{15}<[^>]*>")', [Action.BLOCK]),                  # X G* S shape: latch events
        Rule("ua", 'http_request.user_agent.matches("b[ab]{17}a")', [Action.CAPTCHA]),                 # another field
        Rule("wide", 'http_request.url.matches("a[ab]{300}z")', [Action.BLOCK]),                      # 302 positions: 10 state words
        Rule("small", 'http_request.url.contains("../")', [Action.CAPTCHA]),                          # an ordinary neighbour
    ]


def exploding_requests(n, seed=5, long_every=0):
    rng = random.Random(seed)
    reqs = []
    for i in range(n):
        kind = rng.randrange(10)
        if kind < 5:
            url = "".join(rng.choice("ababababc \nqQxXyYzZ<>_/.") for _ in range(rng.randint(0, 60)))
        elif kind < 7:   # near misses and hits of the counted patterns
            body = "".join(rng.choice("ab") for _ in range(rng.choice([14, 15, 15, 16, 17])))
            url = rng.choice(["", "x", " ", "b\n", "zz "]) + rng.choice(["a", "b", "c", "x", "y"]) + body + rng.choice(["c", "", " ", "-", "_", "\n", "\nab", "Z", "z", "<>", "<a b>", "<ab"])
            if rng.random() < 0.3:
                url = url.replace("a", "X").replace("b", "y") if rng.random() < 0.5 else url.upper()
        elif kind < 9:
            body = "".join(rng.choice("ab") for _ in range(rng.choice([299, 300, 300, 301])))
            url = rng.choice(["", "qq", "z"]) + "a" + body + rng.choice(["z", "", "q", "az"])
        else:
            url = "/x/../y" if rng.random() < 0.5 else ""
        if long_every and i % long_every == 0:
            url = url + "c" * 40 + "a" + "".join(rng.choice("ab") for _ in range(15)) + rng.choice(["c", "d"])
        ua = "Mozilla/5.0 " + "".join(rng.choice("ab") for _ in range(rng.randint(0, 30)))
        reqs.append(dict(host="h.example", url=url, path="/p", method="GET", user_agent=ua, ip="10.1.2.3", remote_port=4000 + i % 100, flags=i % 2))
    return pack_requests(reqs)


def test_exploding_patterns_run_as_bitset_units():
    rules = exploding_rules()
    batch = exploding_requests(6000)
    sim, want = _sim_check(rules, batch)
    desc = sim.describe()
    assert desc.count("bitset-nfa") >= 9, desc
    assert "url/bitset-nfa: positions=302" in desc, desc
    # every rule decides some request, and most requests are allowed
    assert len(set((want >> 2).tolist())) >= len(rules), sorted(set((want >> 2).tolist()))


def test_every_pattern_forced_through_the_bitset_unit():
    """max_dfa_states=1 leaves no DFA at all: gap-split latches, complements, early-exit prefixes, `$`, `\\b` and the
    internal captcha-path gate all run as NFA position sets and must give the oracle's verdicts."""
    rules, reqs = scenarios.ragged()
    sim, _ = _sim_check(rules, pack_requests(reqs), max_dfa_states=1, candidate_gate=False)
    assert "bitset-nfa" in sim.describe() and "states=" not in sim.describe().split("gate(")[0].replace("max_dfa_states", "")
    rules, batch = scenarios.gates()
    _sim_check(rules, batch, max_dfa_states=1, candidate_gate=False)
    rules, lists, mmdb, batch, _ = scenarios.config1()
    _sim_check(rules, batch.slice(0, 4000), max_dfa_states=1, candidate_gate=False)
    rules, lists, mmdb, batch, _ = scenarios.config2_sample(6000, attack_rate=0.2)
    _sim_check(rules, batch, max_dfa_states=1, candidate_gate=False)
    # with the gate in front: gated classes lose their bundles to the bitset units, literals stay with the resolve step
    _sim_check(rules, batch.slice(0, 3000), max_dfa_states=1)


def test_gap_split_and_complement_patterns_through_the_bitset_unit():
    # the same rule sets as the DFA-path tests, no DFA allowed
    rng = random.Random(99)
    rules = [Rule(f"tag{i}", 'http_request.url.matches("(?i)<%s[^>]*>")' % tg, [Action.BLOCK]) for i, tg in enumerate(["script", "svg", "a", "ab"])]
    rules += [Rule("quote", 'http_request.url.matches("x=\\"[^\\"]*\\"")', [Action.BLOCK]),
              Rule("dotstar", 'http_request.url.matches("select.*;")', [Action.CAPTCHA]),
              Rule("plus", 'http_request.url.matches("a[^b]+b")', [Action.BLOCK]),
              Rule("not_moz", '!http_request.user_agent.starts_with("Mozilla/") && !http_request.user_agent.contains("curl/")', [Action.CAPTCHA])]
    alphabet = ['<', '>', 'script', 'SVG', 'a', 'b', 'ab', ' ', '/', 'x=', '"', 'select', ';', '\n', '<a', '<ab>', 'aab']
    reqs = []
    for i in range(3000):
        url = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, 14)))
        ua = rng.choice(["Mozilla/5.0", "Mozilla", "curl/8", "x curl/ y", "M", "Nozilla/5"])
        reqs.append(dict(host="h", url=url, path="/p", method="GET", user_agent=ua, ip="1.2.3.4", remote_port=1, flags=i % 2))
    batch = pack_requests(reqs)
    _, want = _sim_check(rules, batch, max_dfa_states=1, candidate_gate=False)
    assert len(set(want.tolist())) > 6


def test_special_word_boundaries_on_both_kinds_of_unit():
    """\\b{start}, \\b{end}, their halves and \\< \\> (regex 1.10): same verdicts from the DFA units and from the bitset unit."""
    rules = [Rule("ws", 'http_request.url.matches("\\\\b{start}cat")', [Action.BLOCK]),
             Rule("we", 'http_request.url.matches("dog\\\\b{end}")', [Action.CAPTCHA]),
             Rule("angle", 'http_request.url.matches("\\\\<a[ab]{2}\\\\>")', [Action.BLOCK]),
             Rule("half", 'http_request.url.matches("\\\\b{start-half}=[ab]+\\\\b{end-half}")', [Action.CAPTCHA, Action.BLOCK]),
             Rule("both", 'http_request.user_agent.matches("(?i)\\\\b{start}bot\\\\b{end}")', [Action.BLOCK])]
    rng = random.Random(3)
    words = ["cat", "dog", "concat", "dogs", "aab", "abb", "xaab", "=ab", "=a", "a=b", " ", "-", "_", ".", "bot", "Bot", "robot", "bots", "\n"]
    reqs = []
    for i in range(4000):
        url = "".join(rng.choice(words) for _ in range(rng.randint(0, 6)))
        ua = "Mozilla/5.0 " + "".join(rng.choice(words[10:18]) for _ in range(rng.randint(0, 4)))
        reqs.append(dict(host="h", url=url, path="/p", method="GET", user_agent=ua, ip="1.2.3.4", remote_port=1, flags=i % 2))
    batch = pack_requests(reqs)
    _, want = _sim_check(rules, batch)
    assert len(set((want >> 2).tolist())) >= 6
    sim, _ = _sim_check(rules, batch, max_dfa_states=1, candidate_gate=False)
    assert sim.describe().count("bitset-nfa") >= 5


def test_position_limit_is_reported():
    rules = [Rule("huge", 'http_request.url.matches("a[ab]{2100}c")', [Action.BLOCK])]
    with pytest.raises(ValueError) as ei:
        Sim(rules, max_dfa_states=64)
    assert "NFA positions" in str(ei.value) and "http_request.url" in str(ei.value)


# ---- the kernel ----------------------------------------------------------------------------------------------------------

def _gpu_check(rules, batch, **opts):
    import torch

    from pingoo_b200 import WafEngine

    eng = WafEngine(rules, device=0, **opts)
    got = eng.evaluate_host(batch)
    want = Oracle(rules).evaluate(batch, threads=16)
    assert np.array_equal(got, want), _explain(batch, want, got)
    t, cb = eng.to_device(batch)
    out = torch.empty(batch.n, dtype=torch.int32, device="cuda")
    eng.evaluate_device(cb, out, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want)
    return eng, want


@pytest.mark.gpu
def test_gpu_exploding_patterns():
    rules = exploding_rules()
    eng, want = _gpu_check(rules, exploding_requests(20_000, seed=6, long_every=97))
    assert "bitset-nfa" in eng.describe()
    assert len(set((want >> 2).tolist())) >= len(rules)


@pytest.mark.gpu
def test_gpu_tables_larger_than_shared_memory():
    """1 202 positions: 38 state words per request and 180 KB of follow rows -- read from global memory, state in local memory."""
    rules = [Rule("huge", 'http_request.url.matches("a[ab]{1200}z")', [Action.BLOCK]),
             Rule("k15", 'http_request.url.matches("a[ab]{15}c")', [Action.CAPTCHA])]
    rng = random.Random(11)
    reqs = []
    for i in range(600):
        n = rng.choice([1199, 1200, 1200, 1201, 50, 0])
        url = rng.choice(["", "q", "zb"]) + "a" + "".join(rng.choice("ab") for _ in range(n)) + rng.choice(["z", "", "qz", "c"])
        reqs.append(dict(host="h", url=url, path="/p", method="GET", user_agent="Mozilla/5.0", ip="10.0.0.1", remote_port=1, flags=0))
    eng, want = _gpu_check(rules, pack_requests(reqs), max_dfa_states=256)   # (a small cap only shortens the doomed DFA attempt)
    assert "positions=1202" in eng.describe()
    hist = np.bincount(want & 3, minlength=4)
    assert hist[1] > 50 and hist[0] > 50, hist


@pytest.mark.gpu
def test_gpu_every_pattern_forced_through_the_bitset_kernel():
    rules, reqs = scenarios.ragged()
    _gpu_check(rules, pack_requests(reqs), max_dfa_states=1, candidate_gate=False)
    rules, lists, mmdb, batch, _ = scenarios.config2_sample(20_000, attack_rate=0.2)
    eng, want = _gpu_check(rules, batch, max_dfa_states=1, candidate_gate=False)
    assert eng.info().n_scan_units == 0 and eng.info().n_bitset_units > 100 and eng.info().bitset_positions > 1000
    _gpu_check(rules, batch.slice(0, 8000), max_dfa_states=1)   # gate + literal confirmation in front, bitset units behind
