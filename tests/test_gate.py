"""The candidate gate (pingoo_b200/csrc/gate.{hpp,cpp}) is a prefilter: a request the gate does not flag is never
walked by the gated DFAs, so the gate must flag EVERY request in which a gated pattern matches.  These CPU tests check
that soundness claim against the oracle (through the test-only table walk, which applies gate, early exit and the
clean / single-atom verdict tables exactly as the kernels do), with the gate on and off, on inputs built to hit the
corner cases of the argument in gate.hpp: matches at even and odd column positions, at the first and last bytes of a
field, matches shorter than a window, case variations, whitespace classes, patterns that begin with a gap, neighbours
whose bytes complete a gram across a field boundary."""
import numpy as np
import pytest

import synth
from helpers import Oracle, Sim
from pingoo_b200 import Action, Rule, pack_requests


def _same(rules, batch, lists=None):
    want = Oracle(rules, lists).evaluate(batch, threads=8)
    on = Sim(rules, lists).evaluate(batch)
    off = Sim(rules, lists, candidate_gate=False).evaluate(batch)
    bad = np.nonzero(on != want)[0]
    assert len(bad) == 0, f"gate on: {len(bad)} differ, first url={batch.field('url', int(bad[0]))!r} ua={batch.field('user_agent', int(bad[0]))!r}"
    assert np.array_equal(off, want)
    return want


RULES = [
    Rule("lit5", 'http_request.url.contains("union")', [Action.BLOCK]),
    Rule("lit3", 'http_request.url.contains("../")', [Action.BLOCK]),
    Rule("lit4end", 'http_request.url.ends_with(".php")', [Action.CAPTCHA]),
    Rule("ci", 'http_request.url.matches("(?i)sel[e3]ct\\\\s+from")', [Action.BLOCK]),
    Rule("short_alt", 'http_request.url.matches("(;|\\\\|)\\\\s*(nc|id)\\\\b")', [Action.BLOCK]),
    Rule("gap_first", 'http_request.url.matches("[0-9]*\\\\.env$")', [Action.BLOCK]),
    Rule("quote", 'http_request.url.matches("(%27|\')\\\\s*(or|and)\\\\s+\\\\d+=\\\\d+")', [Action.BLOCK]),
    Rule("tag", 'http_request.url.matches("(?i)<svg[^>]*>")', [Action.CAPTCHA]),
    Rule("ua_lit", 'http_request.user_agent.contains("curl/")', [Action.CAPTCHA]),
    Rule("ua_neg", '!http_request.user_agent.starts_with("Mozilla/") && http_request.user_agent.contains("bot")', [Action.BLOCK]),
    Rule("path_pre", 'http_request.path.starts_with("/adm")', [Action.BLOCK]),
    Rule("path_dot", 'http_request.path.matches("(?i)\\\\.(git|env)(/|$)")', [Action.BLOCK]),
    Rule("one_byte", 'http_request.url.contains("\\u0001")', [Action.CAPTCHA]),   # cannot be gated: stays an ungated unit
    Rule("empty_ok", 'http_request.path.matches("^a*$")', [Action.CAPTCHA]),     # matches the empty field
]


def test_units_are_classified():
    d = Sim(RULES).describe()
    assert "url/gated" in d and "gate(url:" in d and "gate(user_agent:" in d and "gate(path:" in d
    # the user-agent and path patterns of this set are finite string sets: confirmed by the gate, no gated unit left
    assert "literals=" in d and "user_agent/gated" not in d and "path/gated" not in d
    assert "[url:" in d  # the one-byte pattern keeps an ungated url unit
    off = Sim(RULES, candidate_gate=False).describe()
    assert "gated" not in off and "gate(" not in off


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_matches_everywhere(seed):
    rng = np.random.RandomState(100 + seed)
    frag = ["union", "UNION", "uNiOn", "unio", "nion", "../", "..", "./", ".php", ".ph", "select  from", "SEL3CT\tFROM", "sel", ";nc", "| id", ";  nc ", ";ncx", "|id",
            "7.env", ".env", "12.envx", "%27 or 1=1", "' AND 22=2", "%27or", "'  or  9=", "<svg x>", "<SVG", "<sv>", ">", "a", "b", "/", "?", "=", "&", "x", "\x01",
            "curl/", "cur", "bot", "Bot", "Mozilla/", "Mozilla", "/adm", "/ad", ".git", ".GIT/", ".gi", " ", "\t", "%2", "9", "\n"]
    reqs = []
    for i in range(4000):
        k = rng.randint(0, 7)
        url = "".join(frag[rng.randint(len(frag))] for _ in range(k))
        ua = "".join(frag[rng.randint(len(frag))] for _ in range(rng.randint(1, 4)))
        path = "".join(frag[rng.randint(len(frag))] for _ in range(rng.randint(0, 3)))
        reqs.append(dict(host="h", url=url, path=path, method="GET", user_agent=ua or "x", ip="1.2.3.4", remote_port=1, flags=0))
    want = _same(RULES, pack_requests(reqs))
    assert len(set(want.tolist())) > 8  # most rules decide something


def test_every_alignment_and_boundary():
    """One matching fragment at every position of short fields: every column alignment, first / last bytes, and
    neighbours that end / begin with halves of a gram (a gram completed across a boundary must change nothing)."""
    reqs = []
    for frag in ["union", "../", ".php", ";nc", "'or 1=1", "<svg>", "select from", "3.env"]:
        for pre in range(0, 9):
            for post in range(0, 5):
                url = "q" * pre + frag + "r" * post
                reqs.append(dict(host="h", url=url, path="/" + "p" * (pre % 3), method="GET", user_agent="Mozilla/5.0 z" + "y" * post, ip="1.2.3.4", remote_port=1, flags=0))
        # halves: the previous field ends with the first half, the next one begins with the second
        for cut in range(1, len(frag)):
            reqs.append(dict(host="h", url="zz" + frag[:cut], path="/", method="GET", user_agent="Mozilla/5.0", ip="1.2.3.4", remote_port=1, flags=0))
            reqs.append(dict(host="h", url=frag[cut:] + "zz", path="/", method="GET", user_agent="Mozilla/5.0", ip="1.2.3.4", remote_port=1, flags=0))
    reqs += [dict(host="h", url="", path="", method="GET", user_agent="Mozilla/5.0", ip="1.2.3.4", remote_port=1, flags=0)] * 3
    want = _same(RULES, pack_requests(reqs))
    assert np.count_nonzero(want & 3) > 300


def test_gate_falls_back_when_grams_do_not_fit():
    """A field whose patterns would need more grams than the budget keeps the excess in ungated units: same verdicts."""
    rules = [Rule(f"r{i}", 'http_request.url.matches("[a-f][0-9]%c[x-z]q")' % chr(ord("g") + i), [Action.BLOCK]) for i in range(12)]
    rules.append(Rule("wide", 'http_request.url.matches("[a-z][a-z][0-9][a-z]k")', [Action.CAPTCHA]))  # thousands of grams
    rng = np.random.RandomState(3)
    alphabet = "abcdefghijklmnopqrstuvwxyz0123456789"
    reqs = [dict(host="h", url="".join(alphabet[rng.randint(36)] for _ in range(rng.randint(0, 40))), path="/", method="GET", user_agent="Mozilla/5.0",
                 ip="1.2.3.4", remote_port=1, flags=0) for _ in range(5000)]
    want = _same(rules, pack_requests(reqs))
    assert np.count_nonzero(want & 3) > 10


@pytest.mark.parametrize("n_rules,cfg", [(128, 2), (512, 3), (1024, 4)])
def test_baseline_rule_sets(n_rules, cfg):
    """The BASELINE.json rule sets on their own request streams (attack rate raised so that many patterns fire)."""
    rules, payloads, _ = synth.make_ruleset(n_rules, config_id=cfg)
    batch = synth.RequestStream(config_id=cfg, payloads=payloads, attack_rate=0.25).generate(1_000, 12_000)
    want = _same(rules, batch)
    assert len(set((want >> 2).tolist())) > n_rules // 8


def test_literal_confirmation_is_exclusive():
    """An atom is confirmed by the gate only if no other atom of the field shares one of its grams (one comparison per gram hit,
    DESIGN.md section 5): the synthetic BASELINE rule sets are families of strings with common prefixes, so configs 3 and 4 confirm
    nothing (narrow table slots, the block-append resolve kernel -- the measured-best state), config 2 one user-agent atom."""
    for n_rules, cfg, want_literals in ((128, 2, True), (512, 3, False), (1024, 4, False)):
        rules, _, _ = synth.make_ruleset(n_rules, config_id=cfg)
        d = Sim(rules).describe()
        assert ("literals=" in d) == want_literals and ("wide-slots" in d) == want_literals, (cfg, d[-400:])
    # a family: the second member shares the grams of the first -> both stay with the DFA unit; a loner is confirmed
    fam = [Rule("a", 'http_request.user_agent.contains("sqlmap")', [Action.BLOCK]), Rule("b", 'http_request.user_agent.contains("sqlmap/2")', [Action.BLOCK]),
           Rule("c", 'http_request.user_agent.contains("nikto")', [Action.CAPTCHA])]
    d = Sim(fam).describe()
    assert "literals=1" in d and "user_agent/gated" in d, d
