"""CPU fuzz: random regex rule sets through every way the compiler can realise a pattern -- gated DFA units, literal
confirmation, ungated DFA units, small state caps (many units) and the bit-parallel NFA unit -- against the oracle.
(`tools/fuzz_patterns.py` is the same loop with a seed and a round count on the command line.)"""
import json
import random

import numpy as np
import pytest

from helpers import Oracle, Sim
from pingoo_b200 import Action, Rule, pack_requests
from test_oracle import random_pattern

ALPHABET = "abcABCx1 2/%=.\n_-"
VARIANTS = (("gated", {}), ("no gate", dict(candidate_gate=False)), ("no literal confirmation", dict(literal_confirm=False)),
            ("bitset units only", dict(max_dfa_states=1, candidate_gate=False)), ("bitset units behind the gate", dict(max_dfa_states=1)),
            ("24-state units", dict(max_dfa_states=24)))


def one_round(rng, n_patterns=12, n_requests=300):
    pats = []
    while len(pats) < n_patterns:
        p = random_pattern(rng)
        if len(p) <= 48:
            pats.append(p)
    fields = ["url", "user_agent", "path"]
    rules = [Rule(f"r{i}", f"http_request.{fields[i % 3]}.matches(" + json.dumps(p) + ")", [Action.BLOCK if i % 2 else Action.CAPTCHA])
             for i, p in enumerate(pats)]

    def mk(n):
        return "".join(rng.choice(ALPHABET) for _ in range(rng.randint(0, n)))

    reqs = [dict(host="h", url=mk(40), path="/" + mk(20).replace("\n", "x"), method="GET", user_agent="M" + mk(30).replace("\n", " ").strip() + "z",
                 ip="1.2.3.4", remote_port=1, flags=i % 2) for i in range(n_requests)]
    batch = pack_requests(reqs)
    want = Oracle(rules, eval_gates=False).evaluate(batch, threads=8)
    problems = []
    for label, opts in VARIANTS:
        got = Sim(rules, eval_gates=False, **opts).evaluate(batch)
        d = np.nonzero(got != want)[0]
        if len(d):
            i = int(d[0])
            problems.append(f"{label}: {len(d)} differ, e.g. request {i}: oracle {int(want[i]):#x} tables {int(got[i]):#x} url={batch.field('url', i)!r} "
                            f"ua={batch.field('user_agent', i)!r} path={batch.field('path', i)!r} rules={[r.expression for r in rules]}")
    return problems


@pytest.mark.parametrize("seed", [101, 102, 103, 104, 105, 106])
def test_random_pattern_sets_on_every_unit_kind(seed):
    rng = random.Random(seed)
    for _ in range(4):
        problems = one_round(rng)
        assert not problems, "\n".join(problems)


VOC = ["curl", "curl/", "curl/8", "sqlmap", "sqlmap1", "sqlmaP2", "nikto", "nik", ".php", ".php5", ".phtml", "php", "../", "..", "/etc/passwd", "/etc/", "passwd", "wp-admin",
       "wp-", "admin", "Admin", "ADMIN", "bot", "Bot/", "robot", "abc", "abcd", "abcde", "bcde", "cde", "xyz", "xy", "yz", "-", "_", "/", "%2e", "%2E", "aaa", "aaaa", "aab",
       "baa", "a", "b", "zz"]


def literal_round(rng, n_requests=250):
    """Rule sets made of strings that share prefixes, suffixes and grams (what decides which atoms the gate may confirm itself),
    anchored and case-insensitive variants, negations; requests assembled from the same vocabulary at every alignment."""
    def esc(t):
        return "".join("\\\\" + c if c in ".+*?()[]{}|^$/\\" else c for c in t)

    rules = []
    for i in range(rng.randint(3, 14)):
        f = rng.choice(["url", "user_agent", "path"])
        k, lit = rng.randrange(8), rng.choice(VOC)
        if k < 3:
            ex = f"http_request.{f}.{['contains', 'ends_with', 'starts_with'][k]}(" + json.dumps(lit) + ")"
        elif k == 3:
            ex = f'http_request.{f}.matches("(?i)(' + "|".join(esc(rng.choice(VOC)) for _ in range(rng.randint(1, 4))) + ')")'
        elif k == 4:
            ex = f'http_request.{f}.matches("(' + "|".join(esc(rng.choice(VOC)) for _ in range(rng.randint(1, 3))) + ')$")'
        elif k == 5:
            ex = f'http_request.{f}.matches("^' + esc(lit) + '")'
        elif k == 6:
            ex = f'http_request.{f}.matches("' + esc(lit) + rng.choice(["[0-9]?", "s?"]) + esc(rng.choice(VOC)) + '")'
        else:
            ex = f"http_request.{f} == " + json.dumps(lit)
        if rng.random() < 0.15:
            ex = "!(" + ex + ")"
        rules.append(Rule(f"r{i}", ex, [Action.BLOCK if i % 2 else Action.CAPTCHA]))

    def mk():
        return "".join(rng.choice(VOC + ["q", "Q", " ", "="]) for _ in range(rng.randint(0, 7)))

    reqs = [dict(host="h", url=mk(), path="/" + mk(), method="GET", user_agent=(mk().strip() or "m") + "z", ip="1.2.3.4", remote_port=1, flags=i % 2)
            for i in range(n_requests)]
    batch = pack_requests(reqs)
    want = Oracle(rules, eval_gates=False).evaluate(batch, threads=8)
    problems = []
    for label, opts in VARIANTS[:4]:
        got = Sim(rules, eval_gates=False, **opts).evaluate(batch)
        d = np.nonzero(got != want)[0]
        if len(d):
            i = int(d[0])
            problems.append(f"{label}: {len(d)} differ, e.g. url={batch.field('url', i)!r} ua={batch.field('user_agent', i)!r} path={batch.field('path', i)!r} "
                            f"oracle {int(want[i]):#x} tables {int(got[i]):#x} rules={[r.expression for r in rules]}")
    return problems


@pytest.mark.parametrize("seed", [301, 302, 303])
def test_literal_heavy_rule_sets(seed):
    rng = random.Random(seed)
    for _ in range(12):
        problems = literal_round(rng)
        assert not problems, "\n".join(problems)
