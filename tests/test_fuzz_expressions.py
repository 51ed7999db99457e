"""Randomised rule sets: the compiled program (three-valued lowering, atoms, DFAs, candidate evaluation) must give the
oracle's verdicts and services for arbitrary combinations of predicates, including the error-producing ones (A4-A6)."""
import random

import numpy as np
import pytest

from helpers import Oracle, Sim, fmt_verdict
from pingoo_b200 import Action, ListType, Rule, Service, WafEngine, pack_requests

WORDS = ["admin", "login", ".php", "select", "union", "/api", "wp-", "etc/passwd", "<script", "x", "", "a=b", "../", "%00", "curl", "bot", "Mozilla/"]
REGEX = [r"(?i)union\s+select", r"\.(php|asp)$", r"^/api/v[0-9]+/", r"a+b", r"[0-9]{3,}", r"(?i)<script[^>]*>", r"\bcat\b", r"^$", r"x*", r"(", r"wp-(admin|login)", r"%[0-9a-fA-F]{2}",
         r"\b{start}admin", r"\p{Lu}{3,}", r"\P{L}\p{Nd}+$", r"login\>"]
FIELDS = ["host", "url", "path", "method", "user_agent"]
HOSTS = ["example.com", "api.example.com", "evil.example", "", "h"]
METHODS = ["GET", "POST", "PUT", "DELETE", ""]
UAS = ["Mozilla/5.0 (X11)", "curl/8.0", "bot", "Mozilla/4.0 bot", "x" * 40]


def lit(rng):
    return '"' + rng.choice(WORDS).replace('"', '') + '"'


def predicate(rng):
    f = "http_request." + rng.choice(FIELDS)
    k = rng.randrange(31)
    if k < 4:
        return f'{f}.contains({lit(rng)})'
    if k < 6:
        return f'{f}.starts_with({lit(rng)})'
    if k < 8:
        return f'{f}.ends_with({lit(rng)})'
    if k < 10:
        return f'{f} {rng.choice(["==", "!="])} {lit(rng)}'
    if k < 13:
        return f'{f}.matches("{rng.choice(REGEX)}")'.replace("\\", "\\\\")
    if k == 13:
        return f'{f}.length() {rng.choice(["<", "<=", ">", ">=", "==", "!="])} {rng.randrange(0, 40)}'
    if k == 14:
        return f'client.remote_port {rng.choice(["<", "<=", ">", ">=", "==", "!="])} {rng.choice([0, 80, 443, 1024, 65535])}'
    if k == 15:
        return f'client.asn {rng.choice(["==", "!=", ">"])} {rng.choice([0, 64512, 7])}'
    if k == 16:
        return rng.choice(['client.country == "FR"', '["FR", "DE"].contains(client.country)', 'client.country != "XX"'])
    if k == 17:
        return rng.choice(['lists["nets"].contains(client.ip)', 'lists["ports"].contains(client.remote_port)', 'lists["names"].contains(http_request.host)'])
    if k == 18:  # evaluation errors: missing key, cross-type comparison, method on the wrong type
        return rng.choice(['lists["nope"].contains(client.ip)', 'client.ip == "1.2.3.4"', 'http_request.host == 1', 'http_request.nope == "x"',
                           'client.remote_port.contains("1")', '1 / 0 == 1'])
    if k == 19:
        return rng.choice(["true", "false", "1 == 1", '"a" < "b"', "1 + 1"])  # the last one is not a Bool
    if k == 20:
        return f'lists["names"].contains({lit(rng)})'
    if k == 21:
        return f'{f}.contains({lit(rng)}) == {rng.choice(["true", "false"])}'
    g = "http_request." + rng.choice(FIELDS)
    if k == 22:   # one field against another
        return rng.choice([f'{f} == {g}', f'{f} != {g}', f'{f}.contains({g})', f'{f}.starts_with({g})', f'{f}.ends_with({g})', f'{f} < {g}', f'{f} >= {g}'])
    if k == 23:   # integer arithmetic on request values, with its overflow / division errors
        return rng.choice(['client.remote_port + 1 > http_request.url.length() * 2', 'client.remote_port / (client.asn - 7) == 0',
                           'client.remote_port % 7 == client.asn % 7', '-client.remote_port < 0 - http_request.host.length()',
                           'client.remote_port * 140737488355328 * 65536 > 0'])
    if k == 24:   # a constant receiver, the request variable as argument
        return rng.choice(['"GET POST".contains(http_request.method)', '"api.example.com".ends_with(http_request.host)', f'{lit(rng)}.starts_with({f})',
                           '"FRDE".contains(client.country)', f'{lit(rng)}.contains({f})'])
    if k == 25:   # list literals that hold request variables
        return rng.choice([f'[{f}, "x"].contains({lit(rng)})', f'[{f}, client.remote_port, "GET"].contains({g})', '[client.remote_port + 1, 80].contains(client.asn)',
                           f'[{f}, {g}][1] == {lit(rng)}', f'[{f}].length() == 1'])
    if k == 26:   # integer expressions looked up in lists
        return rng.choice(['lists["ports"].contains(client.remote_port + 363)', '[81, 444].contains(client.remote_port + 1)',
                           '[0].contains(client.remote_port / (client.asn - 7))', 'lists["names"].contains(client.remote_port + 1)'])
    if k == 27:   # key presence
        return rng.choice([f'http_request.contains({f})', f'lists.contains({f})', f'{{"GET": 1, "admin": 2}}.contains({f})'])
    if k == 28:   # conditional with non-boolean branches
        return rng.choice([f'(client.remote_port > 100 ? {f} : {g}) == {lit(rng)}', f'(client.asn == 7 ? "GET" : "POST") == http_request.method',
                           f'(client.remote_port == 80 ? 1 : client.asn) > 5'])
    if k == 29:   # string concatenation with request fields, compared with constants
        cat = "(" + " + ".join(rng.choice([f, g, lit(rng), '"/"', "http_request.method"]) for _ in range(rng.randint(2, 3))) + " + " + f + ")"
        return rng.choice([f'{cat} == {lit(rng)}', f'{cat}.contains({lit(rng)})', f'{cat}.starts_with({lit(rng)})', f'{cat}.ends_with({lit(rng)})',
                           f'{cat}.length() > {rng.randrange(0, 60)}', f'{cat} != "GETGET"', f'(http_request.method + " " + http_request.path).starts_with("GET /api")'])
    return f'{f}.contains({lit(rng)}) == {rng.choice(["true", "false"])}'


def expr(rng, depth):
    if depth == 0 or rng.random() < 0.3:
        return predicate(rng)
    k = rng.randrange(6)
    if k == 0:
        return "!(" + expr(rng, depth - 1) + ")"
    if k in (1, 2):
        return "(" + expr(rng, depth - 1) + " && " + expr(rng, depth - 1) + ")"
    if k in (3, 4):
        return "(" + expr(rng, depth - 1) + " || " + expr(rng, depth - 1) + ")"
    return "(" + expr(rng, depth - 1) + " ? " + expr(rng, depth - 1) + " : " + expr(rng, depth - 1) + ")"


def make_case(seed, n_rules=12, n_services=4, n_requests=300):
    rng = random.Random(seed)
    acts = [[Action.BLOCK], [Action.CAPTCHA], [Action.CAPTCHA, Action.BLOCK], []]
    rules = [Rule(f"r{i}", None if rng.random() < 0.03 else expr(rng, 3), rng.choice(acts)) for i in range(n_rules)]
    svcs = [Service(f"s{i}", None if rng.random() < 0.1 else expr(rng, 2)) for i in range(n_services)]
    lists = {"nets": (ListType.Ip, b"10.0.0.0/8\n192.168.1.7\n2001:db8::/32\n"), "ports": (ListType.Int, b"80\n443\n"),
             "names": (ListType.String, b"evil.example\nadmin\n")}  # (ports + 363: 443 - 80)
    reqs = []
    for _ in range(n_requests):
        parts = [rng.choice(WORDS + ["/", "?q=", "123", "aab", " cat ", "UNION  SELECT", "%2e"]) for _ in range(rng.randrange(0, 6))]
        url = "".join(parts)
        reqs.append(dict(host=rng.choice(HOSTS), url=url, path=url.split("?")[0], method=rng.choice(METHODS), user_agent=rng.choice(UAS),
                         ip=rng.choice(["10.1.2.3", "192.168.1.7", "8.8.8.8", "2001:db8::1", "::1"]), remote_port=rng.choice([0, 80, 443, 1024, 40000, 65535]),
                         asn=rng.choice([0, 7, 64512]), country=rng.choice(["FR", "DE", "US", "XX"]), flags=rng.choice([0, 0, 0, 1, 2, 4, 8])))
    return rules, svcs, lists, pack_requests(reqs)


def _compare(build, seeds):
    compiled = skipped = 0
    for seed in seeds:
        rules, svcs, lists, batch = make_case(seed)
        want_v, want_s = Oracle(rules, lists, services=svcs).evaluate_routed(batch, threads=4)
        try:
            got_v, got_s = build(rules, lists, svcs, batch)
        except Exception as e:  # constructs the engine refuses loudly (SEMANTICS.md) are allowed, silent differences are not
            assert any(k in str(e) for k in ("not supported", "cannot be expressed", "needs a DFA larger", "unsupported", "too deeply")), (seed, str(e))
            skipped += 1
            continue
        compiled += 1
        bad = np.nonzero((got_v != want_v) | (got_s != want_s))[0]
        assert len(bad) == 0, "\n".join(
            [f"seed {seed}: {len(bad)} of {batch.n} differ"] +
            [f"  req {i}: oracle {fmt_verdict(want_v[i])}/svc {want_s[i]} got {fmt_verdict(got_v[i])}/svc {got_s[i]} url={batch.field('url', i)!r} host={batch.field('host', i)!r}" for i in bad[:4]] +
            [f"  {r.name}: {r.expression} -> {[int(a) for a in r.actions]}" for r in rules] + [f"  {s.name}: {s.route}" for s in svcs])
    assert compiled >= len(seeds) // 2, f"only {compiled} of {len(seeds)} random rule sets compiled ({skipped} refused)"


def test_random_rule_sets_compiled_tables_vs_oracle():
    _compare(lambda rules, lists, svcs, batch: Sim(rules, lists, services=svcs).evaluate_routed(batch), range(60))


@pytest.mark.gpu
def test_random_rule_sets_gpu_vs_oracle():
    _compare(lambda rules, lists, svcs, batch: WafEngine(rules, lists, device=0, services=svcs).evaluate_host_routed(batch), range(100, 125))
