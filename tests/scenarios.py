"""Shared parity scenarios: (rules, lists, mmdb, batch, eval_gates).  Used by the GPU parity tests
(CUDA vs oracle) and by the CPU-only compiler tests (compiled tables vs oracle)."""
import numpy as np

import synth
from pingoo_b200 import Service  # noqa: E402
from pingoo_b200 import Action, ListType, Rule, pack_requests


def config1():
    rules, payloads, _ = synth.make_ruleset(16, config_id=1)
    batch = synth.RequestStream(config_id=1, payloads=payloads, get_only=True).generate(0, 10_000)
    return rules, None, None, batch, True


def config2_sample(n=60_000, first=0, attack_rate=0.03):
    rules, payloads, _ = synth.make_ruleset(128, config_id=2)
    batch = synth.RequestStream(config_id=2, payloads=payloads, attack_rate=attack_rate).generate(first, n)
    return rules, None, None, batch, True


def rules256(n=20_000):
    rules, payloads, _ = synth.make_ruleset(256, config_id=2)
    batch = synth.RequestStream(config_id=2, payloads=payloads).generate(5_000, n)
    return rules, None, None, batch, True


def ragged():
    rules = [
        Rule("eq_empty", 'http_request.path == ""', [Action.CAPTCHA]),
        Rule("end", 'http_request.url.ends_with("=")', [Action.BLOCK]),
        Rule("re_end", 'http_request.url.matches("a+b$")', [Action.BLOCK]),
        Rule("len", "http_request.host.length() == 0", [Action.BLOCK]),
        Rule("word", 'http_request.url.matches("\\\\bcat\\\\b")', [Action.CAPTCHA, Action.BLOCK]),
    ]
    reqs = []
    urls = ["", "/", "/x=", "aab", "aaba", "a cat", "concat", "cat", "x" * 15, "y" * 16, "z" * 17, "b" * 31 + "=", "q" * 4097 + "aab",
            # bytes outside printable ASCII (control characters, DEL, UTF-8): the scan leaves its byte-indexed fast path
            "a\tcat", "caf\u00e9 cat", "\x01aab", "aab\x7f", "\x7fa cat\x1f", "x" * 30 + "\u00e9" * 3 + " cat", "aa\x00b"]
    for i, u in enumerate(urls * 7):
        reqs.append(dict(host="" if i % 5 == 0 else "h.example", url=u, path="" if i % 3 == 0 else u[:40], method="GET",
                         user_agent="Mozilla/5.0 t", ip="10.0.0.%d" % (i % 250), remote_port=1000 + i, flags=i % 2))
    return rules, reqs


def gates():
    rules = [
        Rule("cap_then_block", 'http_request.path.starts_with("/a")', [Action.CAPTCHA, Action.BLOCK]),
        Rule("noop", 'http_request.path.starts_with("/b")', []),
        Rule("cap", 'http_request.path.starts_with("/b")', [Action.CAPTCHA]),
        Rule("always", None, [Action.CAPTCHA]),
        Rule("never_reached", 'http_request.path.starts_with("/c")', [Action.BLOCK]),
    ]
    reqs = []
    for path in ["/a", "/b", "/c", "/d", "/__pingoo/captcha/verify", "/__pingoo"]:
        for ua in ["Mozilla/5.0", "", "x" * 255, "y" * 256]:
            for flags in [0, 1, 2, 4, 8, 5, 9]:
                reqs.append(dict(host="h", url=path, path=path, method="GET", user_agent=ua, ip="1.1.1.1", remote_port=1, flags=flags))
    return rules, pack_requests(reqs)


def lists_geo(n=30_000, n_block=3_000, n_geo=1_500):
    csv, members = synth.make_blocklist(n_block, config_id=3)
    mmdb, records = synth.make_geoip(n_geo, config_id=3)
    lists = {"blocked_ips": (ListType.Ip, csv), "bad_asns": (ListType.Int, b"64512\n64513,\"x\"\n7\n"),
             "bad_hosts": (ListType.String, b"evil.example\n bad.example ,note\n")}
    rules, payloads, _ = synth.make_ruleset(64, config_id=3, with_lists=True)
    rules += [Rule("hosts", 'lists["bad_hosts"].contains(http_request.host)', [Action.BLOCK]),
              Rule("asn_big", "client.asn >= 60000 && client.remote_port > 60000", [Action.CAPTCHA]),
              Rule("cc", '["FR", "DE", "ZZ"].contains(client.country) && client.asn != 0', [Action.CAPTCHA]),
              Rule("missing", 'lists["nope"].contains(client.ip) || http_request.path.starts_with("/zz")', [Action.BLOCK])]
    stream = synth.RequestStream(config_id=3, payloads=payloads, blocklist_ips=members, blocklist_rate=0.05, special_ip_rate=0.01)
    batch = stream.generate(0, n)
    rng = np.random.RandomState(5)
    for i in range(0, batch.n, 3):
        net, _ = records[rng.randint(len(records))]
        host = int(net.network_address) + int(rng.randint(0, min(net.num_addresses, 1 << 30)))
        raw = host.to_bytes(4 if net.version == 4 else 16, "big")
        batch.ip[i] = np.frombuffer(raw + b"\0" * (16 - len(raw)), dtype=np.uint8)
        batch.ip_is_v6[i] = net.version == 6
    return rules, lists, mmdb, batch, True, records


def geo_probe_addresses(records, seed=11):
    rng = np.random.RandomState(seed)
    ips, v6 = [], []
    for net, _ in records:
        for _ in range(3):
            host = int(net.network_address) + int(rng.randint(0, min(net.num_addresses, 1 << 30)))
            raw = host.to_bytes(4 if net.version == 4 else 16, "big")
            ips.append(raw + b"\0" * (16 - len(raw)))
            v6.append(net.version == 6)
    for s in ["127.0.0.1", "224.0.0.1", "8.8.8.8", "255.255.255.255", "0.0.0.0", "239.1.2.3", "126.255.255.255", "128.0.0.0"]:
        ips.append(bytes(map(int, s.split("."))) + b"\0" * 12)
        v6.append(0)
    for raw in [b"\0" * 15 + b"\1", b"\xff\x02" + b"\0" * 13 + b"\1", b"\x20\x01" + b"\0" * 14, b"\0" * 12 + bytes([1, 2, 3, 4]),
                b"\0" * 10 + b"\xff\xff" + bytes([1, 2, 3, 4]), b"\xff" * 16, b"\0" * 16]:
        ips.append(raw)
        v6.append(1)
    ip_np = np.frombuffer(b"".join(ips), dtype=np.uint8).reshape(-1, 16).copy()
    return ip_np, np.array(v6, dtype=np.uint8)


def services(n=20_000, catch_all=True):
    """WAF rules of config 2 plus a service table (http_listener.rs:266-272): routes from the docs (two of the documented
    examples refer to names that do not exist and therefore never match), a regex route, a list route, a non-bool route."""
    rules, payloads, _ = synth.make_ruleset(128, config_id=2)
    # host names come from the generator's vocabulary: take frequent ones so that every live route sees traffic
    from collections import Counter

    probe = synth.RequestStream(config_id=2, payloads=payloads).generate(7_000, 3_000)
    common = [h for h, _ in Counter(probe.field("host", i) for i in range(probe.n)).most_common(40)]
    plain = [h for h in common if h.count(b".") == 1][:2]
    sub = [h.split(b".")[0] + b"." for h in common if h.count(b".") >= 2][:2]
    lists = {"static_hosts": (ListType.String, plain[0] + b"\n" + plain[1] + b",cdn\n")}
    svcs = [
        Service("api", 'http_request.host.starts_with("%s") || http_request.host.starts_with("%s")' % (sub[0].decode(), sub[1].decode())),  # getting_started.md:38 shape
        Service("doc_unknown_variable", 'host.starts_with("api")'),               # docs/services.md:18: `host` is not a variable
        Service("doc_method_on_map", 'http_request.starts_with("/api")'),         # docs/configuration.md:53
        Service("images", 'http_request.path.ends_with(".png") || http_request.path.ends_with(".svg")'),
        Service("numbered", 'http_request.path.matches("^/[0-9]+/") && client.remote_port > 1024'),
        Service("static_site", 'lists["static_hosts"].contains(http_request.host) && http_request.method == "GET"'),
        Service("non_bool", "http_request.url.length()"),
        Service("missing_list", 'lists["nope"].contains(http_request.host)'),
        Service("writes", 'http_request.method == "POST" || http_request.method == "PUT"'),
    ]
    if catch_all:
        svcs.append(Service("default"))
        svcs.append(Service("shadowed", 'http_request.method == "GET"'))
    batch = synth.RequestStream(config_id=2, payloads=payloads, attack_rate=0.05).generate(7_000, n)
    return rules, lists, svcs, batch


def value_constructs():
    """Constructs whose operands swap the usual roles: a constant receiver with a request variable as argument, list literals that
    hold request variables, integer expressions looked up in lists, byte-wise ordering between two fields."""
    exprs = [
        '"GET POST".contains(http_request.method)',
        '"/admin/x".starts_with(http_request.path)',
        '"example.com".ends_with(http_request.host)',
        '"USFR".contains(client.country) && client.remote_port == 80',
        '[80, 443].contains(client.remote_port + 1)',
        'lists["ports"].contains(client.remote_port * 2 - 80)',
        '[0].contains(client.remote_port / (client.remote_port - 442))',          # division by zero at port 442: an error, no match
        '!([7].contains(100 / (client.remote_port - 443))) && client.remote_port > 441',   # the error wins over the negation
        'http_request.path < http_request.url',
        'http_request.path >= http_request.host',
        'http_request.method <= http_request.host && http_request.method > http_request.path',
        '[http_request.method, "x"].contains("PUT")',
        '[http_request.method, client.remote_port, "GET"].contains(http_request.host)',
        '[client.remote_port - 1, 79].contains(client.remote_port + 0)',
        '[http_request.path, http_request.url][1] == "/adm"',
        '[http_request.path, 5].length() == 2 && http_request.method == ""',
        '{"GET": 1, "ET": 2}.contains(http_request.method) && client.remote_port == 79',
        'http_request.contains(http_request.method)',
        'lists.contains(http_request.host)',
        # string concatenation with request fields: decided by the ways the constant can be cut along the parts
        'http_request.host + http_request.path == "example.com/admin/x"',
        '(http_request.method + " " + http_request.path).starts_with("GET /adm")',
        '(http_request.host + "|" + http_request.method).ends_with("m|PUT")',
        '(http_request.host + http_request.path).contains("m/a") && http_request.method != "GET"',
        '(http_request.method + http_request.method) == "ETET"',
        '(http_request.host + "ab" + http_request.path).length() == 17 && client.remote_port == 443',
    ]
    rules = [Rule(f"c{i}", ex, [Action.BLOCK if i % 3 else Action.CAPTCHA]) for i, ex in enumerate(exprs)]
    reqs = [dict(host=h, url=u, path=p, method=m, user_agent="Mozilla/5.0", ip="1.2.3.4", remote_port=port, flags=(len(h) + port) % 2, country=c, asn=1)
            for h in ["example.com", "com", "", "US", "a.example.com", "ports", "GET", "host"]
            for (u, p) in [("/admin/x?y", "/admin/x"), ("/", ""), ("/adm", "/adm"), ("", ""), ("/zz/GET", "/zz")]
            for m in ["GET", "POST", "PUT", "ET", "", "host", "zz"] for port in [79, 80, 442, 443] for c in ["US", "FR", "XX"]]
    lists = {"ports": (ListType.Int, b"80\n443\n806\n")}
    return rules, lists, pack_requests(reqs)


def unreferenced_predicates():
    """A predicate whose rule folds to a constant (here: behind a `||` operand that always errors) used to be evaluated
    anyway; being mentioned by no rule it had an empty rule signature, and the two-atom verdict shortcut paired it with an atom
    that a route true on the all-false vector negates -- the wrong service (found by the expression fuzz, seed 3736)."""
    rules = [Rule("dead", '(1 + 1 || lists["nets"].contains(client.ip)) ? http_request.host >= http_request.url : http_request.nope == "x"', []),
             Rule("port", "client.remote_port == 81", [Action.BLOCK])]
    services = [Service("s0", "!(http_request.path < http_request.url)"), Service("s1", 'http_request.method == "GET"'), Service("s2", None)]
    lists = {"nets": (ListType.Ip, b"10.0.0.0/8\n")}
    reqs = [dict(host=h, url=u, path=p, method=m, user_agent="Mozilla/5.0", ip=ip, remote_port=port, flags=0)
            for h in ["", "zz", "a"] for (u, p) in [("selectcurlaab/", "selectcurlaab"), ("a", "b"), ("/x", "/x")] for m in ["GET", "DELETE"]
            for ip in ["8.8.8.8", "10.1.1.1"] for port in [80, 81]]
    return rules, services, lists, pack_requests(reqs)
