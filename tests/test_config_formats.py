"""Config / wire formats of the rule path: a Pingoo configuration directory is consumed unchanged."""
import ctypes as C
import ctypes.util
import os

import numpy as np
import pytest

import synth
import json

import yaml

from helpers import Oracle, Sim, yaml_dump
from pingoo_b200 import Action, Error, ListType
from pingoo_b200.config import load_config, zstd_decode_all

PINGOO_YML = """
listeners:
  http:
    address: http://0.0.0.0:8080
    services: ["webapp", "api"]
  raw:
    address: tcp://0.0.0.0:5432
services:
  db:                       # a TCP service listed first: never offered to HTTP requests (config.rs:217-221)
    tcp_proxy: ["10.0.0.5:5432"]
  api:
    route: http_request.host.starts_with("api.")
    http_proxy: []
  images:
    route: http_request.path.ends_with(".png")
    static:
      root: /var/www
  webapp:
    static:
      root: /var/www
lists:
  blocked_ips:
    type: Ip
    file: {blocked}
  bad_asns:
    type: Int
    file: {asns}
rules:
  captcha_bots:
    expression: |
      !http_request.user_agent.starts_with("Mozilla/") && !http_request.user_agent.contains("curl/")
    actions:
      - action: captcha
  blocked:
    expression: lists["blocked_ips"].contains(client.ip) || lists["bad_asns"].contains(client.asn)
    actions:
      - action: block
"""
RULES_A = """
env_files:
  expression: http_request.path.ends_with(".env") || http_request.path.matches("(?i)/\\\\.git/")
  actions: [{action: block}]
no_expression:
  actions: []
"""


def _zstd_compress(data: bytes) -> bytes:
    z = C.CDLL(ctypes.util.find_library("zstd") or "libzstd.so.1")
    z.ZSTD_compressBound.restype = C.c_size_t
    z.ZSTD_compress.restype = C.c_size_t
    buf = C.create_string_buffer(z.ZSTD_compressBound(C.c_size_t(len(data))))
    n = z.ZSTD_compress(buf, C.c_size_t(len(buf)), data, C.c_size_t(len(data)), 3)
    return buf.raw[:n]


def _write_tree(tmp, geo_zst=True):
    csv, members = synth.make_blocklist(300, config_id=3)
    mmdb, records = synth.make_geoip(200, config_id=3)
    (tmp / "lists").mkdir()
    (tmp / "lists" / "blocked.csv").write_bytes(csv)
    (tmp / "lists" / "asns.csv").write_bytes(b"64512\n64513,\"note\"\n")
    (tmp / "pingoo.yml").write_text(PINGOO_YML.format(blocked=tmp / "lists" / "blocked.csv", asns=tmp / "lists" / "asns.csv"))
    (tmp / "rules").mkdir()
    (tmp / "rules" / "a.yml").write_text(RULES_A)
    (tmp / "rules" / "ignored.yaml").write_text("x: {actions: []}")
    (tmp / ("geoip.mmdb.zst" if geo_zst else "geoip.mmdb")).write_bytes(_zstd_compress(mmdb) if geo_zst else mmdb)
    return members, mmdb


def test_directory_is_loaded_like_the_reference_does(tmp_path):
    members, mmdb = _write_tree(tmp_path)
    cfg = load_config(str(tmp_path), geoip_dirs=[str(tmp_path)])
    assert [r.name for r in cfg.rules] == ["captcha_bots", "blocked", "env_files", "no_expression"]  # file first, folder appended
    assert cfg.rules[0].actions == [Action.CAPTCHA] and cfg.rules[3].expression is None and cfg.rules[3].actions == []
    assert [(s.name, s.route) for s in cfg.services] == [("api", 'http_request.host.starts_with("api.")'),
                                                         ("images", 'http_request.path.ends_with(".png")'), ("webapp", None)]
    assert set(cfg.lists) == {"blocked_ips", "bad_asns"} and cfg.lists["blocked_ips"][0] == ListType.Ip
    assert cfg.geoip_path.endswith("geoip.mmdb.zst") and cfg.geoip_mmdb == mmdb
    # the loaded objects drive the compiler and the oracle to the same answers
    stream = synth.RequestStream(config_id=3, payloads=[], blocklist_ips=members, blocklist_rate=0.1)
    batch = stream.generate(0, 3_000)
    batch.asn = None
    batch.country = None  # resolved from the loaded database
    want_v, want_s = Oracle(cfg.rules, cfg.lists, cfg.geoip_mmdb, services=cfg.services).evaluate_routed(batch, threads=4)
    got_v, got_s = Sim(cfg.rules, cfg.lists, cfg.geoip_mmdb, services=cfg.services).evaluate_routed(batch)
    assert np.array_equal(got_v, want_v) and np.array_equal(got_s, want_s)
    assert len(set((want_v & 3).tolist())) >= 2 and 2 in set(want_s.tolist())  # blocked / captcha / allowed, catch-all service used
    # the engine's own (C++) loader reads the same directory to the same program
    nat = Sim.from_config_dir(str(tmp_path), geoip_dir=str(tmp_path))
    nat_v, nat_s = nat.evaluate_routed(batch)
    assert np.array_equal(nat_v, want_v) and np.array_equal(nat_s, want_s)
    assert nat.describe() == Sim(cfg.rules, cfg.lists, cfg.geoip_mmdb, services=cfg.services).describe()
    # a listener's own `services:` list, in its order: webapp (no route) shadows api
    lis = load_config(str(tmp_path), geoip_dirs=[str(tmp_path)], listener="http")
    assert [s.name for s in lis.services] == ["webapp", "api"]
    lv, ls = Oracle(lis.rules, lis.lists, lis.geoip_mmdb, services=lis.services).evaluate_routed(batch, threads=4)
    nv, ns = Sim.from_config_dir(str(tmp_path), listener="http", geoip_dir=str(tmp_path)).evaluate_routed(batch)
    assert np.array_equal(nv, lv) and np.array_equal(ns, ls) and set(ls[(lv & 3) == 0].tolist()) == {0}


@pytest.mark.gpu
def test_engine_loads_a_configuration_directory_through_the_c_abi(tmp_path):
    """pgw_ruleset_load_dir (csrc/config_dir.cpp) -> finalize -> evaluate: the documented example directory, GeoIP from the
    compressed database, both service sets (default and the listener's own list), against the oracle fed by the Python loader."""
    from pingoo_b200 import WafEngine

    members, mmdb = _write_tree(tmp_path)
    d = str(tmp_path)
    stream = synth.RequestStream(config_id=3, payloads=[], blocklist_ips=members, blocklist_rate=0.1)
    batch = stream.generate(0, 20_000)
    batch.asn = None
    batch.country = None
    for listener in (None, "http"):
        cfg = load_config(d, geoip_dirs=[d], listener=listener)
        want_v, want_s = Oracle(cfg.rules, cfg.lists, cfg.geoip_mmdb, services=cfg.services).evaluate_routed(batch, threads=os.cpu_count() or 1)
        eng = WafEngine.from_config_dir(d, listener=listener, geoip_dirs=[d], device=0)
        got_v, got_s = eng.evaluate_host_routed(batch)
        assert np.array_equal(got_v, want_v) and np.array_equal(got_s, want_s), listener
        assert eng.info().n_rules == 4 and eng.info().geoip_loaded == 1
    with pytest.raises(Error, match="error reading config file"):
        WafEngine.from_config_dir(str(tmp_path / "nope"))


def test_configuration_errors(tmp_path):
    _write_tree(tmp_path, geo_zst=False)
    assert load_config(str(tmp_path), geoip_dirs=[str(tmp_path)]).geoip_path.endswith("geoip.mmdb")
    (tmp_path / "rules" / "b.yml").write_text("captcha_bots: {actions: []}")
    with pytest.raises(Error, match="duplicate rule name: captcha_bots"):
        load_config(str(tmp_path))
    (tmp_path / "rules" / "b.yml").write_text('broken: {expression: "http_request.path ==", actions: []}')
    with pytest.raises(Error, match="error parsing rules: Expression is not valid"):
        load_config(str(tmp_path))
    (tmp_path / "rules" / "b.yml").write_text("x: {actions: [{action: drop}]}")
    with pytest.raises(Error, match="unknown variant `drop`, expected `block` or `captcha`"):
        load_config(str(tmp_path))
    os.remove(tmp_path / "rules" / "b.yml")
    (tmp_path / "geoip.mmdb.zst").write_bytes(b"not zstd")
    os.remove(tmp_path / "geoip.mmdb")
    with pytest.raises(Error, match="error decompressing geoip database"):
        load_config(str(tmp_path), geoip_dirs=[str(tmp_path)])
    with pytest.raises(Error, match="error reading config file"):
        load_config(str(tmp_path / "nope"))


def test_native_loader_reports_the_same_configuration_errors(tmp_path):
    _write_tree(tmp_path, geo_zst=False)
    d = str(tmp_path)
    Sim.from_config_dir(d, geoip_dir=d)
    cases = [
        ("rules/b.yml", "captcha_bots: {actions: []}", "duplicate rule name: captcha_bots"),
        ("rules/b.yml", 'broken: {expression: "http_request.path ==", actions: []}', "error parsing rules: Expression is not valid"),
        ("rules/b.yml", "x: {actions: [{action: drop}]}", "unknown variant `drop`, expected `block` or `captcha`"),
        ("rules/b.yml", "x: {expression: \"true\"}", "missing field `actions`"),
        ("rules/b.yml", "x: [1, 2", "error parsing rules file"),
    ]
    for rel, text, msg in cases:
        (tmp_path / rel).write_text(text)
        with pytest.raises(ValueError, match=msg):
            Sim.from_config_dir(d, geoip_dir=d)
        with pytest.raises(Error, match=msg.replace("rules file", "rules file")):
            load_config(d, geoip_dirs=[d])
    os.remove(tmp_path / "rules" / "b.yml")
    base = (tmp_path / "pingoo.yml").read_text()
    for patch, msg in [
        (("    http_proxy: []", "    http_proxy: []\n    static: {root: /x}"), "services must have exactly 1 http_proxy, tcp_proxy or static field"),
        (('    tcp_proxy: ["10.0.0.5:5432"]', '    tcp_proxy: ["10.0.0.5:5432"]\n    route: "true"'), "TCP proxy can't have a route"),
        (('http_request.path.ends_with(".png")', 'http_request.path.ends_with('), "error parsing route for service images"),
        (("type: Int", "type: Float"), "unknown variant `Float`, expected one of `String`, `Int`, `Ip`"),
        (('services: ["webapp", "api"]', 'services: ["webapp", "nope"]'), "service nope doesn't exist"),
    ]:
        (tmp_path / "pingoo.yml").write_text(base.replace(*patch))
        with pytest.raises(ValueError, match=msg):
            Sim.from_config_dir(d, listener="http", geoip_dir=d)
        with pytest.raises(Error, match=msg):
            load_config(d, geoip_dirs=[d], listener="http")
    (tmp_path / "pingoo.yml").write_text(base)
    (tmp_path / "geoip.mmdb.zst").write_bytes(b"not zstd")
    os.remove(tmp_path / "geoip.mmdb")
    with pytest.raises(ValueError, match="error decompressing geoip database"):
        Sim.from_config_dir(d, geoip_dir=d)
    with pytest.raises(ValueError, match="error reading config file .*No such file or directory \\(os error 2\\)"):
        Sim.from_config_dir(str(tmp_path / "nope"))


def _canon(x):
    if isinstance(x, dict):
        return "{" + ",".join(json.dumps(str(k)) + ":" + _canon(v) for k, v in x.items()) + "}"
    if isinstance(x, list):
        return "[" + ",".join(_canon(v) for v in x) + "]"
    if x is None:
        return "null"
    if isinstance(x, bool):
        return json.dumps("true" if x else "false")
    return json.dumps(str(x), ensure_ascii=False)


YAML_SAMPLES = [
    PINGOO_YML.format(blocked="/tmp/b.csv", asns="/tmp/a.csv"),
    RULES_A,
    "a: 1\nb:\n  - x\n  - y: 2\n    z: [p, 'q r', \"s\\tt\"]\n  -\n    - nested\n    - seq\nc: {k: v, l: [1, 2]}\nd: ~\ne:\n",
    "key: |\n  line one\n    indented\n\n  line three\nfold: >-\n  a b\n  c\n\n  d\nkeep: |+\n  x\n\nafter: 'it''s'\n",
    "# comment\n---\nrules:\n  r1:   # trailing comment\n    expression: http_request.url.contains(\"#not a comment\")\n    actions:\n    - action: block\n    - action: captcha\n",
    "plain: multi\n  line scalar\n  continues\nurl: http://x/y#frag\n\"quoted key\": [a,\n  b,\n  c]\n",
    "seq_at_same_indent:\n- a\n- b\nother: 3\n",
    "",
]


@pytest.mark.parametrize("text", YAML_SAMPLES)
def test_yaml_reader_agrees_with_pyyaml(text):
    ok, out = yaml_dump(text)
    assert ok, out
    assert out == _canon(yaml.safe_load(text))


YAML_WORDS = ['http_request.url.contains("x")', "block", "captcha", "a b", "it's", "x: y", "# no", "-dash", "1", "true", "null", "~", "", " lead", "trail ",
              "multi\nline", "tab\there", "100%", "[a]", "{b}", "a,b", '"q"', "k#v", "path/to/file.csv", "0x1F", "1e3", "\u00e9", "yes", "No", "*star", "&amp", "!bang",
              "|pipe", ">gt", "@at", "`tick", "%pct", "?q", ":colon", "a: ", "- x"]
YAML_KEYS = ["rules", "services", "lists", "name", "expression", "actions", "action", "route", "http_proxy", "static", "root", "file", "type", "listeners", "address",
             "a b", "k-1", "x_y", "K"]


@pytest.mark.parametrize("seed", range(8))
def test_yaml_reader_agrees_with_pyyaml_on_random_documents(seed):
    """Random nested documents written by PyYAML in block, flow and mixed styles, narrow and wide, with plain / quoted / literal /
    folded scalars (multi-line quoted scalars, apostrophes in plain scalars, indicator characters, empty collections): the
    engine's reader must read what PyYAML reads.  Styles that make PyYAML emit tags (`!!int "1"`) are refused by design."""
    import random

    rng = random.Random(seed)

    def val(d):
        r = rng.random()
        if d <= 0 or r < 0.45:
            return rng.choice(YAML_WORDS) if rng.random() < 0.8 else rng.choice([1, 0, -3, True, False, None, 2.5])
        if r < 0.75:
            return {rng.choice(YAML_KEYS) + (str(rng.randrange(3)) if rng.random() < 0.5 else ""): val(d - 1) for _ in range(rng.randrange(0, 4))}
        return [val(d - 1) for _ in range(rng.randrange(0, 4))]

    read = refused = 0
    for _ in range(150):
        doc = {rng.choice(YAML_KEYS) + str(k): val(3) for k in range(rng.randrange(1, 4))}
        text = yaml.safe_dump(doc, default_flow_style=rng.choice([True, False, None]), default_style=rng.choice([None, None, None, '"', "'", "|", ">"]),
                              width=rng.choice([20, 80, 1000]), indent=rng.choice([2, 4]), allow_unicode=rng.random() < 0.5, sort_keys=False)
        ok, out = yaml_dump(text)
        if not ok:
            assert "tags are not supported" in out and "!!" in text, (out, text)
            refused += 1
            continue
        assert out == _canon(yaml.safe_load(text)), text
        read += 1
    assert read > 80


@pytest.mark.parametrize("text", ["a: [1, 2", "a: 'x", "a: &anchor 1", "a: *alias", "a:\n    b: 1\n  c: 2\n", "- a\nb: 1\n", "a: 1\na: 2\n", "? complex\n: key\n", "next: value with: colon inside\n"])
def test_yaml_reader_rejects_what_it_does_not_read(text):
    ok, out = yaml_dump(text)
    assert not ok and out.startswith("line "), out


def test_zstd_round_trip():
    data = bytes(range(256)) * 4000 + b"tail"
    assert zstd_decode_all(_zstd_compress(data)) == data


CFG_EXPRS = ['http_request.path.starts_with("/a")', 'http_request.url.contains("x: y")', "client.remote_port == 80", 'lists["l1"].contains(client.ip)',
             'http_request.host == "it\'s"', 'http_request.url.matches("a|b")', '!(http_request.method == "GET")', "http_request.path ==", "1 +",
             'http_request.user_agent.contains("# not a comment")', 'lists["nope"].contains(client.ip)', "true", '"multi\nline" == http_request.host']


@pytest.mark.parametrize("seed", range(4))
def test_random_directories_python_and_native_loader_agree(seed, tmp_path):
    """Random configuration directories -- rules in the file and in rules/*.yml, services of every kind, lists, a listener's own
    service list, written by PyYAML in assorted styles, with the usual mistakes sprinkled in (unknown action, duplicate rule,
    route that does not compile, TCP service with a route, two service kinds, missing list file, malformed list entry): the
    Python loader and the engine's C++ loader produce the same program, or fail with the same message (the reference's order of
    checks: file, rules folder, duplicates, services with their routes, listeners, rules; config.rs:194-269)."""
    import random
    import re
    import shutil

    rng = random.Random(seed)
    acts = ["block", "captcha", "Block", "allow", ""]

    def rule():
        r = {}
        if rng.random() < 0.9:
            r["expression"] = rng.choice(CFG_EXPRS)
        if rng.random() < 0.9:
            r["actions"] = [{"action": rng.choice(acts[:2] if rng.random() < 0.93 else acts)} for _ in range(rng.randrange(0, 3))]
        return r

    def service():
        sv = {}
        if rng.random() < 0.6:
            sv["route"] = rng.choice(CFG_EXPRS[:7] + CFG_EXPRS[9:])
        for k in rng.sample(["http_proxy", "static", "tcp_proxy"], 1 if rng.random() < 0.93 else rng.choice([0, 2])):
            sv[k] = ["10.0.0.1:80"] if k != "static" else {"root": "/var/www"}
        return sv

    agree_ok = agree_err = 0
    for k in range(60):
        d = tmp_path / f"c{k}"
        d.mkdir()
        doc = {}
        names = [f"r{i}" for i in range(6)]
        if rng.random() < 0.9:
            doc["rules"] = {rng.choice(names): rule() for _ in range(rng.randrange(0, 4))}
        if rng.random() < 0.8:
            doc["services"] = {f"s{i}": service() for i in range(rng.randrange(0, 4))}
        if rng.random() < 0.7:
            (d / "l1.csv").write_text(rng.choice(['10.0.0.0/8\n1.2.3.4,"c"\n', "10.0.0.0/8\n", "bad\n", ""]))
            doc["lists"] = {"l1": {"type": rng.choice(["Ip", "Ip", "Ip", "String", "Int", "ip"]), "file": str(d / ("l1.csv" if rng.random() < 0.9 else "missing.csv"))}}
        if rng.random() < 0.3:
            doc["listeners"] = {"http": {"address": "http://0.0.0.0:80", "services": [rng.choice(["s0", "s1", "zz"])]}}
        style = dict(default_flow_style=rng.choice([False, None, True]), default_style=rng.choice([None, None, None, '"', "'"]), width=rng.choice([30, 80, 1000]),
                     indent=rng.choice([2, 4]), sort_keys=False)
        (d / "pingoo.yml").write_text(yaml.safe_dump(doc, **style))
        if rng.random() < 0.6:
            (d / "rules").mkdir()
            for fn in rng.sample(["a.yml", "b.yml", "c.yaml", "d.txt"], rng.randrange(1, 3)):
                (d / "rules" / fn).write_text(yaml.safe_dump({rng.choice(names): rule() for _ in range(rng.randrange(0, 3))}, **style))
        listener = rng.choice([None, None, "http"])
        res = []
        try:
            cfg = load_config(str(d), geoip_dirs=[str(d)], listener=listener)
            res.append(("ok", Sim(cfg.rules, cfg.lists, cfg.geoip_mmdb, services=cfg.services).describe()))
        except (Error, ValueError) as e:
            res.append(("error", str(e)))
        try:
            res.append(("ok", Sim.from_config_dir(str(d), listener=listener, geoip_dir=str(d)).describe()))
        except (Error, ValueError) as e:
            res.append(("error", str(e)))
        # a malformed list entry is reported by the engine when the list is added: the native loader knows the file's path (as
        # lists.rs does), the Python path only the list's name
        norm = [(kind, re.sub(r"error parsing list \S+ at line", "error parsing list <l> at line", text)) for kind, text in res]
        assert norm[0] == norm[1], (d / "pingoo.yml").read_text()
        agree_ok += res[0][0] == "ok"
        agree_err += res[0][0] == "error"
        shutil.rmtree(d)
    assert agree_ok >= 5 and agree_err >= 5
