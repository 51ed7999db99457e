"""Config / wire formats of the rule path: a Pingoo configuration directory is consumed unchanged."""
import ctypes as C
import ctypes.util
import os

import numpy as np
import pytest

import synth
from helpers import Oracle, Sim
from pingoo_b200 import Action, Error, ListType
from pingoo_b200.config import load_config, zstd_decode_all

PINGOO_YML = """
listeners:
  http:
    address: http://0.0.0.0:8080
services:
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


def test_zstd_round_trip():
    data = bytes(range(256)) * 4000 + b"tail"
    assert zstd_decode_all(_zstd_compress(data)) == data
