"""Columnar request batch = RequestData + ClientData of pingoo/rules.rs:16-34 for n requests.

`RequestBatch` owns numpy columns in the layout `pgw_batch` (include/pingoo_waf.h)
describes; `pack_requests` mirrors what http_listener.rs:139-219 extracts from one
hyper request (host/path trimming rules, user-agent shaping) so tests can start
from "raw" requests.
"""
import ipaddress
from typing import Iterable

import numpy as np

from . import _ffi

FIELDS = _ffi.FIELDS


def _pad16(a: np.ndarray) -> np.ndarray:
    # the kernel reads whole 32-byte chunks: keep the column readable to round_up(len, 32)
    n = (len(a) + 31) // 32 * 32
    if n == 0:
        n = 32
    out = np.zeros(n, dtype=np.uint8)
    out[: len(a)] = a
    return out


class RequestBatch:
    """SoA batch on the host (numpy).  String columns are byte arrays + n+1 uint32 offsets."""

    def __init__(self, n, cols, ip, ip_is_v6, remote_port, asn=None, country=None, flags=None):
        self.n = int(n)
        self.cols = {}
        for f in FIELDS:
            b, o = cols[f]
            b = np.ascontiguousarray(b, dtype=np.uint8)
            o = np.ascontiguousarray(o, dtype=np.uint32)
            assert len(o) == self.n + 1, f"{f}: need n+1 offsets"
            self.total = getattr(self, "total", {})
            self.total[f] = int(o[-1]) if len(o) else 0
            if len(b) % 32 or len(b) < self.total[f] or len(b) == 0:
                b = _pad16(b[: self.total[f]])
            self.cols[f] = (b, o)
        self.ip = np.ascontiguousarray(ip, dtype=np.uint8).reshape(self.n, 16)
        self.ip_is_v6 = np.ascontiguousarray(ip_is_v6, dtype=np.uint8)
        self.remote_port = np.ascontiguousarray(remote_port, dtype=np.int32)
        self.asn = None if asn is None else np.ascontiguousarray(asn, dtype=np.int64)
        self.country = None if country is None else np.ascontiguousarray(country, dtype=np.uint16)
        self.flags = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)

    # ---- views ---------------------------------------------------------------------------------
    def field(self, f, i) -> bytes:
        b, o = self.cols[f]
        return bytes(b[o[i]:o[i + 1]])

    def slice(self, lo, hi) -> "RequestBatch":
        cols = {}
        for f in FIELDS:
            b, o = self.cols[f]
            o2 = (o[lo:hi + 1] - o[lo]).astype(np.uint32)
            cols[f] = (b[o[lo]:o[hi]].copy(), o2)
        return RequestBatch(
            hi - lo, cols, self.ip[lo:hi], self.ip_is_v6[lo:hi], self.remote_port[lo:hi],
            None if self.asn is None else self.asn[lo:hi], None if self.country is None else self.country[lo:hi],
            None if self.flags is None else self.flags[lo:hi])

    def as_ctypes(self) -> _ffi.Batch:
        """pgw_batch over the HOST arrays (keep `self` alive while it is in use)."""
        b = _ffi.Batch()
        b.n = self.n
        for f in FIELDS:
            by, of = self.cols[f]
            sc = getattr(b, f)
            sc.bytes = by.ctypes.data
            sc.offsets = of.ctypes.data
        b.ip = self.ip.ctypes.data
        b.ip_is_v6 = self.ip_is_v6.ctypes.data
        b.remote_port = self.remote_port.ctypes.data
        b.asn = None if self.asn is None else self.asn.ctypes.data
        b.country = None if self.country is None else self.country.ctypes.data
        b.flags = None if self.flags is None else self.flags.ctypes.data
        return b

    def nbytes(self, fields=FIELDS) -> int:
        return sum(self.total[f] for f in fields)


def country_code(cc: str) -> int:
    """CountryCode([u8;2]) (pingoo/geoip.rs:39-40) as the uint16 the batch carries."""
    b = cc.encode()
    return b[0] | (b[1] << 8)


def _ip16(ip):
    a = ipaddress.ip_address(ip)
    if a.version == 4:
        return a.packed + b"\0" * 12, 0
    return a.packed, 1


def pack_requests(requests: Iterable[dict], with_geo: bool = True) -> RequestBatch:
    """Build a batch from dicts with keys host,url,path,method,user_agent,ip,remote_port[,asn,country,flags].

    Applies the shaping the listener applies before rules see a request:
      path: trailing '/' trimmed (services/http_utils.rs:114-116, "/" -> "")
      host: header `to_str` (visible ASCII or tab, else ""), trimmed, longer than 256 bytes -> ""   (http_listener.rs:284-296)
      user_agent: trimmed; non-visible-ASCII or longer than 256 bytes -> "" (http_listener.rs:159-165)
    """
    reqs = list(requests)
    n = len(reqs)
    cols = {}
    shaped = []
    for r in reqs:
        host = r.get("host", "")
        hb = host.encode("utf-8", "surrogateescape")
        if any(not (32 <= c < 127 or c == 9) for c in hb):  # HeaderValue::to_str fails -> unwrap_or_default()
            host = ""
        host = host.strip()
        if len(host.encode()) > 256:
            host = ""
        ua = r.get("user_agent", "")
        try:
            ub = ua.encode("ascii")
            if any((c < 32 and c != 9) or c == 127 for c in ub):
                raise UnicodeError
            ua = ua.strip()
            if len(ua) > 256:
                ua = ""
        except UnicodeError:
            ua = ""
        path = r.get("path", "").rstrip("/")
        shaped.append({"host": host, "url": r.get("url", ""), "path": path, "method": r.get("method", "GET"), "user_agent": ua})
    for f in FIELDS:
        parts = [s[f].encode() for s in shaped]
        offs = np.zeros(n + 1, dtype=np.uint32)
        if n:
            offs[1:] = np.cumsum([len(p) for p in parts], dtype=np.uint64)
        cols[f] = (np.frombuffer(b"".join(parts), dtype=np.uint8), offs)
    ip = np.zeros((n, 16), dtype=np.uint8)
    v6 = np.zeros(n, dtype=np.uint8)
    port = np.zeros(n, dtype=np.int32)
    asn = np.zeros(n, dtype=np.int64)
    cc = np.full(n, country_code("XX"), dtype=np.uint16)
    flags = np.zeros(n, dtype=np.uint8)
    for i, r in enumerate(reqs):
        raw, is6 = _ip16(r.get("ip", "0.0.0.0"))
        ip[i] = np.frombuffer(raw, dtype=np.uint8)
        v6[i] = is6
        port[i] = r.get("remote_port", 0)
        asn[i] = r.get("asn", 0)
        cc[i] = country_code(r.get("country", "XX"))
        flags[i] = r.get("flags", 0)
    return RequestBatch(n, cols, ip, v6, port, asn if with_geo else None, cc if with_geo else None, flags)
