"""WafEngine: the batched replacement for the per-request rule loop.

Mirrors, for a batch, what the reference does per request:
  lists  = load_lists(config.lists)                        pingoo/lists.rs:48-60
  geoip  = GeoipDB::load()                                 pingoo/geoip.rs:44-71
  rules  = config.rules (compile_expression each)          pingoo/config/config.rs:255-269
  for rule in rules: rule.match_request(ctx) -> actions    pingoo/listeners/http_listener.rs:251-264
All evaluation happens in CUDA kernels behind the C ABI (include/pingoo_waf.h).
"""
import ctypes as C
from typing import Dict, Iterable, Optional, Tuple

import numpy as np

from . import _ffi
from .batch import FIELDS, RequestBatch
from .rules import Error, ListType, Rule, Service


class WafEngine:
    def __init__(self, rules: Iterable[Rule], lists: Optional[Dict[str, Tuple[ListType, bytes]]] = None,
                 geoip_mmdb: Optional[bytes] = None, device: int = 0, eval_gates: bool = True,
                 max_dfa_states: int = 0, max_unit_table_bytes: int = 0, services: Optional[Iterable[Service]] = None,
                 candidate_gate: bool = True, literal_confirm: bool = True):
        self._lib = _ffi.load()
        self._h = C.c_void_p()
        self.rules = list(rules)
        descs = (_ffi.RuleDesc * max(1, len(self.rules)))()
        self._keep = []
        for i, r in enumerate(self.rules):
            acts = (C.c_uint8 * max(1, len(r.actions)))(*[int(a) for a in r.actions])
            self._keep.append(acts)
            descs[i].name = r.name.encode()
            descs[i].expression = None if r.expression is None else r.expression.encode()
            descs[i].actions = C.cast(acts, C.POINTER(C.c_uint8))
            descs[i].n_actions = len(r.actions)
        opt = _ffi.Options(max_dfa_states, max_unit_table_bytes, 1 if eval_gates else 0, (0 if candidate_gate else 1) | (0 if literal_confirm else 2))
        err = C.create_string_buffer(1024)
        if self._lib.pgw_ruleset_create(descs, len(self.rules), C.byref(opt), C.byref(self._h), err, len(err)):
            raise Error(err.value.decode(errors="replace"))
        for name, (ltype, csv) in (lists or {}).items():
            if self._lib.pgw_lists_add(self._h, name.encode(), int(ltype), csv, len(csv), err, len(err)):
                msg = err.value.decode(errors="replace")
                self.close()
                raise Error(msg)
        if geoip_mmdb is not None:
            if self._lib.pgw_geoip_load(self._h, geoip_mmdb, len(geoip_mmdb), err, len(err)):
                msg = err.value.decode(errors="replace")
                self.close()
                raise Error(msg)
        self.services = list(services or [])
        if self.services:
            # http_listener.rs:266-270: services are tried in configuration order
            sd = (_ffi.ServiceDesc * len(self.services))()
            for i, sv in enumerate(self.services):
                sd[i].name = sv.name.encode()
                sd[i].route = None if sv.route is None else sv.route.encode()
            if self._lib.pgw_services_set(self._h, sd, len(self.services), err, len(err)):
                msg = err.value.decode(errors="replace")
                self.close()
                raise Error(msg)
        if self._lib.pgw_ruleset_finalize(self._h, device, err, len(err)):
            msg = err.value.decode(errors="replace")
            self.close()
            raise Error(msg)
        self.device = device

    @classmethod
    def from_config_dir(cls, folder: str, listener: Optional[str] = None, geoip_dirs: Optional[Iterable[str]] = None, device: int = 0,
                        eval_gates: bool = True, max_dfa_states: int = 0, max_unit_table_bytes: int = 0, candidate_gate: bool = True,
                        literal_confirm: bool = True):
        """pgw_ruleset_load_dir + finalize: a Pingoo configuration directory consumed by the engine's own (C++) loader."""
        self = cls.__new__(cls)
        self._lib = _ffi.load()
        self._h = C.c_void_p()
        self.rules, self.services, self._keep = [], [], []
        opt = _ffi.Options(max_dfa_states, max_unit_table_bytes, 1 if eval_gates else 0, (0 if candidate_gate else 1) | (0 if literal_confirm else 2))
        err = C.create_string_buffer(2048)
        dirs = [d.encode() for d in (geoip_dirs or [])]
        arr = (C.c_char_p * max(1, len(dirs)))(*dirs)
        if self._lib.pgw_ruleset_load_dir(folder.encode(), None if listener is None else listener.encode(), arr, len(dirs), C.byref(opt),
                                          C.byref(self._h), err, len(err)):
            raise Error(err.value.decode(errors="replace"))
        if self._lib.pgw_ruleset_finalize(self._h, device, err, len(err)):
            msg = err.value.decode(errors="replace")
            self.close()
            raise Error(msg)
        self.device = device
        return self

    # ---- lifecycle ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.pgw_ruleset_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> _ffi.Info:
        i = _ffi.Info()
        self._lib.pgw_ruleset_info(self._h, C.byref(i))
        return i

    def set_profiling(self, enable: bool) -> None:
        """Measurement hook: bracket every scan-kernel launch with CUDA events (see include/pingoo_waf.h)."""
        if self._lib.pgw_ruleset_set_profiling(self._h, 1 if enable else 0) != 0:
            raise Error(self._lib.pgw_last_error().decode(errors="replace"))

    def profile(self) -> tuple[float, int]:
        """(summed scan-kernel milliseconds, launches covered) since profiling was enabled / last read."""
        ms, n = C.c_double(0.0), C.c_uint32(0)
        if self._lib.pgw_ruleset_profile(self._h, C.byref(ms), C.byref(n)) != 0:
            raise Error(self._lib.pgw_last_error().decode(errors="replace"))
        return ms.value, n.value

    def profile_kernels(self) -> tuple[list, int]:
        """([pre-pass ms, scan ms, epilogue ms] summed, batches covered) since profiling was enabled / last read."""
        ms, n = (C.c_double * 3)(), C.c_uint32(0)
        if self._lib.pgw_ruleset_profile_kernels(self._h, ms, C.byref(n)) != 0:
            raise Error(self._lib.pgw_last_error().decode(errors="replace"))
        return [ms[0], ms[1], ms[2]], n.value

    def describe(self) -> str:
        n = self._lib.pgw_ruleset_describe(self._h, None, 0)
        buf = C.create_string_buffer(n + 1)
        self._lib.pgw_ruleset_describe(self._h, buf, n + 1)
        return buf.value.decode(errors="replace")

    # ---- evaluation --------------------------------------------------------------------------
    def evaluate_host(self, batch: RequestBatch) -> np.ndarray:
        """Host columns in, host verdicts out (H2D + kernel + D2H inside the call)."""
        out = np.empty(batch.n, dtype=np.uint32)
        cb = batch.as_ctypes()
        if self._lib.pgw_evaluate_batch_host(self._h, C.byref(cb), out.ctypes.data):
            raise Error(self._lib.pgw_last_error().decode(errors="replace"))
        return out

    def evaluate_host_routed(self, batch: RequestBatch):
        """Verdicts plus, for allowed requests, the index of the first matching service (NO_SERVICE = 404)."""
        out = np.empty(batch.n, dtype=np.uint32)
        svc = np.empty(batch.n, dtype=np.uint16)
        cb = batch.as_ctypes()
        if self._lib.pgw_evaluate_batch_routed_host(self._h, C.byref(cb), out.ctypes.data, svc.ctypes.data):
            raise Error(self._lib.pgw_last_error().decode(errors="replace"))
        return out, svc

    def to_device(self, batch: RequestBatch):
        """Copy a host batch into torch CUDA tensors; returns (tensors, pgw_batch of device pointers)."""
        import torch

        dev = torch.device("cuda", self.device)
        t = {}
        cb = _ffi.Batch()
        cb.n = batch.n
        for f in FIELDS:
            by, of = batch.cols[f]
            t[f + "_b"] = torch.from_numpy(by).to(dev)
            t[f + "_o"] = torch.from_numpy(of.view(np.int32)).to(dev)
            sc = getattr(cb, f)
            sc.bytes = t[f + "_b"].data_ptr()
            sc.offsets = t[f + "_o"].data_ptr()
        for name, arr, view in (("ip", batch.ip, None), ("ip_is_v6", batch.ip_is_v6, None), ("remote_port", batch.remote_port, None),
                                ("asn", batch.asn, None), ("country", batch.country, np.int16), ("flags", batch.flags, None)):
            if arr is None:
                setattr(cb, name, None)
                continue
            a = arr if view is None else arr.view(view)
            t[name] = torch.from_numpy(a).to(dev)
            setattr(cb, name, t[name].data_ptr())
        return t, cb

    def evaluate_device(self, cbatch: _ffi.Batch, verdict_tensor, stream: int = 0):
        """Enqueue the kernel on `stream` (cudaStream_t handle) for device-resident columns."""
        if self._lib.pgw_evaluate_batch(self._h, C.byref(cbatch), verdict_tensor.data_ptr(), stream):
            raise Error(self._lib.pgw_last_error().decode(errors="replace"))

    def evaluate_device_routed(self, cbatch: _ffi.Batch, verdict_tensor, service_tensor, stream: int = 0):
        """Like evaluate_device, also writing the uint16 service index per request (int16 tensor of n elements)."""
        if self._lib.pgw_evaluate_batch_routed(self._h, C.byref(cbatch), verdict_tensor.data_ptr(), service_tensor.data_ptr(), stream):
            raise Error(self._lib.pgw_last_error().decode(errors="replace"))

    def client_ids_device(self, cbatch: _ffi.Batch, out_tensor, stream: int = 0):
        """generate_captcha_client_id for every request of a device-resident batch: out_tensor is uint8 [n, 44]."""
        if self._lib.pgw_captcha_client_id_batch(C.byref(cbatch), out_tensor.data_ptr(), stream):
            raise Error(self._lib.pgw_last_error().decode(errors="replace"))

    def geoip_lookup_device(self, ip_t, v6_t, asn_t, cc_t, stream: int = 0):
        n = ip_t.shape[0]
        if self._lib.pgw_geoip_lookup_batch(self._h, ip_t.data_ptr(), v6_t.data_ptr(), n, asn_t.data_ptr(), cc_t.data_ptr(), stream):
            raise Error(self._lib.pgw_last_error().decode(errors="replace"))


def decode_verdict(v: int):
    """(action, rule_index or None) from a verdict word."""
    rule = int(v) >> 2
    return int(v) & 3, (None if rule == _ffi.NO_RULE else rule)


def make_request(host=b"", url=b"", path=b"", method=b"GET", user_agent=b"", ip="0.0.0.0", remote_port=0, flags=0) -> _ffi.Request:
    """A pgw_request from raw byte strings (what the listener has in hand before any shaping)."""
    import ipaddress

    def b(x):
        return x if isinstance(x, bytes) else str(x).encode("utf-8", "surrogateescape")

    r = _ffi.Request()
    for name, val in (("host", host), ("url", url), ("path", path), ("method", method), ("user_agent", user_agent)):
        raw = b(val)
        setattr(r, name, raw)
        setattr(r, name + "_len", len(raw))
    a = ipaddress.ip_address(ip)
    packed = a.packed + (b"\0" * 12 if a.version == 4 else b"")
    r.ip = (C.c_uint8 * 16)(*packed)
    r.ip_is_v6 = 0 if a.version == 4 else 1
    r.remote_port = int(remote_port)
    r.flags = int(flags)
    return r


def shape_request(req: _ffi.Request):
    """The listener's shaping of one request (no device needed): dict of the five rule-visible strings."""
    lib = _ffi.load()
    ptr = (C.c_char_p * 5)()
    ln = (C.c_size_t * 5)()
    raw = (C.c_void_p * 5)()
    if lib.pgw_shape_request(C.byref(req), C.cast(raw, C.POINTER(C.c_char_p)), ln):
        raise Error("pgw_shape_request failed")
    return {f: C.string_at(raw[i], ln[i]) if ln[i] else b"" for i, f in enumerate(FIELDS)}


class RequestQueue:
    """Micro-batching front of a WafEngine (include/pingoo_waf.h: pgw_queue_*): `evaluate` is thread-safe and blocks until
    the batch that contains the request has been evaluated (ctypes releases the GIL while it waits)."""

    def __init__(self, engine: WafEngine, max_batch: int = 4096, max_delay_us: int = 200):
        self._engine = engine  # keeps the ruleset alive
        self._lib = engine._lib
        self._q = C.c_void_p()
        err = C.create_string_buffer(256)
        if self._lib.pgw_queue_create(engine._h, max_batch, max_delay_us, C.byref(self._q), err, len(err)):
            raise Error(err.value.decode(errors="replace"))

    def evaluate(self, req: _ffi.Request):
        v, s = C.c_uint32(0), C.c_uint16(0)
        rc = self._lib.pgw_queue_evaluate(self._q, C.byref(req), C.byref(v), C.byref(s))
        if rc:
            buf = C.create_string_buffer(512)
            self._lib.pgw_queue_last_error(self._q, buf, len(buf))
            raise Error(f"pgw_queue_evaluate failed ({rc}): " + buf.value.decode(errors="replace"))
        return v.value, s.value

    def stats(self) -> _ffi.QueueStats:
        st = _ffi.QueueStats()
        self._lib.pgw_queue_get_stats(self._q, C.byref(st))
        return st

    def close(self):
        if getattr(self, "_q", None) and self._q.value:
            self._lib.pgw_queue_destroy(self._q)
            self._q = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
