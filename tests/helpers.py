"""Test-side loaders: the oracle (oracle/liboracle.so) and the test-only program
simulator (tests/sim/libpgw_sim.so).  Neither is importable from the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

from pingoo_b200 import _ffi
from pingoo_b200.batch import RequestBatch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
SIM_SO = os.environ.get("PGW_SIM_SO") or os.path.join(ROOT, "tests", "sim", "libpgw_sim.so")   # PGW_SIM_SO: e.g. a sanitizer build (tools/README.md)


def _ensure(path, make_dir):
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", make_dir], stdout=subprocess.DEVNULL)
    return path


def _descs(rules):
    rules = list(rules)
    descs = (_ffi.RuleDesc * max(1, len(rules)))()
    keep = []
    for i, r in enumerate(rules):
        acts = (C.c_uint8 * max(1, len(r.actions)))(*[int(a) for a in r.actions])
        keep.append(acts)
        descs[i].name = r.name.encode()
        descs[i].expression = None if r.expression is None else r.expression.encode()
        descs[i].actions = C.cast(acts, C.POINTER(C.c_uint8))
        descs[i].n_actions = len(r.actions)
    return descs, keep, len(rules)


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(_ensure(ORACLE_SO, os.path.join(ROOT, "oracle")))
        p = C.c_void_p
        lib.orc_create.restype = p
        lib.orc_create.argtypes = [C.POINTER(_ffi.RuleDesc), C.c_uint32, C.c_int, C.c_char_p, C.c_size_t]
        lib.orc_lists_add.argtypes = [p, C.c_char_p, C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        lib.orc_geoip_load.argtypes = [p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        lib.orc_evaluate.argtypes = [p, C.POINTER(_ffi.Batch), p, C.c_int]
        lib.orc_services_set.argtypes = [p, C.POINTER(_ffi.ServiceDesc), C.c_uint32, C.c_char_p, C.c_size_t]
        lib.orc_evaluate_routed.argtypes = [p, C.POINTER(_ffi.Batch), p, p, C.c_int]
        lib.orc_geoip_lookup.argtypes = [p, C.c_char_p, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint16)]
        lib.orc_geoip_lookup.restype = None
        lib.orc_destroy.argtypes = [p]
        lib.orc_destroy.restype = None
        lib.orc_compile_expression.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        lib.orc_validate_expression.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        lib.orc_regex_is_match.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        lib.orc_ipnet_contains.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        lib.orc_eval_kind.argtypes = [C.c_char_p, C.POINTER(_ffi.Batch)]
        _oracle = lib
    return _oracle


def _svc_descs(services):
    services = list(services or [])
    sd = (_ffi.ServiceDesc * max(1, len(services)))()
    for i, sv in enumerate(services):
        sd[i].name = sv.name.encode()
        sd[i].route = None if sv.route is None else sv.route.encode()
    return sd, len(services)


class Oracle:
    """CPU restatement of the reference path (the checker)."""

    def __init__(self, rules, lists=None, geoip_mmdb=None, eval_gates=True, services=None):
        self.lib = oracle_lib()
        descs, keep, n = _descs(rules)
        err = C.create_string_buffer(1024)
        self.h = self.lib.orc_create(descs, n, 1 if eval_gates else 0, err, len(err))
        if not self.h:
            raise ValueError(err.value.decode(errors="replace"))
        for name, (ltype, csv) in (lists or {}).items():
            if self.lib.orc_lists_add(self.h, name.encode(), int(ltype), csv, len(csv), err, len(err)):
                raise ValueError(err.value.decode(errors="replace"))
        if geoip_mmdb is not None:
            if self.lib.orc_geoip_load(self.h, geoip_mmdb, len(geoip_mmdb), err, len(err)):
                raise ValueError(err.value.decode(errors="replace"))
        if services:
            sd, ns = _svc_descs(services)
            if self.lib.orc_services_set(self.h, sd, ns, err, len(err)):
                raise ValueError(err.value.decode(errors="replace"))

    def evaluate(self, batch: RequestBatch, threads=1) -> np.ndarray:
        out = np.empty(batch.n, dtype=np.uint32)
        cb = batch.as_ctypes()
        self.lib.orc_evaluate(self.h, C.byref(cb), out.ctypes.data, threads)
        return out

    def evaluate_routed(self, batch: RequestBatch, threads=1):
        out = np.empty(batch.n, dtype=np.uint32)
        svc = np.empty(batch.n, dtype=np.uint16)
        cb = batch.as_ctypes()
        self.lib.orc_evaluate_routed(self.h, C.byref(cb), out.ctypes.data, svc.ctypes.data, threads)
        return out, svc

    def geoip_lookup(self, ip16: bytes, is_v6: int):
        a = C.c_uint32()
        c = C.c_uint16()
        self.lib.orc_geoip_lookup(self.h, ip16, is_v6, C.byref(a), C.byref(c))
        return a.value, bytes([c.value & 0xFF, c.value >> 8]).decode()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None


_sim = None


def sim_lib():
    global _sim
    if _sim is None:
        csrc = os.path.join(ROOT, "pingoo_b200", "csrc")
        if not os.path.exists(os.path.join(csrc, "build", "ruleset.o")):
            subprocess.check_call(["make", "-C", csrc, "host"], stdout=subprocess.DEVNULL)
        lib = C.CDLL(_ensure(SIM_SO, os.path.join(ROOT, "tests", "sim")))
        p = C.c_void_p
        lib.pgwsim_create.restype = p
        lib.pgwsim_create.argtypes = [C.POINTER(_ffi.RuleDesc), C.c_uint32, C.POINTER(_ffi.Options), C.c_char_p, C.c_size_t]
        lib.pgwsim_lists_add.argtypes = [p, C.c_char_p, C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        lib.pgwsim_geoip_load.argtypes = [p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        lib.pgwsim_finalize.argtypes = [p, C.c_char_p, C.c_size_t]
        lib.pgwsim_describe.restype = C.c_size_t
        lib.pgwsim_describe.argtypes = [p, C.c_char_p, C.c_size_t]
        lib.pgwsim_evaluate.argtypes = [p, C.POINTER(_ffi.Batch), p]
        lib.pgwsim_services_set.argtypes = [p, C.POINTER(_ffi.ServiceDesc), C.c_uint32, C.c_char_p, C.c_size_t]
        lib.pgwsim_evaluate_routed.argtypes = [p, C.POINTER(_ffi.Batch), p, p]
        lib.pgwsim_geoip_lookup.argtypes = [p, p, p, C.c_uint32, p, p]
        lib.pgwsim_destroy.argtypes = [p]
        lib.pgwsim_destroy.restype = None
        lib.pgwsim_load_dir.restype = p
        lib.pgwsim_load_dir.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]
        lib.pgwsim_yaml_dump.restype = C.c_size_t
        lib.pgwsim_yaml_dump.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]
        lib.pgwsim_evaluate_mt.argtypes = [p, C.POINTER(_ffi.Batch), p, C.c_int]
        lib.pgwsim_set_gate_stats.argtypes = [p, p]
        lib.pgwsim_set_gate_stats.restype = None
        _sim = lib
    return _sim


class Sim:
    """CPU walk over the tables the product compiler emits (compiler check without a GPU)."""

    def __init__(self, rules, lists=None, geoip_mmdb=None, eval_gates=True, max_dfa_states=0, max_unit_table_bytes=0, services=None,
                 candidate_gate=True, literal_confirm=True):
        self.lib = sim_lib()
        descs, keep, n = _descs(rules)
        err = C.create_string_buffer(2048)
        opt = _ffi.Options(max_dfa_states, max_unit_table_bytes, 1 if eval_gates else 0, (0 if candidate_gate else 1) | (0 if literal_confirm else 2))
        self.h = self.lib.pgwsim_create(descs, n, C.byref(opt), err, len(err))
        if not self.h:
            raise ValueError(err.value.decode(errors="replace"))
        for name, (ltype, csv) in (lists or {}).items():
            if self.lib.pgwsim_lists_add(self.h, name.encode(), int(ltype), csv, len(csv), err, len(err)):
                raise ValueError(err.value.decode(errors="replace"))
        if geoip_mmdb is not None:
            if self.lib.pgwsim_geoip_load(self.h, geoip_mmdb, len(geoip_mmdb), err, len(err)):
                raise ValueError(err.value.decode(errors="replace"))
        if services:
            sd, ns = _svc_descs(services)
            if self.lib.pgwsim_services_set(self.h, sd, ns, err, len(err)):
                raise ValueError(err.value.decode(errors="replace"))
        if self.lib.pgwsim_finalize(self.h, err, len(err)):
            raise ValueError(err.value.decode(errors="replace"))

    @classmethod
    def from_config_dir(cls, folder, listener=None, geoip_dir=None):
        """The engine's C++ configuration-directory loader (csrc/config_dir.cpp) feeding the table walk."""
        self = cls.__new__(cls)
        self.lib = sim_lib()
        err = C.create_string_buffer(2048)
        self.h = self.lib.pgwsim_load_dir(folder.encode(), None if listener is None else listener.encode(),
                                          None if geoip_dir is None else geoip_dir.encode(), err, len(err))
        if not self.h:
            raise ValueError(err.value.decode(errors="replace"))
        if self.lib.pgwsim_finalize(self.h, err, len(err)):
            raise ValueError(err.value.decode(errors="replace"))
        return self

    def gate_stats(self):
        """Start counting, per field, how many requests the gate saw and how many it made candidates."""
        self._stats = np.zeros(10, dtype=np.uint64)
        self.lib.pgwsim_set_gate_stats(self.h, self._stats.ctypes.data)
        return self._stats

    def describe(self):
        n = self.lib.pgwsim_describe(self.h, None, 0)
        buf = C.create_string_buffer(n + 1)
        self.lib.pgwsim_describe(self.h, buf, n + 1)
        return buf.value.decode(errors="replace")

    def evaluate(self, batch: RequestBatch) -> np.ndarray:
        out = np.empty(batch.n, dtype=np.uint32)
        cb = batch.as_ctypes()
        if self.lib.pgwsim_evaluate(self.h, C.byref(cb), out.ctypes.data):
            raise RuntimeError("sim evaluate failed")
        return out

    def evaluate_mt(self, batch: RequestBatch, threads: int) -> np.ndarray:
        out = np.empty(batch.n, dtype=np.uint32)
        cb = batch.as_ctypes()
        if self.lib.pgwsim_evaluate_mt(self.h, C.byref(cb), out.ctypes.data, threads):
            raise RuntimeError("sim evaluate failed")
        return out

    def evaluate_routed(self, batch: RequestBatch):
        out = np.empty(batch.n, dtype=np.uint32)
        svc = np.empty(batch.n, dtype=np.uint16)
        cb = batch.as_ctypes()
        if self.lib.pgwsim_evaluate_routed(self.h, C.byref(cb), out.ctypes.data, svc.ctypes.data):
            raise RuntimeError("sim evaluate failed")
        return out, svc

    def geoip_lookup(self, ip: np.ndarray, v6: np.ndarray):
        n = len(v6)
        asn = np.zeros(n, dtype=np.uint32)
        cc = np.zeros(n, dtype=np.uint16)
        self.lib.pgwsim_geoip_lookup(self.h, ip.ctypes.data, v6.ctypes.data, n, asn.ctypes.data, cc.ctypes.data)
        return asn, cc

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.pgwsim_destroy(self.h)
            self.h = None


def fmt_verdict(v):
    act = ["allow", "block", "captcha", "bypass"][int(v) & 3]
    r = int(v) >> 2
    return f"{act}@{'-' if r == _ffi.NO_RULE else r}"


def yaml_dump(text: str):
    """(ok, canonical dump or error text) from the engine's YAML subset reader (csrc/yaml.cpp)."""
    lib = sim_lib()
    raw = text.encode()
    ok = C.c_int(0)
    n = lib.pgwsim_yaml_dump(raw, len(raw), None, 0, C.byref(ok))
    buf = C.create_string_buffer(n + 1)
    lib.pgwsim_yaml_dump(raw, len(raw), buf, n + 1, C.byref(ok))
    return bool(ok.value), buf.value.decode(errors="replace")
