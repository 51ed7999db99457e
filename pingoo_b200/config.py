"""Config / wire formats of the rule path: consume a Pingoo configuration directory unchanged.

Restates, for the inputs of the rule path only, what `config::load_and_validate` and `Server::run` do at start-up:
  /etc/pingoo/pingoo.yml  rules: {name: {expression?, actions: [{action: block|captcha}]}}     config_file.rs:97-101
                          services: {name: {route?, ...}}   (order = YAML order)                 config_file.rs:49-66
                          lists: {name: {type: String|Int|Ip, file: path}}                       config.rs:158-161
  /etc/pingoo/rules/*.yml more rules, appended after the file's own (directory order of the OS)  config.rs:206-213, 378-422
  duplicate rule names are a configuration error ("duplicate rule name: X")                     config.rs:207-212, 411-416
  a rule expression / service route that does not compile is a configuration error              config.rs:255-269, config_file.rs:257-265
  list files are header-less CSV (parsed by the engine: pgw_lists_add)                           lists.rs:62-113
  GeoIP: first existing of geoip.mmdb[.zst] in /etc/pingoo, /usr/share/pingoo; `.zst` = Zstandard  config.rs:31-36, geoip.rs:44-58, 94-109
Listeners, TLS, service back-ends and the rest of the file are ignored here (out of scope).
"""
import ctypes as C
import ctypes.util
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import yaml

from .rules import Error, ListType, Rule, Service, compile_expression

DEFAULT_CONFIG_FOLDER = "/etc/pingoo"
GEOIP_DATABASE_NAMES = ("geoip.mmdb", "geoip.mmdb.zst")


@dataclass
class LoadedConfig:
    rules: List[Rule] = field(default_factory=list)
    services: List[Service] = field(default_factory=list)
    lists: Dict[str, Tuple[ListType, bytes]] = field(default_factory=dict)
    geoip_mmdb: Optional[bytes] = None
    geoip_path: Optional[str] = None

    def engine(self, device: int = 0, **kw):
        from .engine import WafEngine

        return WafEngine(self.rules, self.lists, self.geoip_mmdb, device=device, services=self.services, **kw)


def zstd_decode_all(data: bytes) -> bytes:
    """zstd::decode_all (geoip.rs:52-56) through the system libzstd; streams, so frames without a content size work."""
    name = ctypes.util.find_library("zstd") or "libzstd.so.1"
    try:
        z = C.CDLL(name)
    except OSError as e:
        raise Error(f"libzstd is not available: {e}")
    z.ZSTD_createDStream.restype = C.c_void_p
    z.ZSTD_freeDStream.argtypes = [C.c_void_p]
    z.ZSTD_decompressStream.restype = C.c_size_t
    z.ZSTD_isError.argtypes = [C.c_size_t]
    z.ZSTD_getErrorName.restype = C.c_char_p
    z.ZSTD_getErrorName.argtypes = [C.c_size_t]

    class Buf(C.Structure):
        _fields_ = [("p", C.c_void_p), ("size", C.c_size_t), ("pos", C.c_size_t)]

    z.ZSTD_decompressStream.argtypes = [C.c_void_p, C.POINTER(Buf), C.POINTER(Buf)]
    ds = z.ZSTD_createDStream()
    src = C.create_string_buffer(data, len(data))
    inb = Buf(C.cast(src, C.c_void_p), len(data), 0)
    chunk = C.create_string_buffer(1 << 20)
    out = bytearray()
    try:
        ret = 1
        while inb.pos < inb.size or ret != 0:
            outb = Buf(C.cast(chunk, C.c_void_p), len(chunk), 0)
            ret = z.ZSTD_decompressStream(ds, C.byref(outb), C.byref(inb))
            if z.ZSTD_isError(ret):
                raise Error(z.ZSTD_getErrorName(ret).decode())
            out += chunk.raw[:outb.pos]
            if inb.pos >= inb.size and outb.pos < outb.size:
                if ret != 0:
                    raise Error("incomplete frame")
                break
    finally:
        z.ZSTD_freeDStream(ds)
    return bytes(out)


def _os_error(e: OSError) -> str:
    """std::io::Error's Display: `No such file or directory (os error 2)`."""
    return f"{e.strerror} (os error {e.errno})"


def _rules_from_mapping(mapping, where) -> List[Rule]:
    if mapping is None:
        return []
    if not isinstance(mapping, dict):
        raise Error(f"error parsing {where}: invalid type: expected a map of rules")
    out = []
    for name, cfg in mapping.items():
        if not isinstance(cfg, dict) or "actions" not in cfg:
            raise Error(f"error parsing {where}: {name}: missing field `actions`")
        expr = cfg.get("expression")
        try:
            actions = Rule.from_config(str(name), cfg).actions
        except Error as e:   # serde reports an unknown action while the file is being deserialised: part of the parse error
            raise Error(f"error parsing {where}: {name}: {e}")
        out.append(Rule(name=str(name), expression=None if expr is None else str(expr), actions=actions))
    return out


def load_config(folder: str = DEFAULT_CONFIG_FOLDER, geoip_dirs: Optional[List[str]] = None, listener: Optional[str] = None) -> LoadedConfig:
    """The rule-path inputs of a configuration directory.  `services` = what an HTTP listener offers a request to
    (config.rs:217-244, server.rs:62-74, 104-110): `listener` None -> every service with `http_proxy` or `static`, in
    configuration order (tcp_proxy services never carry a route and are not offered to HTTP requests); else the
    `services:` list of that listener, in its own order.  The product-side loader with the same behaviour is the C ABI's
    pgw_ruleset_load_dir (csrc/config_dir.cpp); this Python one feeds the oracle and the tests."""
    cfg_path = os.path.join(folder, "pingoo.yml")
    try:
        raw = open(cfg_path, "rb").read()
    except OSError as e:
        raise Error(f"error reading config file ({cfg_path}): {_os_error(e)}")
    try:
        doc = yaml.safe_load(raw) or {}
    except yaml.YAMLError as e:
        raise Error(f"error parsing config file ({cfg_path}): {e}")
    out = LoadedConfig()
    out.rules = _rules_from_mapping(doc.get("rules"), f"config file ({cfg_path})")

    # rules folder: *.yml only, in the directory order the OS returns (config.rs:378-422)
    rules_dir = os.path.join(folder, "rules")
    folder_rules: List[Rule] = []
    if os.path.isdir(rules_dir):
        with os.scandir(rules_dir) as it:
            for ent in it:
                if not ent.name.endswith(".yml") or os.path.splitext(ent.name)[1] != ".yml":
                    continue
                try:
                    content = open(ent.path, "rb").read()
                except OSError as e:
                    raise Error(f'error reading rules file "{ent.path}": {_os_error(e)}')
                try:
                    mapping = yaml.safe_load(content)
                except yaml.YAMLError as e:
                    raise Error(f'error parsing rules file "{ent.path}": {e}')
                new = _rules_from_mapping(mapping, f'rules file "{ent.path}"')
                seen = {r.name for r in folder_rules}
                for r in new:
                    if r.name in seen:
                        raise Error(f"duplicate rule name: {r.name}")
                folder_rules += new
    names = {r.name for r in out.rules}
    for r in folder_rules:
        if r.name in names:
            raise Error(f"duplicate rule name: {r.name}")
    out.rules += folder_rules

    all_services = {}
    for name, cfg in (doc.get("services") or {}).items():
        cfg = cfg or {}
        kinds = [k for k in ("http_proxy", "static", "tcp_proxy") if cfg.get(k) is not None]
        if len(kinds) != 1:
            raise Error(f"invalid service definition for {name}: services must have exactly 1 http_proxy, tcp_proxy or static field")
        sv = Service.from_config(str(name), cfg)
        if kinds[0] == "tcp_proxy" and sv.route is not None:
            raise Error(f"Invalid service definition for {name}: TCP proxy can't have a route")
        if sv.route is not None:
            try:
                compile_expression(sv.route)
            except Error as e:
                raise Error(f"error parsing route for service {sv.name}: {e}")
        all_services[str(name)] = (sv, kinds[0] != "tcp_proxy")
    offered = None
    if listener is not None:
        lcfg = (doc.get("listeners") or {}).get(listener)
        if lcfg is None:
            raise Error(f"config: listeners: {listener}: no such listener")
        if lcfg.get("services") is not None:
            offered = []
            for nm in lcfg["services"]:
                if nm not in all_services:
                    raise Error(f"config: listeners: {listener}: service {nm} doesn't exist")
                if nm in [o.name for o in offered]:
                    raise Error(f"config: listeners: {listener}: duplicate services are not allowed ({nm})")
                if not all_services[nm][1]:
                    raise Error(f"config: listeners: {listener}: service {nm} is not an HTTP service")
                offered.append(all_services[nm][0])
    out.services = offered if offered is not None else [sv for sv, is_http in all_services.values() if is_http]
    # the rules are compiled after the services were parsed (routes included) and the listeners validated (config.rs:217-269)
    for r in out.rules:  # config.rs:255-269: compile errors are fatal
        if r.expression is not None:
            try:
                compile_expression(r.expression)
            except Error as e:
                raise Error(f"error parsing rules: {e}")

    for name, lc in (doc.get("lists") or {}).items():
        try:
            ltype = ListType[str(lc["type"])]
        except KeyError:
            raise Error(f"error parsing config file ({cfg_path}): lists.{name}: unknown variant `{lc.get('type')}`, expected one of `String`, `Int`, `Ip`")
        path = lc["file"]
        try:
            out.lists[str(name)] = (ltype, open(path, "rb").read())
        except OSError as e:
            raise Error(f"error reading list {path}: {_os_error(e)}")   # lists.rs:62-66

    for d in (geoip_dirs if geoip_dirs is not None else [folder, "/usr/share/pingoo"]):
        for nm in GEOIP_DATABASE_NAMES:
            p = os.path.join(d, nm)
            if os.path.exists(p):
                try:
                    data = open(p, "rb").read()
                except OSError as e:
                    raise Error(f"error reading geoip database ({p}): {_os_error(e)}")
                if p.endswith(".zst"):
                    try:
                        data = zstd_decode_all(data)
                    except Error as e:
                        raise Error(f"error decompressing geoip database ({p}): {e}")
                out.geoip_mmdb, out.geoip_path = data, p
                return out
    return out
