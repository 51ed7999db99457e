"""ctypes mirror of include/pingoo_waf.h and the loader for the CUDA library.

The product library is `pingoo_b200/libpingoo_waf.so` (built in-tree by
`__graft_entry__.build()` / `make -C pingoo_b200/csrc`).  There is no CPU
fallback: if the library is missing, importing an engine fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PGW_LIB", os.path.join(_HERE, "libpingoo_waf.so"))  # PGW_LIB: tuning experiments only


class RuleDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("expression", C.c_char_p), ("actions", C.POINTER(C.c_uint8)), ("n_actions", C.c_uint32)]


class ServiceDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("route", C.c_char_p)]


class Request(C.Structure):
    _fields_ = [
        ("host", C.c_char_p), ("host_len", C.c_size_t), ("url", C.c_char_p), ("url_len", C.c_size_t),
        ("path", C.c_char_p), ("path_len", C.c_size_t), ("method", C.c_char_p), ("method_len", C.c_size_t),
        ("user_agent", C.c_char_p), ("user_agent_len", C.c_size_t),
        ("ip", C.c_uint8 * 16), ("ip_is_v6", C.c_uint8), ("flags", C.c_uint8), ("remote_port", C.c_int32),
    ]


class QueueStats(C.Structure):
    _fields_ = [("batches", C.c_uint64), ("requests", C.c_uint64), ("full_flushes", C.c_uint64), ("deadline_flushes", C.c_uint64),
                ("largest_batch", C.c_uint32), ("reserved", C.c_uint32)]


DONE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_uint16, C.c_int)


class Options(C.Structure):
    _fields_ = [("max_dfa_states", C.c_int32), ("max_unit_table_bytes", C.c_uint64), ("eval_gates", C.c_int32),
                ("disable_candidate_gate", C.c_int32)]


class StrCol(C.Structure):
    _fields_ = [("bytes", C.c_void_p), ("offsets", C.c_void_p)]


class Batch(C.Structure):
    _fields_ = [
        ("n", C.c_uint32),
        ("host", StrCol), ("url", StrCol), ("path", StrCol), ("method", StrCol), ("user_agent", StrCol),
        ("ip", C.c_void_p), ("ip_is_v6", C.c_void_p), ("remote_port", C.c_void_p),
        ("asn", C.c_void_p), ("country", C.c_void_p), ("flags", C.c_void_p),
    ]


class Info(C.Structure):
    _fields_ = [
        ("n_rules", C.c_uint32), ("n_atoms", C.c_uint32), ("n_scan_units", C.c_uint32), ("n_nonscan_atoms", C.c_uint32),
        ("scanned_fields_mask", C.c_uint32), ("offset_fields_mask", C.c_uint32),
        ("reads_ip", C.c_uint32), ("reads_port", C.c_uint32), ("reads_geo_columns", C.c_uint32),
        ("table_arena_bytes", C.c_uint64), ("smem_bytes", C.c_uint64),
        ("tables_in_smem", C.c_uint32), ("hot_dfa_states", C.c_uint32), ("grid", C.c_uint32), ("threads", C.c_uint32),
        ("total_dfa_states", C.c_uint32), ("lpm_present", C.c_uint32), ("geoip_loaded", C.c_uint32),
        ("kernel_launches", C.c_uint64), ("last_h2d_bytes", C.c_uint64), ("last_d2h_bytes", C.c_uint64),
        ("gated_fields_mask", C.c_uint32), ("gate_grams", C.c_uint32), ("gate_smem_bytes", C.c_uint64),
        ("n_bitset_units", C.c_uint32), ("bitset_positions", C.c_uint32),
    ]


FIELDS = ("host", "url", "path", "method", "user_agent")
ACTION_BLOCK, ACTION_CAPTCHA = 1, 2
ALLOW, BLOCK, CAPTCHA, BYPASS = 0, 1, 2, 3
NO_RULE = 0x3FFFFFFF
NO_SERVICE = 0xFFFF
FLAG_CAPTCHA_VERIFIED, FLAG_PRE_BLOCK, FLAG_PRE_CAPTCHA, FLAG_BYPASS = 1, 2, 4, 8

EXPORTS = (
    "pgw_compile_expression", "pgw_validate_expression", "pgw_ruleset_create", "pgw_ruleset_load_dir", "pgw_lists_add", "pgw_geoip_load",
    "pgw_ruleset_finalize", "pgw_evaluate_batch", "pgw_evaluate_batch_host", "pgw_geoip_lookup_batch",
    "pgw_services_set", "pgw_evaluate_batch_routed", "pgw_evaluate_batch_routed_host",
    "pgw_captcha_client_id_batch", "pgw_queue_create", "pgw_queue_evaluate", "pgw_queue_submit", "pgw_queue_get_stats", "pgw_queue_last_error", "pgw_queue_destroy", "pgw_shape_request",
    "pgw_host_alloc", "pgw_host_free", "pgw_ruleset_info", "pgw_ruleset_set_profiling", "pgw_ruleset_profile", "pgw_ruleset_profile_kernels",
    "pgw_ruleset_describe", "pgw_ruleset_destroy", "pgw_last_error",
)

_lib = None


def declare(lib, prefix="pgw_"):
    """Attach argtypes/restypes for the pgw_* ABI to `lib`."""
    p = C.c_void_p
    sig = {
        "compile_expression": (C.c_int, [C.c_char_p, C.c_char_p, C.c_size_t]),
        "validate_expression": (C.c_int, [C.c_char_p, C.c_char_p, C.c_size_t]),
        "ruleset_create": (C.c_int, [C.POINTER(RuleDesc), C.c_uint32, C.POINTER(Options), C.POINTER(p), C.c_char_p, C.c_size_t]),
        "lists_add": (C.c_int, [p, C.c_char_p, C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]),
        "geoip_load": (C.c_int, [p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]),
        "ruleset_finalize": (C.c_int, [p, C.c_int, C.c_char_p, C.c_size_t]),
        "evaluate_batch": (C.c_int, [p, C.POINTER(Batch), p, p]),
        "evaluate_batch_host": (C.c_int, [p, C.POINTER(Batch), p]),
        "services_set": (C.c_int, [p, C.POINTER(ServiceDesc), C.c_uint32, C.c_char_p, C.c_size_t]),
        "evaluate_batch_routed": (C.c_int, [p, C.POINTER(Batch), p, p, p]),
        "evaluate_batch_routed_host": (C.c_int, [p, C.POINTER(Batch), p, p]),
        "captcha_client_id_batch": (C.c_int, [C.POINTER(Batch), p, p]),
        "queue_create": (C.c_int, [p, C.c_uint32, C.c_uint32, C.POINTER(p), C.c_char_p, C.c_size_t]),
        "queue_evaluate": (C.c_int, [p, C.POINTER(Request), C.POINTER(C.c_uint32), C.POINTER(C.c_uint16)]),
        "queue_submit": (C.c_int, [p, C.POINTER(Request), DONE_FN, p]),
        "queue_get_stats": (C.c_int, [p, C.POINTER(QueueStats)]),
        "queue_last_error": (C.c_size_t, [p, C.c_char_p, C.c_size_t]),
        "queue_destroy": (None, [p]),
        "shape_request": (C.c_int, [C.POINTER(Request), C.POINTER(C.c_char_p), C.POINTER(C.c_size_t)]),
        "geoip_lookup_batch": (C.c_int, [p, p, p, C.c_uint32, p, p, p]),
        "host_alloc": (p, [C.c_size_t]),
        "host_free": (None, [p]),
        "ruleset_load_dir": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(Options), C.POINTER(p), C.c_char_p, C.c_size_t]),
        "ruleset_info": (C.c_int, [p, C.POINTER(Info)]),
        "ruleset_set_profiling": (C.c_int, [p, C.c_int]),
        "ruleset_profile": (C.c_int, [p, C.POINTER(C.c_double), C.POINTER(C.c_uint32)]),
        "ruleset_profile_kernels": (C.c_int, [p, C.POINTER(C.c_double), C.POINTER(C.c_uint32)]),
        "ruleset_describe": (C.c_size_t, [p, C.c_char_p, C.c_size_t]),
        "ruleset_destroy": (None, [p]),
        "last_error": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, prefix + name)
        fn.restype = res
        fn.argtypes = args
    return lib


def load():
    """Load the CUDA library or raise: this package has no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build the sm_100a extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C pingoo_b200/csrc). "
            "pingoo_b200 has no CPU fallback."
        )
    _lib = declare(C.CDLL(LIB_PATH))
    return _lib
