"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/pingoo_waf.h declares,
parses expressions like rules::compile_expression, and refuses to evaluate without a CUDA device."""
import ctypes as C
import os
import re

import pytest

from pingoo_b200 import Action, Error, ExpressionIsNotValid, Rule, WafEngine, _ffi, compile_expression, validate_expression

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    lib = _ffi.load()
    header = open(os.path.join(ROOT, "include", "pingoo_waf.h")).read()
    declared = set(re.findall(r"\b(pgw_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_ffi.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_struct_layout_matches_header():
    # sizes the C compiler gives the public structs (LP64)
    assert C.sizeof(_ffi.RuleDesc) == 32
    assert C.sizeof(_ffi.StrCol) == 16
    assert C.sizeof(_ffi.Batch) == 8 + 5 * 16 + 6 * 8
    assert C.sizeof(_ffi.Options) == 24  # 4 (+4 pad) + 8 + 4 + 4
    assert C.sizeof(_ffi.ServiceDesc) == 16
    assert C.sizeof(_ffi.Info) == 9 * 4 + 4 + 2 * 8 + 4 * 4 + 3 * 4 + 4 + 3 * 8 + 2 * 4 + 8 + 2 * 4


def test_compile_and_validate_expression_need_no_gpu():
    assert compile_expression('http_request.path.starts_with("/.env")')
    with pytest.raises(ExpressionIsNotValid, match="Expression is not valid"):
        compile_expression('http_request.path ==')
    with pytest.raises(ExpressionIsNotValid, match="expression is empty"):
        validate_expression("")
    compile_expression("1 in [1]")
    with pytest.raises(ExpressionIsNotValid, match="unknown operator: in"):
        validate_expression("1 in [1]")


def test_rule_with_bad_expression_is_a_config_error():
    # config.rs:255-269: a rule that does not compile aborts start-up with "error parsing rules: ..."
    with pytest.raises(Error, match="error parsing rules: Expression is not valid"):
        WafEngine([Rule("r", "http_request.path == (", [Action.BLOCK])], device=0)


@pytest.mark.skipif(_has_cuda(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_device():
    with pytest.raises(Error, match="no CUDA device available.*no CPU fallback"):
        WafEngine([Rule("r", 'http_request.path == "/x"', [Action.BLOCK])], device=0)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_ffi, "_lib", None)
    monkeypatch.setattr(_ffi, "LIB_PATH", "/nonexistent/libpingoo_waf.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _ffi.load()


def test_null_and_invalid_arguments_are_errors_not_crashes():
    """Every entry point that works without a device, called with null pointers / absurd values: an error code and a message,
    never a crash (a proxy must survive a bad configuration reload)."""
    L = _ffi.load()
    err = C.create_string_buffer(256)
    p = C.c_void_p
    sigs = {
        "pgw_compile_expression": [C.c_char_p, C.c_char_p, C.c_size_t], "pgw_validate_expression": [C.c_char_p, C.c_char_p, C.c_size_t],
        "pgw_ruleset_create": [p, C.c_uint32, p, p, C.c_char_p, C.c_size_t], "pgw_lists_add": [p, C.c_char_p, C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t],
        "pgw_geoip_load": [p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t], "pgw_ruleset_finalize": [p, C.c_int, C.c_char_p, C.c_size_t],
        "pgw_ruleset_info": [p, p], "pgw_evaluate_batch": [p, p, p, p], "pgw_evaluate_batch_host": [p, p, p], "pgw_services_set": [p, p, C.c_uint32, C.c_char_p, C.c_size_t],
        "pgw_queue_create": [p, C.c_uint32, C.c_uint32, p, C.c_char_p, C.c_size_t], "pgw_ruleset_load_dir": [C.c_char_p, C.c_char_p, p, C.c_uint32, p, p, C.c_char_p, C.c_size_t],
        "pgw_shape_request": [p, p, p], "pgw_geoip_lookup_batch": [p, p, p, C.c_uint32, p, p, p], "pgw_captcha_client_id_batch": [p, p, p],
        "pgw_ruleset_destroy": [p], "pgw_queue_destroy": [p], "pgw_host_free": [p],
    }
    for name, at in sigs.items():
        getattr(L, name).argtypes = at
    L.pgw_ruleset_describe.argtypes, L.pgw_ruleset_describe.restype = [p, C.c_char_p, C.c_size_t], C.c_size_t
    assert L.pgw_compile_expression(None, err, 256) != 0 and L.pgw_validate_expression(None, None, 0) != 0
    assert L.pgw_compile_expression(b"true", None, 0) == 0
    h = C.c_void_p()
    assert L.pgw_ruleset_create(None, 3, None, C.byref(h), err, 256) != 0 and b"rules is null" in err.value
    assert L.pgw_ruleset_create(None, 0, None, None, err, 256) != 0
    bad = (_ffi.RuleDesc * 1)()
    bad[0].name, bad[0].expression, bad[0].actions, bad[0].n_actions = None, b"true", None, 2
    assert L.pgw_ruleset_create(bad, 1, None, C.byref(h), err, 256) != 0 and b"actions is null" in err.value
    assert L.pgw_ruleset_create(None, 0, None, C.byref(h), err, 256) == 0 and h.value
    assert L.pgw_lists_add(h, None, 0, b"x", 1, err, 256) != 0 and L.pgw_lists_add(None, b"l", 0, b"", 0, err, 256) != 0
    assert L.pgw_lists_add(h, b"l", 7, b"x", 1, err, 256) != 0 and b"not a valid ListType" in err.value
    assert L.pgw_lists_add(h, b"l", 0, None, 0, err, 256) == 0   # an empty list file
    assert L.pgw_geoip_load(h, None, 0, err, 256) != 0 and L.pgw_geoip_load(None, b"x", 1, err, 256) != 0
    assert L.pgw_services_set(h, None, 2, err, 256) != 0 and L.pgw_services_set(None, None, 0, err, 256) != 0
    assert L.pgw_ruleset_finalize(None, 0, err, 256) != 0
    assert L.pgw_ruleset_info(None, None) != 0 and L.pgw_ruleset_info(h, None) != 0
    assert L.pgw_ruleset_describe(None, None, 0) == 0 and L.pgw_ruleset_describe(h, None, 0) > 0
    b = _ffi.Batch()
    assert L.pgw_evaluate_batch(None, None, None, None) != 0 and L.pgw_evaluate_batch(h, C.byref(b), None, None) != 0   # not finalized
    assert L.pgw_evaluate_batch_host(h, C.byref(b), None) != 0
    q = C.c_void_p()
    assert L.pgw_queue_create(None, 16, 100, C.byref(q), err, 256) != 0 and L.pgw_queue_create(h, 0, 100, C.byref(q), err, 256) != 0
    assert L.pgw_ruleset_load_dir(None, None, None, 0, None, C.byref(q), err, 256) != 0
    assert L.pgw_ruleset_load_dir(b"/nonexistent-dir", None, None, 0, None, C.byref(q), err, 256) != 0 and b"error reading config file" in err.value
    assert L.pgw_shape_request(None, None, None) != 0
    assert L.pgw_geoip_lookup_batch(None, None, None, 0, None, None, None) != 0 and L.pgw_captcha_client_id_batch(None, None, None) != 0
    L.pgw_ruleset_destroy(None)
    L.pgw_queue_destroy(None)
    L.pgw_host_free(None)
    L.pgw_ruleset_destroy(h)
