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
