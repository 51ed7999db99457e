"""Python mirror of the reference's rule surface.

  rules::Action / Error / compile_expression / validate_expression   rules/rules.rs:30-77
  pingoo::rules::Rule{name, expression, actions}                      pingoo/rules.rs:9-14
  pingoo::lists::ListType                                             pingoo/lists.rs:17-22

Expressions are parsed by the engine's C++ front-end through the C ABI; nothing
here evaluates a rule on the CPU.
"""
import ctypes as C
import enum
from dataclasses import dataclass, field
from typing import List, Optional

from . import _ffi


class Error(Exception):
    """rules::Error (Unspecified / ExpressionIsNotValid)."""


class ExpressionIsNotValid(Error):
    pass


class Action(enum.IntEnum):
    """rules::Action, serde tag `action: block | captcha` (rules/rules.rs:30-35)."""

    BLOCK = _ffi.ACTION_BLOCK
    CAPTCHA = _ffi.ACTION_CAPTCHA

    @classmethod
    def from_config(cls, obj):
        name = obj["action"] if isinstance(obj, dict) else obj
        try:
            return {"block": cls.BLOCK, "captcha": cls.CAPTCHA}[name]
        except KeyError:
            raise Error(f"unknown variant `{name}`, expected `block` or `captcha`")


class ListType(enum.IntEnum):
    String = 0
    Int = 1
    Ip = 2

    @classmethod
    def from_str(cls, value):
        try:
            return {"String": cls.String, "Int": cls.Int, "Ip": cls.Ip}[value]
        except KeyError:
            raise Error(f"{value} is not a valid ListType")


def compile_expression(expression: str) -> str:
    """rules::compile_expression: returns the (validated) source or raises ExpressionIsNotValid."""
    lib = _ffi.load()
    err = C.create_string_buffer(512)
    if lib.pgw_compile_expression(expression.encode(), err, len(err)):
        raise ExpressionIsNotValid(err.value.decode(errors="replace"))
    return expression


def validate_expression(expression: str) -> None:
    """rules::validate_expression: additionally rejects empty input and the `in` operator."""
    lib = _ffi.load()
    err = C.create_string_buffer(512)
    if lib.pgw_validate_expression(expression.encode(), err, len(err)):
        raise ExpressionIsNotValid(err.value.decode(errors="replace"))


@dataclass
class Rule:
    """pingoo::rules::Rule (pingoo/rules.rs:9-14); `expression=None` matches every request."""

    name: str
    expression: Optional[str] = None
    actions: List[Action] = field(default_factory=list)

    @classmethod
    def from_config(cls, name, cfg):
        """RuleConfigFile {expression?, actions} (pingoo/config/config_file.rs:97-101)."""
        return cls(name=name, expression=cfg.get("expression"), actions=[Action.from_config(a) for a in cfg.get("actions", [])])


@dataclass
class Service:
    """The routing view of an HTTP service: `HttpService::match_request` (pingoo/services/mod.rs:33-37,
    http_proxy_service.rs:84-95, http_static_site_service.rs:70-81). `route=None` matches every request."""

    name: str
    route: Optional[str] = None

    @classmethod
    def from_config(cls, name, cfg):
        """ServiceConfigFile {route?, ...} (pingoo/config/config_file.rs:257-265): only the route matters here."""
        return cls(name=name, route=cfg.get("route"))
