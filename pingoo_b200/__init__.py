"""pingoo_b200: B200-native batched WAF verdict engine for Pingoo's rules/lists/GeoIP hot path."""
from . import _ffi  # noqa: F401
from .batch import RequestBatch, country_code, pack_requests  # noqa: F401
from .engine import RequestQueue, WafEngine, decode_verdict, make_request, shape_request  # noqa: F401
from .rules import Action, Error, ExpressionIsNotValid, ListType, Rule, Service, compile_expression, validate_expression  # noqa: F401
