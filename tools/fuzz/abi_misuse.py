import sys, ctypes as C
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0,R)
from pingoo_b200 import _ffi
L=_ffi.load()
err=C.create_string_buffer(256)
def call(name, *a):
    f=getattr(L,name)
    try:
        r=f(*a); print(name, "->", r, err.value[:80])
    except Exception as e:
        print(name, "EXC", e)
L.pgw_compile_expression.argtypes=[C.c_char_p,C.c_char_p,C.c_size_t]
call("pgw_compile_expression", None, err, 256)
call("pgw_compile_expression", b"true", None, 0)
call("pgw_validate_expression", None, None, 0)
h=C.c_void_p()
L.pgw_ruleset_create.argtypes=[C.c_void_p,C.c_uint32,C.c_void_p,C.c_void_p,C.c_char_p,C.c_size_t]
call("pgw_ruleset_create", None, 0, None, C.byref(h), err, 256); print(" handle", h.value)
call("pgw_ruleset_create", None, 3, None, C.byref(h), err, 256)
call("pgw_ruleset_create", None, 0, None, None, err, 256)
h=C.c_void_p(); L.pgw_ruleset_create(None,0,None,C.byref(h),err,256)
L.pgw_lists_add.argtypes=[C.c_void_p,C.c_char_p,C.c_int,C.c_char_p,C.c_size_t,C.c_char_p,C.c_size_t]
call("pgw_lists_add", h, None, 0, b"x", 1, err, 256)
call("pgw_lists_add", h, b"l", 7, b"x", 1, err, 256)
call("pgw_lists_add", h, b"l", 0, None, 0, err, 256)
call("pgw_lists_add", None, b"l", 0, b"", 0, err, 256)
L.pgw_geoip_load.argtypes=[C.c_void_p,C.c_char_p,C.c_size_t,C.c_char_p,C.c_size_t]
call("pgw_geoip_load", h, None, 0, err, 256)
call("pgw_geoip_load", h, b"garbage", 7, err, 256)
L.pgw_ruleset_finalize.argtypes=[C.c_void_p,C.c_int,C.c_char_p,C.c_size_t]
call("pgw_ruleset_finalize", None, 0, err, 256)
call("pgw_ruleset_finalize", h, 0, err, 256)
info=_ffi.Info()
L.pgw_ruleset_info.argtypes=[C.c_void_p,C.c_void_p]
call("pgw_ruleset_info", h, C.byref(info)); call("pgw_ruleset_info", None, C.byref(info)); call("pgw_ruleset_info", h, None)
L.pgw_ruleset_describe.argtypes=[C.c_void_p,C.c_char_p,C.c_size_t]; L.pgw_ruleset_describe.restype=C.c_size_t
call("pgw_ruleset_describe", h, None, 0); call("pgw_ruleset_describe", None, None, 0)
L.pgw_evaluate_batch.argtypes=[C.c_void_p,C.c_void_p,C.c_void_p,C.c_void_p]
b=_ffi.Batch()
call("pgw_evaluate_batch", h, C.byref(b), None, None); call("pgw_evaluate_batch", None, None, None, None)
L.pgw_evaluate_batch_host.argtypes=[C.c_void_p,C.c_void_p,C.c_void_p]
call("pgw_evaluate_batch_host", h, C.byref(b), None)
q=C.c_void_p()
L.pgw_queue_create.argtypes=[C.c_void_p,C.c_uint32,C.c_uint32,C.c_void_p,C.c_char_p,C.c_size_t]
call("pgw_queue_create", h, 16, 100, C.byref(q), err, 256); call("pgw_queue_create", None, 16, 100, C.byref(q), err, 256)
L.pgw_ruleset_load_dir.argtypes=[C.c_char_p,C.c_char_p,C.c_void_p,C.c_uint32,C.c_void_p,C.c_void_p,C.c_char_p,C.c_size_t]
call("pgw_ruleset_load_dir", b"/nonexistent", None, None, 0, None, C.byref(h), err, 256)
call("pgw_ruleset_load_dir", None, None, None, 0, None, C.byref(h), err, 256)
L.pgw_shape_request.argtypes=[C.c_void_p,C.c_void_p,C.c_void_p]
call("pgw_shape_request", None, None, None)
L.pgw_ruleset_destroy.argtypes=[C.c_void_p]
call("pgw_ruleset_destroy", None)
L.pgw_queue_destroy.argtypes=[C.c_void_p]
call("pgw_queue_destroy", None)
L.pgw_host_free.argtypes=[C.c_void_p]; call("pgw_host_free", None)
L.pgw_geoip_lookup_batch.argtypes=[C.c_void_p]*3+[C.c_uint32]+[C.c_void_p]*3
call("pgw_geoip_lookup_batch", None, None, None, 0, None, None, None)
L.pgw_captcha_client_id_batch.argtypes=[C.c_void_p]*3
call("pgw_captcha_client_id_batch", None, None, None)
print("survived")
