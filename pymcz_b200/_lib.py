"""ctypes binding of libmcz_b200.so (the C ABI in include/mcz_b200.h).

There is no CPU fallback: if the library is missing it is a hard error telling the user to
build it (``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C pymcz_b200/csrc``).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCZ_B200_LIB: an instrumented / experimental build of the same library (tools/build_variant.sh)
LIB_PATH = os.environ.get("MCZ_B200_LIB") or os.path.join(_HERE, "libmcz_b200.so")

_lib = None

c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_i32p = ctypes.POINTER(ctypes.c_int)
c_i8p = ctypes.POINTER(ctypes.c_byte)
c_u8p = ctypes.POINTER(ctypes.c_ubyte)
c_u64p = ctypes.POINTER(ctypes.c_ulonglong)

#: every symbol include/mcz_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "mcz_version": (ctypes.c_char_p, []),
    "mcz_last_error": (ctypes.c_char_p, []),
    "mcz_key_name": (ctypes.c_char_p, [ctypes.c_int]),
    "mcz_line_name": (ctypes.c_char_p, [ctypes.c_int]),
    "mcz_device_count": (ctypes.c_int, []),
    "mcz_launch_count": (ctypes.c_longlong, []),
    "mcz_replay_scratch_bytes": (ctypes.c_size_t, [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int]),
    "mcz_forward_replay": (ctypes.c_int, [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_uint,
        ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "mcz_forward_replay_aux": (ctypes.c_int, [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_uint,
        ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "mcz_segmented_percentiles": (ctypes.c_int, [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "mcz_histogram_assist": (ctypes.c_int, [
        ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "mcz_ks_2samp": (ctypes.c_int, [
        ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p]),
    "mcz_production_scratch_bytes": (ctypes.c_size_t, [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_uint]),
    "mcz_production_kernel_name": (ctypes.c_char_p, [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_uint]),
    "mcz_production_last_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "mcz_forward_production": (ctypes.c_int, [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_uint,
        ctypes.c_int, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_int,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "mcz_forward_production_gather": (ctypes.c_int, [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_uint,
        ctypes.c_int, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_int,
        ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "mcz_samples_shared_bytes": (ctypes.c_size_t, [ctypes.c_uint]),
    "mcz_samples_scratch_bytes": (ctypes.c_size_t, [ctypes.c_longlong, ctypes.c_uint]),
    "mcz_forward_production_samples": (ctypes.c_int, [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_uint,
        ctypes.c_int, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_int,
        ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "mcz_ctx_create": (ctypes.c_void_p, [ctypes.c_int]),
    "mcz_ctx_destroy": (None, [ctypes.c_void_p]),
    "mcz_ctx_run_host": (ctypes.c_int, [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong,
        ctypes.c_uint, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_int,
        ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p]),
}

#: every symbol include/mcz_b200_debug.h declares (test / profiling entry points)
DEBUG_SIGNATURES = {
    "mcz_debug_select_columns": (ctypes.c_int, [
        ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "mcz_debug_production_samples": (ctypes.c_int, [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_uint,
        ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p]),
    "mcz_debug_counters": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "mcz_debug_phase_cycles": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "mcz_debug_col_cycles": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
}


class MczLibraryError(RuntimeError):
    pass


def lib():
    """The loaded library (loaded once)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MczLibraryError(
            "%s not found: the CUDA library is not built and pymcz_b200 has no CPU fallback. "
            "Build it with `make -C pymcz_b200/csrc` (needs nvcc, sm_100a)." % LIB_PATH)
    l = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in list(SIGNATURES.items()) + list(DEBUG_SIGNATURES.items()):
        fn = getattr(l, name)          # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


def check(rc, what=""):
    if rc != 0:
        raise MczLibraryError("%s failed (%d): %s" % (what, rc, lib().mcz_last_error().decode()))
