"""ctypes binding of libsuews_b200.so (include/suews_b200.h).  No PyTorch, no CPU fallback:
if the CUDA library is missing or no device is present this fails loudly."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SUEWS_B200_LIB: another build of the same library (launch-geometry tuning builds of tools/variants.sh)
LIB_PATH = os.environ.get("SUEWS_B200_LIB") or os.path.join(_HERE, "libsuews_b200.so")

_lib = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


class SuewsError(RuntimeError):
    """Raised where the reference would STOP 'Fatal error in SUEWS!' (suews_ctrl_error.f95:392-415;
    supy_driver turns that STOP into a RuntimeError, supy-driver/nonstopf2py.py:54-60)."""

    def __init__(self, code, msg):
        super().__init__("SUEWS error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "suews_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C suews_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        H = C.c_void_p
        L.suews_b200_device_count.restype = C.c_int
        L.suews_b200_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(H)]
        L.suews_b200_destroy.argtypes = [H]
        L.suews_b200_field_info.argtypes = [C.c_char_p, ip]
        L.suews_b200_set_class_param.argtypes = [H, C.c_int, C.c_char_p, dp, C.c_int]
        L.suews_b200_set_cell_param.argtypes = [H, C.c_char_p, dp, C.c_int]
        L.suews_b200_set_cell_class.argtypes = [H, ip]
        L.suews_b200_set_state.argtypes = [H, C.c_char_p, dp, C.c_int]
        L.suews_b200_get_state.argtypes = [H, C.c_char_p, dp, C.c_int]
        L.suews_b200_set_clock.argtypes = [H, C.c_int, C.c_int]
        L.suews_b200_get_clock.argtypes = [H, ip, ip]
        L.suews_b200_upload.argtypes = [H]
        L.suews_b200_resort.argtypes = [H]
        L.suews_b200_set_forcing.argtypes = [H, dp, C.c_int]
        L.suews_b200_set_estm_forcing.argtypes = [H, dp, C.c_int]
        L.suews_b200_set_anohm_forcing.argtypes = [H, dp, C.c_int]
        L.suews_b200_set_cbl_initial.argtypes = [H, dp, C.c_int]
        L.suews_b200_set_forcing_orig.argtypes = [H, dp] + [C.c_int] * 6
        L.suews_b200_apply_kdown_zen.argtypes = [H] + [C.c_double] * 4
        L.suews_b200_set_estm_forcing_orig.argtypes = [H, dp, C.c_int, C.c_int, C.c_int]
        L.suews_b200_get_estm_forcing_block.argtypes = [H, dp, C.c_int]
        L.suews_b200_get_forcing_block.argtypes = [H, dp, C.c_int]
        L.suews_b200_advance.argtypes = [H, C.c_int, C.c_int, C.c_int]
        L.suews_b200_sync.argtypes = [H]
        L.suews_b200_set_output_freq.argtypes = [H, C.c_int]
        L.suews_b200_get_agg_block.argtypes = [H, C.c_int, dp, C.c_int]
        L.suews_b200_get_agg_column.argtypes = [H, C.c_int, C.c_int, dp]
        L.suews_b200_agg_daily_rows.argtypes = [H, ip]
        L.suews_b200_get_agg_daily_block.argtypes = [H, C.c_int, dp, C.c_int]
        L.suews_b200_get_sums.argtypes = [H, dp]
        L.suews_b200_reset_sums.argtypes = [H]
        L.suews_b200_get_sums_cols.argtypes = [H, ip, C.c_int, dp]
        L.suews_b200_device_sync.argtypes = [H]
        L.suews_b200_fp64_peak.argtypes = [C.c_int, dp]
        L.suews_b200_get_block.argtypes = [H, C.c_char_p, C.c_int, dp, C.c_int, C.c_int]
        L.suews_b200_get_diag.argtypes = [H, ip, C.POINTER(C.c_ulonglong), C.POINTER(C.c_longlong)]
        L.suews_b200_last_advance_ms.argtypes = [H, C.POINTER(C.c_float)]
        L.suews_b200_device_ptrs.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]
        M = C.c_void_p
        L.suews_b200_multi_create.argtypes = [C.c_int, C.c_int, C.c_int, ip, C.POINTER(M)]
        L.suews_b200_multi_destroy.argtypes = [M]
        L.suews_b200_multi_ndev.argtypes = [M, ip]
        L.suews_b200_multi_part.argtypes = [M, C.c_int, C.POINTER(H), ip, ip]
        L.suews_b200_multi_set_class_param.argtypes = [M, C.c_int, C.c_char_p, dp, C.c_int]
        L.suews_b200_multi_set_cell_param.argtypes = [M, C.c_char_p, dp, C.c_int]
        L.suews_b200_multi_set_cell_class.argtypes = [M, ip]
        L.suews_b200_multi_set_state.argtypes = [M, C.c_char_p, dp, C.c_int]
        L.suews_b200_multi_get_state.argtypes = [M, C.c_char_p, dp, C.c_int]
        L.suews_b200_multi_set_clock.argtypes = [M, C.c_int, C.c_int]
        L.suews_b200_multi_upload.argtypes = [M]
        L.suews_b200_multi_set_forcing.argtypes = [M, dp, C.c_int]
        L.suews_b200_multi_set_forcing_orig.argtypes = [M, dp] + [C.c_int] * 6
        L.suews_b200_multi_apply_kdown_zen.argtypes = [M] + [C.c_double] * 4
        L.suews_b200_multi_set_estm_forcing.argtypes = [M, dp, C.c_int]
        L.suews_b200_multi_set_estm_forcing_orig.argtypes = [M, dp, C.c_int, C.c_int, C.c_int]
        L.suews_b200_multi_set_anohm_forcing.argtypes = [M, dp, C.c_int]
        L.suews_b200_multi_set_cbl_initial.argtypes = [M, dp, C.c_int]
        L.suews_b200_multi_get_block.argtypes = [M, C.c_char_p, C.c_int, dp, C.c_int, C.c_int]
        L.suews_b200_multi_advance.argtypes = [M, C.c_int, C.c_int, C.c_int]
        L.suews_b200_multi_sync.argtypes = [M]
        L.suews_b200_multi_last_advance_ms.argtypes = [M, C.POINTER(C.c_float)]
        L.suews_b200_multi_gather_sums.argtypes = [M, ip, C.c_int, dp, C.POINTER(C.c_float)]
        L.suews_b200_multi_get_diag.argtypes = [M, ip, C.POINTER(C.c_ulonglong), C.POINTER(C.c_longlong)]
        L.suews_cal_sunposition_c.argtypes = [C.c_double] * 6 + [dp, dp]
        L.suews_last_error.argtypes = [ip, C.c_char_p, C.c_int]
        _lib = L
    return _lib


def last_error():
    code = C.c_int()
    buf = C.create_string_buffer(256)
    lib().suews_last_error(C.byref(code), buf, 256)
    return code.value, buf.value.decode(errors="replace")


def check(rc):
    if rc != 0:
        code, msg = last_error()
        raise SuewsError(rc if rc else code, msg)
