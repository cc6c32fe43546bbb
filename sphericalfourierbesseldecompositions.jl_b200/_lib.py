"""ctypes binding of the C ABI in include/sfb_b200.h.

The shared library is built in-tree (`python -m sfb_b200.build` / `__graft_entry__.build()`).  There is no
CPU fallback: if the library is missing, or no CUDA device is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsfb_b200.so")

_i32, _i64, _f64p, _i64p, _vp = C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p

# name -> (restype, argtypes); must list every symbol include/sfb_b200.h declares
SIGNATURES = {
    "sfb_version": (_i32, []),
    "sfb_last_error": (C.c_char_p, []),
    "sfb_device_count": (_i32, [C.POINTER(_i32)]),
    "sfb_set_device": (_i32, [_i32]),
    "sfb_set_devices": (_i32, [_i32]),
    "sfb_get_devices": (_i32, [C.POINTER(_i32)]),
    "sfb_host_alloc": (_i32, [C.POINTER(_vp), _i64]),
    "sfb_host_free": (_i32, [_vp]),
    "sfb_host_register": (_i32, [_vp, _i64]),
    "sfb_host_unregister": (_i32, [_vp]),
    "sfb_get_timings": (_i32, [_f64p, _i32]),
    "sfb_probe_dmma_tflops": (_i32, [_f64p]),
    "sfb_calc_wr_lm": (_i32, [_f64p, _i64, _i64, _i64, _i64, _i64, _i64, _i32, _f64p]),
    "sfb_calc_wlm_mask": (_i32, [_f64p, _i64, _i64, _i64, _i64, _f64p]),
    "sfb_power_win_mix_from_wrlm": (_i32, [_f64p, _f64p, _i64, _i64, _i32, _f64p, _i64, _i64, _i64p, _i64, _i64,
                                           _i32, _i32, _f64p]),
    "sfb_power_win_mix": (_i32, [_f64p, _f64p, _i64, _i64, _i64, _i64, _f64p, _i64, _i64, _i64p, _i64, _i64, _i32,
                                 _i32, _f64p]),
    "sfb_power_win_mix_binned": (_i32, [_f64p, _i64, _i64, _i64, _i64, _f64p, _i64, _i64, _i64p, _i64,
                                        _i64p, _i64p, _f64p, _i64, _i64p, _i64p, _f64p, _i64, _i32, _i32, _f64p]),
    "sfb_power_win_mix_separable": (_i32, [_f64p, _f64p, _i64, _i64, _i64, _f64p, _i64, _i64, _i64p, _i64,
                                           _i64p, _i64p, _f64p, _i64, _i64p, _i64p, _f64p, _i64, _i32, _i32, _f64p]),
    "sfb_solve": (_i32, [_f64p, _i64, _f64p, _i64, _f64p]),
    "sfb_power_win_mix_binned_solve": (_i32, [_f64p, _i64, _i64, _i64, _i64, _f64p, _i64, _i64, _i64p, _i64,
                                              _i64p, _i64p, _f64p, _i64, _i64p, _i64p, _f64p, _i64, _i32, _i32,
                                              _f64p, _i64, _f64p, _f64p]),
    "sfb_win_lnn": (_i32, [_f64p, _i64, _i64, _i64, _i64, _f64p, _i64, _i64, _i64p, _i64, _f64p]),
    "sfb_calc_wmix": (_i32, [_f64p, _i64, _i64, _i64, _i64, _f64p, _i64, _i64, _i64p, _i64p, _i32, _f64p]),
    "sfb_field2anlm": (_i32, [_f64p, _i64, _i64, _i64, _f64p, _i64, _i64, _i64p, _i64p, _f64p]),
    "sfb_anlm2field": (_i32, [_f64p, _f64p, _i64, _i64, _i64, _i64, _i64p, _i64p, _f64p, _i64]),
    "sfb_win_rhat_ln": (_i32, [_f64p, _i64, _i64, _i64, _f64p, _i64, _i64, _i64p, _i64p, _f64p]),
    "sfb_cat2amln": (_i32, [_i64p, _i64p, _i64, _f64p, _i64p, _i64p, _i64, C.c_double, _f64p, _i64, _i64, _i64, _i64p,
                            _i64p, _f64p]),
    "sfb_win_lnn_dev": (_i32, [_vp, _f64p, _f64p, _vp]),
    "sfb_sht_plan_create": (_i32, [C.POINTER(_vp), _i64, _i64, _i64, _i64]),
    "sfb_sht_plan_destroy": (_i32, [_vp]),
    "sfb_sht_alm_doubles": (_i64, [_vp]),
    "sfb_calc_wr_lm_dev": (_i32, [_vp, _f64p, _i64, _i64, _f64p, _vp]),
    "sfb_alm_to_complex_dev": (_i32, [_vp, _f64p, _i32, _f64p, _vp]),
    "sfb_alm_gather_shards_dev": (_i32, [_vp, _i64p, _i64p, _i32, _i64, _i64, _f64p, _vp]),
    "sfb_cmix_plan_create": (_i32, [C.POINTER(_vp), _i64p, _i64, _i64, _f64p, _i64, _i64, _i64]),
    "sfb_cmix_plan_destroy": (_i32, [_vp]),
    "sfb_power_win_mix_dev": (_i32, [_vp, _f64p, _f64p, _i32, _i32, _i64, _i64, _f64p, _i64, _vp]),
    "sfb_cmix_row_costs": (_i32, [_vp, _f64p, _i64]),
    "sfb_cmix_col_costs": (_i32, [_vp, _f64p, _i64]),
    "sfb_cmix_col_costs_upper": (_i32, [_vp, _f64p, _i64]),
    "sfb_cmix_packed_offsets": (_i32, [_vp, _vp, _i64]),
    "sfb_power_win_mix_upper_packed_dev": (_i32, [_vp, _f64p, _i32, _i32, _i64, _i64, _f64p, _vp]),
    "sfb_cmix_mirror_rows_dev": (_i32, [_vp, _f64p, _i64, _i64, _i32, _i32, _f64p, _i64, _vp]),
    "sfb_cmix_unpack_mirror_dev": (_i32, [_vp, _f64p, _i32, _i32, _f64p, _i64, _vp]),
    "sfb_cmix_unpack_mirror_peers_dev": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _f64p, _i64, _vp]),
    "sfb_power_win_mix_block_dev": (_i32, [_vp, _f64p, _f64p, _i32, _i32, _i64, _i64, _i64, _i64, _f64p, _i64, _vp]),
    "sfb_ipc_alloc": (_i32, [C.POINTER(_vp), _i64, _vp]),
    "sfb_ipc_open": (_i32, [_vp, C.POINTER(_vp)]),
    "sfb_ipc_close": (_i32, [_vp]),
    "sfb_ipc_free": (_i32, [_vp]),
    "sfb_memcpy_dev": (_i32, [_vp, _vp, _i64, _vp]),
}

_lib = None


class SFBError(RuntimeError):
    """Raised for any nonzero status from the library (the Julia shim raises ErrorException)."""


def load():
    """Load libsfb_b200.so; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SFBError(f"{LIB_PATH} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; "
                       f"g.build()').  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise SFBError(load().sfb_last_error().decode("utf-8", "replace"))


def ptr(a):
    """Host pointer of a numpy array (None -> NULL)."""
    return None if a is None else a.ctypes.data


def timings():
    import numpy as np
    out = np.zeros(9)
    check(load().sfb_get_timings(ptr(out), 9))
    keys = ["stage1_ms", "wl_ms", "fill_ms", "what_ms", "block_ms", "block_flops", "launches", "binned_ms", "k3_ms"]
    return dict(zip(keys, out.tolist()))
