"""ctypes wrapper of oracle/cmix_ref.c (TEST INFRASTRUCTURE / CPU baseline, see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libcmix_ref.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH) or os.path.getmtime(_PATH) < os.path.getmtime(os.path.join(_HERE, "cmix_ref.c")):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_PATH)
        _lib.sfbo_wigner3j000.restype = C.c_double
        _lib.sfbo_wigner3j000.argtypes = [C.c_int64] * 3
        _lib.sfbo_max_threads.restype = C.c_int32
    return _lib


def max_threads():
    return int(load().sfbo_max_threads())


def wigner3j000(l, lp, L):
    return load().sfbo_wigner3j000(l, lp, L)


def calc_wrl_wrl(W1_mfast, W2_mfast, LMAX):
    """[LMAX+1, nr, nr] (index [L1, j, i] = Julia W[i,j,L1])."""
    W1 = np.asfortranarray(W1_mfast, dtype=np.complex128)
    W2 = np.asfortranarray(W2_mfast, dtype=np.complex128)
    nr = W1.shape[0]
    out = np.empty((LMAX + 1, nr, nr))
    load().sfbo_calc_wrl_wrl(C.c_void_p(W1.ctypes.data), C.c_void_p(W2.ctypes.data), C.c_int64(nr), C.c_int64(LMAX),
                             C.c_void_p(out.ctypes.data))
    return out


def calc_cmix_rows(lnn, rows, G, W_julia, div2Lp1=False, interchange=False, col_lo=1, col_hi=None, nthreads=0):
    """Rows `rows` (1-based) x columns [col_lo, col_hi] of M.  G: (nr, nmax, lmax+1); W_julia: [LMAX+1, nr(j), nr(i)]
    as returned by calc_wrl_wrl (memory = Julia's W[i,j,L1])."""
    lnn = np.asfortranarray(lnn, dtype=np.int64)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    G = np.asfortranarray(G, dtype=np.float64)
    W = np.ascontiguousarray(W_julia, dtype=np.float64)
    nr, nmax = G.shape[0], G.shape[1]
    col_hi = lnn.shape[1] if col_hi is None else col_hi
    out = np.empty((rows.size, col_hi - col_lo + 1), order="F")
    load().sfbo_calc_cmix_rows(C.c_void_p(lnn.ctypes.data), C.c_void_p(rows.ctypes.data), C.c_int64(rows.size),
                               C.c_int64(col_lo), C.c_int64(col_hi), C.c_void_p(G.ctypes.data), C.c_int64(nr),
                               C.c_int64(nmax), C.c_void_p(W.ctypes.data), C.c_int32(int(div2Lp1)),
                               C.c_int32(int(interchange)), C.c_int32(nthreads), C.c_void_p(out.ctypes.data))
    return out
