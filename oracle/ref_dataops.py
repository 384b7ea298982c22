"""ctypes front-end of the data-side entry points of oracle/_ref/libcufd_ref.so (oracle/ref_dataops_shim.cu): the reference's OWN
window / band-pass / cross-correlation / source-update kernels (Src/utilities.cu:733-1356) run on the GPU in the order of the commented
call sites Src/libCUFD.cu:353-457.  TEST INFRASTRUCTURE ONLY (needs a GPU): pins oracle/dataops.py and the product's sepfwi_condition.
"""
import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libcufd_ref.so")
_lib = None


def available():
    if not os.path.exists(_SO):
        return False
    try:
        return hasattr(_load(), "ref_dataops_chain")
    except OSError:
        return False


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_SO)
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def bp_filter(data, dt, f):
    d = _f32(data).copy()
    nrec, nt = d.shape
    f4 = _f32(f)
    rc = _load().ref_bp_filter(C.c_int(nt), C.c_float(dt), C.c_int(nrec), C.c_void_p(d.ctypes.data), C.c_void_p(f4.ctypes.data))
    assert rc == 0, rc
    return d


def window_traces(data, dt, win_start, win_end, weights, src_weight, ratio):
    d = _f32(data).copy()
    nrec, nt = d.shape
    a, b, w = _f32(win_start), _f32(win_end), _f32(weights)
    rc = _load().ref_window_traces(C.c_int(nt), C.c_int(nrec), C.c_float(dt), C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data),
                                   C.c_void_p(w.ctypes.data), C.c_float(src_weight), C.c_float(ratio), C.c_void_p(d.ctypes.data))
    assert rc == 0, rc
    return d


def window_simple(data, dt, ratio):
    d = _f32(data).copy()
    nrec, nt = d.shape
    rc = _load().ref_window_simple(C.c_int(nt), C.c_int(nrec), C.c_float(dt), C.c_float(ratio), C.c_void_p(d.ctypes.data))
    assert rc == 0, rc
    return d


def normfact(a, b):
    a, b = _f32(a), _f32(b)
    nrec, nt = a.shape
    out = np.zeros(nrec, np.float32)
    rc = _load().ref_normfact(C.c_int(nt), C.c_int(nrec), C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data), C.c_void_p(out.ctypes.data))
    assert rc == 0, rc
    return out


def condition(obs, cal, src, dt, if_win=False, win_start=None, win_end=None, weights=None, src_weight=1.0, win_ratio=0.005, filt=None,
              if_cross_misfit=False, if_src_update=False):
    """Same arguments and result keys as oracle.dataops.condition; everything is computed by the reference's kernels."""
    o, c, s = _f32(obs).copy(), _f32(cal).copy(), _f32(src).copy()
    nrec, nt = o.shape
    ws = _f32(win_start if win_start is not None else np.zeros(nrec))
    we = _f32(win_end if win_end is not None else np.zeros(nrec))
    wt = _f32(weights if weights is not None else np.ones(nrec))
    f4 = _f32(filt if filt is not None else np.zeros(4))
    res = np.zeros_like(o)
    J, ratio = np.zeros(1, np.float32), C.c_float(1.0)
    p = lambda a: C.c_void_p(a.ctypes.data)
    rc = _load().ref_dataops_chain(C.c_int(nt), C.c_int(nrec), C.c_float(dt), p(o), p(c), p(s), C.c_int(1 if if_win else 0), p(ws), p(we), p(wt),
                                   C.c_float(src_weight), C.c_float(win_ratio), C.c_int(0 if filt is None else 1), p(f4),
                                   C.c_int(1 if if_cross_misfit else 0), C.c_int(1 if if_src_update else 0), p(res), p(J), C.byref(ratio))
    assert rc == 0, rc
    return dict(res=res, misfit=float(J[0]), cal=c, src=s, obs=o, amp_ratio=float(ratio.value))
