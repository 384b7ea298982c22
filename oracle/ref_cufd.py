"""ctypes front-end of oracle/_ref/libcufd_ref.so -- the reference's OWN CUDA shot driver
(`cufd`, DAS_Waveform_Inversion/Ops/FWI/Src/libCUFD.cu:32) compiled in place from
/root/reference by oracle/Makefile.  TEST INFRASTRUCTURE ONLY: used on the GPU box to pin the
oracle and the product against the real reference (tests/golden/make_cufd_golden.py and the
`-m gpu` parity tests).  Needs a GPU; the reference calls exit(1) on any error.
"""
import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libcufd_ref.so")
_lib = None


def available():
    return os.path.exists(_SO)


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_SO)
        _lib.ref_cufd.argtypes = [C.c_void_p] * 9 + [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_char_p]
        _lib.ref_cufd.restype = None
    return _lib


def cufd(calc_id, lam, mu, den, stf, shot_ids, para_fname, gpu_id=0):
    """Returns (misfit, glam, gmu, gden, gstf); gradients are zero arrays unless calc_id == 1."""
    lam, mu, den = (np.ascontiguousarray(a, np.float32) for a in (lam, mu, den))
    stf = np.ascontiguousarray(stf, np.float32)
    ids = np.ascontiguousarray(shot_ids, np.int32)
    misfit = np.zeros(1, np.float32)
    g = [np.zeros_like(lam) for _ in range(3)]
    gstf = np.zeros_like(stf)
    p = lambda a: a.ctypes.data
    _load().ref_cufd(p(misfit), p(g[0]), p(g[1]), p(g[2]), p(gstf), p(lam), p(mu), p(den), p(stf),
                     int(calc_id), int(gpu_id), int(ids.size), p(ids), para_fname.encode())
    return float(misfit[0]), g[0], g[1], g[2], gstf
