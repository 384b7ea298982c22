"""ctypes front-end of oracle/_ref/libcufd_ref.so -- the reference's OWN CUDA shot driver
(`cufd`, DAS_Waveform_Inversion/Ops/FWI/Src/libCUFD.cu:32) compiled in place from
/root/reference by oracle/Makefile.  TEST INFRASTRUCTURE ONLY: used on the GPU box to pin the
oracle and the product against the real reference (tests/golden/make_cufd_golden.py and the
`-m gpu` parity tests).  Needs a GPU; the reference calls exit(1) on any error.
"""
import ctypes as C
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
# fiber 0: the stock library (horizontal fiber, recording_exx / res_injection_exx); fiber 1: the same sources with the reference's
# two call sites switched to recording_ezz / res_injection_ezz by macro definitions on the compile line (oracle/Makefile)
_SO = {0: os.path.join(_DIR, "libcufd_ref.so"), 1: os.path.join(_DIR, "libcufd_ref_ezz.so")}
_libs = {}


def available(fiber=0):
    return os.path.exists(_SO[fiber])


def _load(fiber=0):
    if fiber not in _libs:
        lib = C.CDLL(_SO[fiber])
        lib.ref_cufd.argtypes = [C.c_void_p] * 9 + [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_char_p]
        lib.ref_cufd.restype = None
        _libs[fiber] = lib
    return _libs[fiber]


def cufd(calc_id, lam, mu, den, stf, shot_ids, para_fname, gpu_id=0, fiber=0):
    """Returns (misfit, glam, gmu, gden, gstf); gradients are zero arrays unless calc_id == 1."""
    lam, mu, den = (np.ascontiguousarray(a, np.float32) for a in (lam, mu, den))
    stf = np.ascontiguousarray(stf, np.float32)
    ids = np.ascontiguousarray(shot_ids, np.int32)
    misfit = np.zeros(1, np.float32)
    g = [np.zeros_like(lam) for _ in range(3)]
    gstf = np.zeros_like(stf)
    p = lambda a: a.ctypes.data
    _load(fiber).ref_cufd(p(misfit), p(g[0]), p(g[1]), p(g[2]), p(gstf), p(lam), p(mu), p(den), p(stf),
                          int(calc_id), int(gpu_id), int(ids.size), p(ids), para_fname.encode())
    return float(misfit[0]), g[0], g[1], g[2], gstf
