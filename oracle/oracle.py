"""ctypes front-end of the CPU oracle (oracle/oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.

The multi-shot drivers below restate the reference's host logic:
  * shot sharding `linspace(0, group_size, ngpu+1)` truncated to int
    (DAS_Waveform_Inversion/Ops/FWI/Src/Torch_Fwi.cpp:59-60,78-80),
  * gradients accumulated over the shots of a group and summed over groups,
    misfit = 0.5 * sum (Src/libCUFD.cu:427,776; Src/Torch_Fwi.cpp:96-101),
  * grad_stf written at the LOCAL shot index of GPU 0's group only
    (Src/libCUFD.cu:671-673; Src/Torch_Fwi.cpp:103)  [reference quirk, `ref_stf_bug`].
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OraPar(C.Structure):
    _fields_ = [("nz", C.c_int), ("nx", C.c_int), ("nPml", C.c_int), ("nPad", C.c_int),
                ("nSteps", C.c_int), ("dz", C.c_float), ("dx", C.c_float), ("dt", C.c_float),
                ("f0", C.c_float), ("mixed", C.c_int), ("fiber", C.c_int), ("src_rxz", C.c_float), ("race", C.c_int)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, os.path.join(_HERE, "liboracle.so")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "oracle.c")):
            build()
        _LIB = C.CDLL(path)
        _LIB.ora_gradient.restype = C.c_double
        _LIB.ora_courant_model.restype = C.c_float
        _LIB.ora_ring_len.restype = C.c_int
    return _LIB


def _f(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_par(nz, nx, nPml, nPad, nSteps, dz, dx, dt, f0, mixed=1, fiber=0, src_rxz=1.0, race=0):
    return OraPar(nz, nx, nPml, nPad, nSteps, dz, dx, dt, f0, mixed, fiber, src_rxz, race)


def shard_bounds(group_size, ngpu):
    """Torch_Fwi.cpp:59-60,78-80: float32 linspace truncated to int."""
    bars = np.linspace(0, group_size, ngpu + 1, dtype=np.float32)
    return [int(b) for b in bars]


def cpml(N, nPml, dh, f0, dt):
    out = [np.zeros(N, np.float32) for _ in range(6)]
    lib().ora_cpml(C.c_int(N), C.c_int(nPml), C.c_float(dh), C.c_float(f0), C.c_float(dt), *[_f(o) for o in out])
    return dict(zip(["K", "a", "b", "K_half", "a_half", "b_half"], out))


def model_prep(lam_mpa, mu_mpa, den):
    nz, nx = lam_mpa.shape
    out = [np.zeros((nz, nx), np.float32) for _ in range(6)]
    lib().ora_model_prep(C.c_int(nz), C.c_int(nx), _f(_c32(lam_mpa)), _f(_c32(mu_mpa)), _f(_c32(den)),
                         *[_f(o) for o in out])
    return dict(zip(["lam", "mu", "muave", "byca", "bycb", "cp"], out))


def stf_taper(stf, dt, ratio=0.001):
    s = np.ascontiguousarray(stf, np.float32).copy()
    lib().ora_stf_taper(C.c_int(s.size), C.c_float(dt), C.c_float(ratio), _f(s))
    return s


def ring_len(par):
    return lib().ora_ring_len(C.byref(par))


def ring_cells(par):
    """(len,2) int array of (z,x) per ring index, Src/utilities.cu:362-392."""
    n = ring_len(par)
    out = np.zeros((n, 2), np.int32)
    z, x = C.c_int(), C.c_int()
    L = lib()
    for i in range(n):
        L.ora_ring_cell(C.byref(par), C.c_int(i), C.byref(z), C.byref(x))
        out[i] = (z.value, x.value)
    return out


def courant(par, lam_mpa, mu_mpa, den):
    return float(lib().ora_courant_model(C.byref(par), _f(_c32(lam_mpa)), _f(_c32(mu_mpa)), _f(_c32(den))))


def _c32(a):
    return np.ascontiguousarray(a, np.float32)


def forward(par, lam_mpa, mu_mpa, den, stf_row, zs, xs, zrec, xrec, comps=("pr", "vx", "vz", "ett"),
            want_ring=False, want_fields=False):
    """One shot of forward modelling (cufd calc_id=2).  zs/xs/zrec/xrec are INTERIOR
    indices as in survey_file.json; +nPml is applied here (Src/Src_Rec.cu:87-115)."""
    lam_mpa, mu_mpa, den = _c32(lam_mpa), _c32(mu_mpa), _c32(den)
    zr = np.ascontiguousarray(np.asarray(zrec) + par.nPml, np.int32)
    xr = np.ascontiguousarray(np.asarray(xrec) + par.nPml, np.int32)
    nrec = zr.size
    out = {k: (np.zeros((nrec, par.nSteps), np.float32) if k in comps else None) for k in ("pr", "vx", "vz", "ett")}
    bnd = np.zeros((5, par.nSteps, ring_len(par)), np.float32) if want_ring else None
    fields = np.zeros((5, par.nz, par.nx), np.float32) if want_fields else None
    lib().ora_forward(C.byref(par), _f(lam_mpa), _f(mu_mpa), _f(den), _f(_c32(stf_row)),
                      C.c_int(zs + par.nPml), C.c_int(xs + par.nPml), C.c_int(nrec), _i(zr), _i(xr),
                      _f(out["pr"]), _f(out["vx"]), _f(out["vz"]), _f(out["ett"]), _f(bnd), _f(fields))
    res = {k: v for k, v in out.items() if v is not None}
    if want_ring:
        res["ring"] = bnd
    if want_fields:
        res["fields"] = fields
    return res


def gradient_shot(par, lam_mpa, mu_mpa, den, stf_row, zs, xs, zrec, xrec, obs_ett, with_adj=True, acc=None):
    """One shot of misfit (+gradient).  Returns dict; gradients accumulate into `acc` if given."""
    lam_mpa, mu_mpa, den = _c32(lam_mpa), _c32(mu_mpa), _c32(den)
    zr = np.ascontiguousarray(np.asarray(zrec) + par.nPml, np.int32)
    xr = np.ascontiguousarray(np.asarray(xrec) + par.nPml, np.int32)
    nrec = zr.size
    shape = (par.nz, par.nx)
    if acc is None:
        acc = {k: np.zeros(shape, np.float32) for k in ("glam", "gmu", "gden")}
    gstf = np.zeros(par.nSteps, np.float32)
    syn = np.zeros((nrec, par.nSteps), np.float32)
    res = np.zeros((nrec, par.nSteps), np.float32)
    J = lib().ora_gradient(C.byref(par), _f(lam_mpa), _f(mu_mpa), _f(den), _f(_c32(stf_row)),
                           C.c_int(zs + par.nPml), C.c_int(xs + par.nPml), C.c_int(nrec), _i(zr), _i(xr),
                           _f(_c32(obs_ett)), C.c_int(1 if with_adj else 0),
                           _f(acc["glam"]), _f(acc["gmu"]), _f(acc["gden"]), _f(gstf), _f(syn), _f(res))
    return dict(J=J, gstf=gstf, syn_ett=syn, res_ett=res, **acc)


def fwi_backward(par, lam_mpa, mu_mpa, den, stf, ngpu, shot_ids, survey, obs_ett, ref_stf_bug=False):
    """Restatement of fwi_ops.backward (Torch_Fwi.cpp:38-104) on top of gradient_shot.
    survey: dict shot_id -> (zs, xs, zrec, xrec); obs_ett: dict shot_id -> [nrec][nSteps].
    Returns misfit, glam, gmu, gden, gstf (nSrc,nSteps)."""
    shot_ids = list(np.asarray(shot_ids).tolist())
    if ngpu > len(shot_ids):
        raise RuntimeError("The number of GPUs should be smaller than the number of shots!")
    bars = shard_bounds(len(shot_ids), ngpu)
    shape = (par.nz, par.nx)
    tot = {k: np.zeros(shape, np.float32) for k in ("glam", "gmu", "gden")}
    gstf = np.zeros_like(np.asarray(stf, np.float32))
    misfit = np.float32(0.0)
    for g in range(ngpu):
        acc = {k: np.zeros(shape, np.float32) for k in ("glam", "gmu", "gden")}
        h = np.float32(0.0)
        for local, sid in enumerate(shot_ids[bars[g]:bars[g + 1]]):
            zs, xs, zrec, xrec = survey[sid]
            r = gradient_shot(par, lam_mpa, mu_mpa, den, np.asarray(stf, np.float32)[sid], zs, xs, zrec, xrec,
                              obs_ett[sid], True, acc)
            h = np.float32(h + np.float32(r["J"]))
            if ref_stf_bug:
                if g == 0:
                    gstf[local] = r["gstf"]
            else:
                gstf[sid] = r["gstf"]
        misfit = np.float32(misfit + np.float32(0.5) * h)
        for k in tot:
            tot[k] += acc[k]
    return float(misfit), tot["glam"], tot["gmu"], tot["gden"], gstf


# --------------------------------------------------------------------------- Numba flavour
def numba_damp(nx_pad, nz_pad, ndamp):
    """Sponge profile, DAS_Waveform_Modeling/src/elasticSolver.py:74-79."""
    damp = np.ones((nx_pad, nz_pad))
    for i in range(ndamp):
        w = np.sin(np.pi / 2 * i / ndamp) ** 2
        damp[i, :] *= w
        damp[-i - 1, :] *= w
        damp[:, i] *= w
        damp[:, -i - 1] *= w
    return damp


def numba_forward(nx, nz, ndamp, dx, dz, dt, nt, f0, vp, vs, rho, src_coord, das_coord, geo_coord,
                  das_sensitivity, isrc=0):
    """Restatement of elasticSolver(...).forward_it(isrc, False)
    (DAS_Waveform_Modeling/src/elasticSolver.py:33-92,185-305).  Arrays are (nx, nz)."""
    vp = np.pad(np.asarray(vp, np.float64), ndamp, "edge")
    vs = np.pad(np.asarray(vs, np.float64), ndamp, "edge")
    rho = np.pad(np.asarray(rho, np.float64), ndamp, "edge")
    NX, NZ = nx + 2 * ndamp, nz + 2 * ndamp
    mu = rho * vs ** 2
    lam = rho * vp ** 2 - 2 * mu
    t = np.arange(0, nt * dt, dt)
    stf = (1.0 - 2.0 * np.pi ** 2 * f0 ** 2 * (t - 1.2 / f0) ** 2) * np.exp(-np.pi ** 2 * f0 ** 2 * (t - 1.2 / f0) ** 2)

    def grid(c):
        c = np.asarray(c, np.float64)
        return (np.round(c[:, 0] / dx).astype(np.int32) + ndamp, np.round(c[:, 1] / dz).astype(np.int32) + ndamp)

    sx, sz = grid(src_coord)
    gx, gz = grid(geo_coord)
    ddx, ddz = grid(das_coord)
    damp = numba_damp(NX, NZ, ndamp)
    sens = np.ascontiguousarray(das_sensitivity, np.float64)
    ng, nd = gx.size, ddx.size
    o = {k: np.zeros((ng, nt)) for k in ("vx", "vz", "pr")}
    o.update({k: np.zeros((nd, nt)) for k in ("exx", "ezz", "exz", "ett")})
    cc = np.ascontiguousarray
    lib().ora_numba_forward(C.c_int(NX), C.c_int(NZ), C.c_double(dx), C.c_double(dz), C.c_double(dt), C.c_int(nt),
                            _d(cc(lam)), _d(cc(mu)), _d(cc(rho)), _d(cc(damp)), _d(cc(stf[:nt])),
                            C.c_int(int(sx[isrc])), C.c_int(int(sz[isrc])),
                            C.c_int(ng), _i(cc(gx)), _i(cc(gz)), C.c_int(nd), _i(cc(ddx)), _i(cc(ddz)), _d(sens),
                            _d(o["vx"]), _d(o["vz"]), _d(o["pr"]), _d(o["exx"]), _d(o["ezz"]), _d(o["exz"]),
                            _d(o["ett"]))
    o["t"] = t
    return o
