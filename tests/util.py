"""Helpers shared by the test modules."""
import numpy as np


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    n = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / n) if n > 0 else float(np.linalg.norm(a))


def oracle_par(O, prob, mixed=1, race=0):
    return O.make_par(prob.nz, prob.nx, prob.nPml, prob.nPad, prob.nSteps, prob.dz, prob.dx, prob.dt, prob.f0,
                      mixed=mixed, fiber=prob.fiber, race=race)


def oracle_observed(O, prob, comps=("ett",)):
    """Observed data of the true model for every shot: dict sid -> dict comp -> [nrec][nSteps]."""
    par = oracle_par(O, prob)
    out = {}
    for sid, (zs, xs, zr, xr) in prob.survey().items():
        out[sid] = O.forward(par, *prob.true, prob.stf[sid], zs, xs, zr, xr, comps=comps)
    return out


def cuda_shots(prob, ShotSpec, ids=None):
    ids = range(prob.nshots) if ids is None else ids
    return [ShotSpec(prob.z_src[i] + prob.nPml, prob.x_src[i] + prob.nPml, prob.z_rec + prob.nPml,
                     prob.x_rec + prob.nPml, prob.stf[i]) for i in ids]


def make_prop(Propagator, prob, **kw):
    return Propagator(prob.nz, prob.nx, prob.nPml, prob.nPad, prob.nSteps, prob.dz, prob.dx, prob.dt, prob.f0,
                      fiber=prob.fiber, max_nrec=len(prob.x_rec), **kw)
