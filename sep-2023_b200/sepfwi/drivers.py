"""End-to-end inversion drivers: the reference's three anomaly experiments
(DAS_Waveform_Inversion/notebooks/00{1,2,3}-*.ipynb cell 3 for the models, Main-00{1,2,3}-*.py for the workflow)
on top of sepfwi.FWI_ops + sepfwi.obj_wrapper + scipy's L-BFGS-B, with the reference's optimiser options
(Main-001-FWI-Anomaly-Vp-Vs-Den.py:157-168).

    prob = anomaly_problem("001")                       # Vp / Vs / Den   ("002": Lame / Den, "003": IP / IS / Den)
    files = write_files(prob, exp_dir)                  # para_file.json, survey_file.json (fwi_utils.paraGen / surveyGen)
    generate_data(prob, files, ngpu=1)                  # FWI_obscalc -> Data/Shot_*.bin
    fwi, obj, log = invert(prob, files, nIter=5, ngpu=1)

Models are built in memory exactly as the notebooks build them before `np.savetxt` / `np.loadtxt(...).astype('float32')`.
"""
import os
import time

import numpy as np
import torch

from . import FWI_ops as F
from . import fwi_utils as ft
from .obj_wrapper import PyTorchObjective


class AnomalyProblem(object):
    nz, nx, dz, dx, dt, nt, nPml, f0 = 101, 201, 20.0, 20.0, 0.002, 1501, 32, 10.0

    def __init__(self, kind):
        self.kind = kind
        nx, nz = self.nx, self.nz
        vp = np.ones((nx, nz)) * 4000.0
        vs = np.ones((nx, nz)) * 4000.0 / 1.732
        rho = np.ones((nx, nz)) * 2500.0
        if kind == "001":            # 001-...ipynb cell 3: boxes in Vp, Vs, rho
            vp[42:58, 42:58] += 80.0
            vs[92:108, 42:58] -= 80.0 / 1.732
            rho[142:158, 42:58] += 40
        elif kind == "002":          # 002-...ipynb cell 3: +-2.5 % in lambda / mu, +2 % in rho
            mu = rho * vs ** 2
            lam = rho * vp ** 2 - 2 * mu
            lam[42:58, 42:58] += lam[0, 0] * 0.025
            mu[92:108, 42:58] -= mu[0, 0] * 0.025
            rho[142:158, 42:58] += rho[0, 0] * 0.020
            vp = np.sqrt((lam + 2 * mu) / rho)
            vs = np.sqrt(mu / rho)
        elif kind == "003":          # 003-...ipynb cell 3: +-2.5 % in the impedances, +50 in rho
            IP, IS = vp * rho, vs * rho
            IP[42:58, 42:58] += IP[0, 0] * 0.025
            IS[92:108, 42:58] -= IS[0, 0] * 0.025
            rho[142:158, 42:58] += 50
            vp, vs = IP / rho, IS / rho
        else:
            raise ValueError("kind must be '001', '002' or '003'")
        # saved transposed with savetxt, loaded back as float32 (Main-00x:78-80,113-115)
        self.true = [a.T.astype("float32") for a in (vp, vs, rho)]
        self.init = [np.full((nz, nx), v, "float32") for v in (4000.0, 4000.0 / 1.732, 2500.0)]
        self.nz_pad, self.nx_pad, self.nPad = ft.padded_shape(nz, nx, self.nPml)
        self.x_src = np.arange(10, nx - 10, 10).astype(int)
        self.z_src = np.ones(len(self.x_src), int)
        self.x_rec = np.arange(10, nx - 10).astype(int)
        self.z_rec = 95 * np.ones(len(self.x_rec), int)
        mask = np.zeros((self.nz_pad, self.nx_pad))
        mask[self.nPml:self.nPml + nz, self.nPml:self.nPml + nx] = 1.0
        mask[self.nPml:self.nPml + 4, :] = 0.0
        self.mask = mask
        self.stf = torch.tensor(ft.sourceGene(self.f0, self.nt, self.dt), dtype=torch.float32).repeat(len(self.x_src), 1)
        self.shot_ids = torch.tensor(np.arange(len(self.x_src)), dtype=torch.int32)


def anomaly_problem(kind="001"):
    return AnomalyProblem(kind)


def write_files(prob, exp_dir, **para_kw):
    os.makedirs(exp_dir, exist_ok=True)
    files = dict(para=os.path.join(exp_dir, "para_file.json"), survey=os.path.join(exp_dir, "survey_file.json"),
                 data=os.path.join(exp_dir, "Data"))
    ft.paraGen(prob.nz_pad, prob.nx_pad, prob.dz, prob.dx, prob.nt, prob.dt, prob.f0, prob.nPml, prob.nPad,
               files["para"], files["survey"], files["data"], **para_kw)
    ft.surveyGen(prob.z_src, prob.x_src, prob.z_rec, prob.x_rec, files["survey"])
    return files


def generate_data(prob, files, ngpu=1, device=None):
    pads = [torch.tensor(ft.padding_numpy_array(a, prob.nPml, prob.nPad), dtype=torch.float32, device=device) for a in prob.true]
    F.FWI_obscalc(*pads, prob.stf, files["para"])(prob.shot_ids, ngpu=ngpu)


def build_fwi(prob, files, device=None):
    """The FWI module of the experiment: 001 FWI(Vp, Vs, Den), 002 FWI_Lame_Den, 003 FWI_IP_IS_Den (Main-00x:113-133)."""
    opt = dict(nz=prob.nz, nx=prob.nx, nz_orig=prob.nz, nx_orig=prob.nx, nPml=prob.nPml, nPad=prob.nPad, para_fname=files["para"])
    vp, vs, den = prob.init
    if prob.kind == "001":
        fields, cls = (vp, vs, den), F.FWI
    elif prob.kind == "002":
        fields, cls = (den * (vp ** 2 - 2.0 * vs ** 2) / 1e6, den * vs ** 2 / 1e6, den), F.FWI_Lame_Den
    else:
        fields, cls = (vp / 1e3 * den, vs / 1e3 * den, den), F.FWI_IP_IS_Den
    th = [torch.tensor(a, dtype=torch.float32, device=device, requires_grad=True) for a in fields]
    return cls(*th, prob.stf, opt, Mask=torch.tensor(prob.mask, dtype=torch.float32, device=device))


def invert(prob, files, nIter=5, ngpu=1, device=None, callback=None, disp=False):
    """L-BFGS-B with the reference's options; returns (module, objective, log) where log lists (iterate, f, max|g|, seconds)
    at every accepted iterate -- the quantities scipy prints as `At iterate k  f= ...  |proj g|= ...`."""
    from scipy import optimize
    fwi = build_fwi(prob, files, device)
    obj = PyTorchObjective(fwi, lambda: fwi(prob.shot_ids, ngpu=ngpu))
    t0 = time.perf_counter()
    log = [(0, obj.fun(obj.x0), float(np.abs(obj.jac(obj.x0)).max()), time.perf_counter() - t0)]

    def cb(x):
        log.append((len(log), obj.fun(x), float(np.abs(obj.jac(x)).max()), time.perf_counter() - t0))
        if callback is not None:
            callback(x, fwi, obj)

    options = {'gtol': 1e-16, 'maxiter': nIter, 'ftol': 1e-12, 'maxcor': 5, 'maxfun': 1500, 'maxls': 6}
    if disp:
        options.update(disp=True, iprint=101)
    res = optimize.minimize(obj.fun, obj.x0, method='L-BFGS-B', jac=obj.jac, bounds=obj.bounds, tol=None,
                            callback=cb, options=options)
    obj.result = res
    return fwi, obj, log
