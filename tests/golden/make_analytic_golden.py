"""Golden vectors for BASELINE.json configs[0]: the reference's Numba solver on the 2-D homogeneous elastic 201 x 201 grid with
an explosive (isotropic moment tensor) source, next to the reference's Aki & Richards analytical solution -- the independent
check of DAS_Waveform_Modeling/notebooks/000-Solver-Benchmark.ipynb (cells 4-13).

Runs only in the build container (needs /root/reference).  Imports DAS_Waveform_Modeling/src/elasticSolver.py and
analyticalSolution.py unmodified (matplotlib stubbed); the 2-D analytical solution integrates the 3-D point-source solution
over ~620 out-of-plane positions and takes about two minutes per receiver and component set.

    python tests/golden/make_analytic_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/DAS_Waveform_Modeling/src"


def main():
    for name in ("matplotlib", "matplotlib.animation", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    import elasticSolver as es
    from analyticalSolution import AnalyticalSolution

    # C1: 201 x 201, dx = dz = 10 m, ndamp = 40, dt = 1 ms, f0 = 10 Hz, source at the centre;
    # receivers 500 m to the left (the survey's point) and at (-400, -300) m so that both components are excited
    nx = nz = 201
    dx = dz = 10.0
    nt, dt, f0 = 601, 1.0e-3, 10.0
    vp0, rho0 = 4000.0, 2500.0
    vp = np.full((nx, nz), vp0)
    src = np.array([[100 * dx, 100 * dz]])
    rec = np.array([[50 * dx, 100 * dz], [60 * dx, 70 * dz]])
    sens = np.array([[1.0, 0, 0, 0, 0, 0], [0.5, 0, 0.3, 0, 0, 0.2]])
    kw = dict(nx=nx, nz=nz, ndamp=40, dx=dx, dz=dz, dt=dt, nt=nt, f0=f0, vp=vp, vs=vp / np.sqrt(3.0),
              rho=np.full((nx, nz), rho0), src_coord=src, das_coord=rec, geo_coord=rec, das_sensitivity=sens)
    solu = es.elasticSolver(**kw).forward_it(0, False)
    store = {"in_" + k: np.asarray(v) for k, v in kw.items()}
    for k in ("vx", "vz", "pr", "ett", "exx", "ezz", "exz"):
        store["numba_" + k] = solu[k]
    M = np.eye(3)
    for i in range(rec.shape[0]):
        x, z = abs(rec[i, 0] - src[0, 0]), abs(rec[i, 1] - src[0, 1])
        U = AnalyticalSolution(vp0, vp0 / np.sqrt(3.0), rho0, x, 0.0, z, 0.0, (nt - 1) * dt, dt, f0, 1e16, M,
                               dim='2D', comp='displacement', verbose=False)
        S = AnalyticalSolution(vp0, vp0 / np.sqrt(3.0), rho0, x, 0.0, z, 0.0, (nt - 1) * dt, dt, f0, 1e16, M,
                               dim='2D', comp='strain', verbose=False)
        for k in ("Ux", "Uz"):
            store["ana%d_%s" % (i, k)] = U[k][:nt]
        for k in ("Exx", "Ezz", "Exz"):
            store["ana%d_%s" % (i, k)] = S[k][:nt]
    path = os.path.join(HERE, "analytic_c1.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
