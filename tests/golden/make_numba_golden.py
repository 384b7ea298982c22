"""Generate golden vectors from the reference's OWN Numba solver.

Runs only in the build container (needs /root/reference).  Imports
DAS_Waveform_Modeling/src/elasticSolver.py unmodified -- matplotlib is absent in
this image and is only used by plot_wavefield, so three empty stub modules are
placed in sys.modules first -- and stores inputs + seismograms of
elasticSolver(...).forward_it(0, False) as small .npz fixtures.

    python tests/golden/make_numba_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/DAS_Waveform_Modeling/src"


def import_reference_solver():
    for name in ("matplotlib", "matplotlib.animation", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    import elasticSolver as es  # noqa: E402
    return es


def cases():
    """name -> kwargs of elasticSolver.__init__ (arrays are (nx, nz))."""
    out = {}
    # homogeneous, the reference benchmark's medium (000-Solver-Benchmark.ipynb cell 4) on a small grid
    nx, nz = 61, 51
    vp = np.full((nx, nz), 4000.0)
    out["numba_homog"] = dict(
        nx=nx, nz=nz, ndamp=20, dx=10.0, dz=10.0, dt=1e-3, nt=300, f0=15.0,
        vp=vp, vs=vp / np.sqrt(3.0), rho=np.full((nx, nz), 2500.0),
        src_coord=np.array([[300.0, 250.0]]),
        das_coord=np.array([[100.0, 100.0], [200.0, 120.0], [450.0, 400.0], [310.0, 250.0]]),
        geo_coord=np.array([[100.0, 100.0], [500.0, 50.0]]),
        das_sensitivity=np.array([[1, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 1], [0.5, 0.3, 0, 0, 0, 0.2],
                                  [0.7, -0.4, 0, 0, 0, 0.1]], dtype=np.float64))
    # heterogeneous: layers + random perturbation, rho = 310 vp^0.25
    rng = np.random.default_rng(2023)
    nx, nz = 80, 60
    vp = np.empty((nx, nz))
    for j in range(nz):
        vp[:, j] = 2000.0 + 1500.0 * (j // 15) / 3.0
    vp *= 1.0 + 0.05 * rng.standard_normal((nx, nz))
    out["numba_hetero"] = dict(
        nx=nx, nz=nz, ndamp=15, dx=8.0, dz=10.0, dt=8e-4, nt=400, f0=12.0,
        vp=vp, vs=vp / 1.732, rho=310.0 * vp ** 0.25,
        src_coord=np.array([[320.0, 30.0], [100.0, 300.0]]),
        das_coord=np.stack([np.arange(40.0, 600.0, 80.0), np.full(7, 200.0)], axis=1),
        geo_coord=np.array([[64.0, 500.0], [560.0, 20.0], [320.0, 30.0]]),
        das_sensitivity=np.tile(np.array([[1.0, 0.25, 0, 0, 0, 0.5]]), (7, 1)))
    # small heterogeneous case stored WITH wavefield snapshots (forward_it(isrc, True): every save_step = 10 steps)
    nx, nz = 50, 40
    vp = 2200.0 + 900.0 * (np.arange(nz)[None, :] / nz) + 60.0 * rng.standard_normal((nx, nz))
    out["numba_wavefield"] = dict(
        nx=nx, nz=nz, ndamp=10, dx=10.0, dz=10.0, dt=1e-3, nt=95, f0=20.0,
        vp=vp, vs=vp / 1.732, rho=310.0 * vp ** 0.25,
        src_coord=np.array([[250.0, 120.0]]),
        das_coord=np.array([[100.0, 100.0], [300.0, 250.0]]),
        geo_coord=np.array([[400.0, 300.0]]),
        das_sensitivity=np.array([[1.0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 1.0]]))
    return out


def main():
    es = import_reference_solver()
    only = sys.argv[1:]
    for name, kw in cases().items():
        if only and name not in only:
            continue
        solver = es.elasticSolver(**kw)
        nshot = kw["src_coord"].shape[0]
        store = {"in_" + k: np.asarray(v) for k, v in kw.items()}
        for isrc in range(nshot):
            solu = solver.forward_it(isrc, name == "numba_wavefield")
            for k in ("vx", "vz", "pr", "ett", "exx", "ezz", "exz"):
                store["out%d_%s" % (isrc, k)] = solu[k]
            if name == "numba_wavefield":
                for k in ("sxx_wavefield", "szz_wavefield", "vx_wavefield", "vz_wavefield"):
                    store["out%d_%s" % (isrc, k)] = solu[k].astype(np.float32)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **store)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
