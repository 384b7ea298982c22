"""Front-end of the reference's OWN Numba CPU propagator (DAS_Waveform_Modeling/src/elasticSolver.py), placed unmodified in the
git-ignored oracle/_ref/numba_ref/ by oracle/Makefile.  TEST / BENCH INFRASTRUCTURE ONLY: bench.py times it on the GPU box's
host cores as the CPU baseline north_star names (`numba_cpu` section); nothing in the product imports it.

matplotlib is absent from this image and only used by the reference's plot_wavefield: three empty stub modules stand in.
"""
import importlib.util
import os
import sys
import time
import types

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "numba_ref", "elasticSolver.py")
_mod = None


def available():
    if not os.path.exists(_PATH):
        return False
    try:
        import numba  # noqa: F401
        return True
    except ImportError:
        return False


def load():
    global _mod
    if _mod is None:
        for name in ("matplotlib", "matplotlib.animation", "matplotlib.pyplot"):
            sys.modules.setdefault(name, types.ModuleType(name))
        spec = importlib.util.spec_from_file_location("elasticSolver_ref", _PATH)
        _mod = importlib.util.module_from_spec(spec)
        sys.modules["elasticSolver_ref"] = _mod      # picklable methods for the reference's multiprocessing Pool
        spec.loader.exec_module(_mod)
    return _mod


def time_forward(nx, nz, ndamp, dx, dz, dt, nt, f0, vp, nshots):
    """elasticSolver(...).forward() -- the reference's Pool over shots (elasticSolver.py:156-182) -- for `nshots` identical
    shots of one model (vp (nx, nz); vs = vp/1.732, rho = 310 vp^0.25), one DAS channel per 10 cells.
    Returns (cell-updates/s over the PADDED grid, wall seconds); JIT compilation is excluded by a warm-up solve."""
    es = load()
    vp = np.ascontiguousarray(vp, np.float64)
    vs, rho = vp / 1.732, 310.0 * vp ** 0.25
    xs = (0.5 * nx + np.arange(nshots) % 5) * dx
    src = np.stack([xs, np.full(nshots, 2.0 * dz)], 1)
    das = np.stack([np.arange(10, nx - 10, 10) * dx, np.full(len(range(10, nx - 10, 10)), 0.5 * nz * dz)], 1)
    sens = np.tile(np.array([[1.0, 0, 0, 0, 0, 0]]), (das.shape[0], 1))
    warm = es.elasticSolver(nx, nz, ndamp, dx, dz, dt, 5, f0, vp, vs, rho, src[:1], das, das[:1], sens)
    warm.forward_it(0, False)                    # compiles the two jitted updates in THIS process; the Pool forks inherit them
    S = es.elasticSolver(nx, nz, ndamp, dx, dz, dt, nt, f0, vp, vs, rho, src, das, das[:1], sens)
    t0 = time.perf_counter()
    out = S.forward(False)
    wall = time.perf_counter() - t0
    assert len(out) == nshots
    cells = float((nx + 2 * ndamp) * (nz + 2 * ndamp))
    return cells * nt * nshots / wall, wall
