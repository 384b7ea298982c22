"""Seeded synthetic problems shared by the oracle tests, the GPU parity tests, the golden-vector
generators and bench.py.  Everything is float32 and defined on the PADDED grid the TorchFWI op
receives: (nz_pad, nx_pad) row-major, lambda/mu in MPa, rho in kg/m^3.
"""
import numpy as np


def pad_rule(nz, nx, nPml):
    nPad = int(32 - np.mod(nz + 2 * nPml, 32))        # notebooks/Main-001-...py:35
    return nz + 2 * nPml + nPad, nx + 2 * nPml, nPad


def ricker(f0, nSteps, dt, amp=1.0e7):
    """sourceGene, Ops/FWI/fwi_utils.py:127-140."""
    e = np.pi * np.pi * f0 * f0
    t = dt * np.arange(nSteps) - 1.2 / f0
    return ((1 - 2 * e * t ** 2) * np.exp(-e * t ** 2) * amp).astype(np.float32)


def layered_vp(nz, nx, vtop, vbot, nlayers, rng=None, nlens=0, lens_amp=0.1, sigma=(3, 12)):
    """Flat layers vtop -> vbot plus optional random gaussian lenses."""
    z = np.arange(nz)
    layer = np.minimum((z * nlayers) // nz, nlayers - 1)
    vp = (vtop + (vbot - vtop) * layer / max(nlayers - 1, 1))[:, None] * np.ones((1, nx))
    if rng is not None and nlens > 0:
        zz, xx = np.meshgrid(np.arange(nz), np.arange(nx), indexing="ij")
        for _ in range(nlens):
            cz, cx = rng.uniform(0, nz), rng.uniform(0, nx)
            s = rng.uniform(*sigma)
            a = rng.uniform(-lens_amp, lens_amp)
            vp = vp * (1.0 + a * np.exp(-((zz - cz) ** 2 + (xx - cx) ** 2) / (2 * s * s)))
    return vp


def smooth(a, n):
    """Separable box smoothing repeated 3x (approximately gaussian), edge-replicated."""
    out = a.astype(np.float64)
    k = np.ones(2 * n + 1) / (2 * n + 1)
    for _ in range(3):
        p = np.pad(out, ((n, n), (0, 0)), mode="edge")
        out = np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), 0, p)
        p = np.pad(out, ((0, 0), (n, n)), mode="edge")
        out = np.apply_along_axis(lambda v: np.convolve(v, k, mode="valid"), 1, p)
    return out


def lame_from_vp(vp_pad):
    """vs = vp/1.732, rho = 310 vp^0.25 (MOD/notebooks/000-Solver-Benchmark.ipynb cell 8);
    lambda, mu in MPa like FWI.forward (Ops/FWI/FWI_ops.py:124-125)."""
    vp = vp_pad.astype(np.float32)
    vs = (vp / np.float32(1.732)).astype(np.float32)
    rho = (310.0 * vp.astype(np.float64) ** 0.25).astype(np.float32)
    lam = ((vp ** 2 - 2.0 * vs ** 2) * rho / 1e6).astype(np.float32)
    mu = (vs ** 2 * rho / 1e6).astype(np.float32)
    return lam, mu, rho


def pad_model(a, nPml, nPad):
    return np.pad(a, ((nPml, nPml + nPad), (nPml, nPml)), mode="edge")


class Problem(object):
    """A complete FWI problem: padded true / start models, survey (interior indices), stf."""

    def __init__(self, name, nz, nx, nPml, dz, dx, dt, nSteps, f0, vp_true, vp_start, z_src, x_src, z_rec, x_rec,
                 fiber=0):
        self.name = name
        self.nz_orig, self.nx_orig, self.nPml = nz, nx, nPml
        self.nz, self.nx, self.nPad = pad_rule(nz, nx, nPml)
        self.dz, self.dx, self.dt, self.nSteps, self.f0 = dz, dx, dt, nSteps, f0
        self.true = lame_from_vp(pad_model(vp_true, nPml, self.nPad))
        self.start = lame_from_vp(pad_model(vp_start, nPml, self.nPad))
        self.z_src, self.x_src = np.asarray(z_src, np.int32), np.asarray(x_src, np.int32)
        self.z_rec, self.x_rec = np.asarray(z_rec, np.int32), np.asarray(x_rec, np.int32)
        self.nshots = len(self.x_src)
        self.stf = np.tile(ricker(f0, nSteps, dt)[None, :], (self.nshots, 1))
        self.fiber = fiber

    def survey(self):
        """shot id -> (zs, xs, zrec, xrec), interior indices (as in survey_file.json)."""
        return {i: (int(self.z_src[i]), int(self.x_src[i]), self.z_rec, self.x_rec) for i in range(self.nshots)}


def tiny(fiber=0):
    """40 x 56 interior, nPml 8, 2 shots, 12 receivers, 260 steps: finishes in < 1 s in the oracle."""
    rng = np.random.default_rng(2023)
    nz, nx = 40, 56
    vt = layered_vp(nz, nx, 1800.0, 3200.0, 4, rng, nlens=6, lens_amp=0.08, sigma=(2, 8))
    vs = smooth(vt, 4)
    if fiber == 0:
        z_rec, x_rec = np.full(12, 30), np.arange(6, 54, 4)
    else:
        z_rec, x_rec = np.arange(8, 32, 2), np.full(12, 40)
    return Problem("tiny" + ("_ezz" if fiber else ""), nz, nx, 8, 10.0, 10.0, 1.0e-3, 260, 18.0, vt, vs,
                   [2, 3], [14, 40], z_rec, x_rec, fiber)


def small(adjacent=False):
    """64 x 96 interior, nPml 16, 3 shots, 40 receivers, 420 steps.
    adjacent=False: receivers two cells apart, no two receivers touch the same cell.
    adjacent=True ("small_adj"): 40 ADJACENT receivers.  Receivers 31 and 32 then sit in different
    32-thread blocks of the reference's res_injection_exx launch and both update cell x_31 with a plain
    load-add-store (Src/utilities.cu:613-614): the reference loses one update there on every time step,
    which changes its gradient by ~15%.  The oracle reproduces that with par.race = 2."""
    rng = np.random.default_rng(7)
    nz, nx = 64, 96
    vt = layered_vp(nz, nx, 1700.0, 3600.0, 5, rng, nlens=10, lens_amp=0.1, sigma=(3, 10))
    vs = smooth(vt, 6)
    x_rec = np.arange(28, 68) if adjacent else np.arange(8, 88, 2)
    return Problem("small_adj" if adjacent else "small", nz, nx, 16, 10.0, 10.0, 1.0e-3, 420, 15.0, vt, vs,
                   [1, 1, 2], [10, 48, 86], np.full(40, 50), x_rec, 0)


def small_adj():
    return small(adjacent=True)


def medium(fiber=0, nx=291):
    """150 x 291 interior (padded width 307: odd, three 120-column strips), nPml 8, 2 shots, 170 ADJACENT
    receivers crossing the strip seams at x = 120 and 240, 180 steps.  Big enough that the streaming kernels run
    their branch-free interior variant next to the CPML variant; too big for the oracle in the CPU suite, so it
    is checked on the GPU against the baseline kernels (which the small problems pin to the oracle)."""
    rng = np.random.default_rng(11)
    nz = 150
    vt = layered_vp(nz, nx, 1800.0, 3400.0, 5, rng, nlens=14, lens_amp=0.1, sigma=(4, 16))
    vs = smooth(vt, 6)
    if fiber == 0:
        z_rec, x_rec = np.full(170, 60), np.arange(70, 240)
    else:
        z_rec, x_rec = np.arange(5, 145), np.full(140, 113)       # padded column 121: first owned quad of strip 1
    return Problem("medium" + ("_ezz" if fiber else ""), nz, nx, 8, 10.0, 10.0, 1.0e-3, 180, 16.0, vt, vs,
                   [2, 70], [30, 231], z_rec, x_rec, fiber)


def reference_test(nSteps=1501, nshots=19):
    """The reference's own test problem (notebooks/Main-001-...py:28-72):
    101 x 201, dx=dz=20, dt=2 ms, nPml 32, 19 shots at z=1, 181 receivers at z=95,
    homogeneous start, Vp +80 box anomaly in the true model."""
    nz, nx = 101, 201
    vs_ = np.full((nz, nx), 4000.0)
    vt = vs_.copy()
    vt[42:58, 42:58] += 80.0
    x_src = np.arange(10, 200, 10)[:nshots]
    return Problem("reftest", nz, nx, 32, 20.0, 20.0, 2.0e-3, nSteps, 10.0, vt, vs_,
                   np.full(len(x_src), 1), x_src, np.full(181, 95), np.arange(10, 191), 0)


# ---- data-side operators (tests/golden/make_dataops_golden.py, tests/test_oracle.py, tests/test_gpu_dataops.py) ----------------
DATAOPS_CASES = {
    "plain": dict(),                                  # only the window-less end tapers of libCUFD.cu:363-367 + L2 residual
    "win": dict(if_win=True),
    "filt": dict(filter=True),
    "cross": dict(if_cross_misfit=True),
    "srcupd": dict(if_src_update=True),
    "win_filt": dict(if_win=True, filter=True),
    "win_filt_cross": dict(if_win=True, filter=True, if_cross_misfit=True),
    "all_l2": dict(if_win=True, filter=True, if_src_update=True),
}


def dataops_traces(nrec, nt, dt, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(nt) * dt
    out = np.zeros((nrec, nt))
    for r in range(nrec):
        for f in rng.uniform(4.0, 60.0, 5):
            out[r] += rng.normal() * np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28))
        out[r] *= np.exp(-((t - 0.45 * nt * dt) / (0.2 * nt * dt)) ** 2) * 1e2
    return out.astype(np.float32)


def dataops_case(nrec=9, nt=301, dt=2.0e-3):
    """Seeded inputs of the data-side golden vectors: two trace gathers, per-trace windows and weights, a tapered Ricker source."""
    obs, cal = dataops_traces(nrec, nt, dt, 1), dataops_traces(nrec, nt, dt, 2)
    cal = (0.6 * obs + 0.4 * cal).astype(np.float32)
    ws = (np.linspace(0.05, 0.2, nrec) * nt * dt).astype(np.float32)
    we = (np.linspace(0.7, 0.95, nrec) * nt * dt).astype(np.float32)
    wt = np.linspace(0.5, 1.5, nrec).astype(np.float32)
    src = ricker(12.0, nt, dt, amp=1.0).astype(np.float32)
    return dict(obs=obs, cal=cal, src=src, dt=dt, win_start=ws, win_end=we, weights=wt, src_weight=0.75,
                filt=np.array([8.0, 15.0, 30.0, 45.0], np.float32))
