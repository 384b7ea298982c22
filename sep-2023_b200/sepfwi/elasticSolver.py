"""GPU-backed `elasticSolver` with the constructor and call signatures of
DAS_Waveform_Modeling/src/elasticSolver.py (class elasticSolver :31-305):

    elasticSolver(nx, nz, ndamp, dx, dz, dt, nt, f0, vp, vs, rho, src_coord, das_coord, geo_coord, das_sensitivity)
    .set_model(vp, vs, rho)   .forward(save_wavefield=False) -> [dict per shot]   .forward_it(isrc, save_wavefield)

Same conventions: arrays are (nx, nz); coordinates in metres as (n, 2) = [x, z], snapped to the nearest grid
point; the model is edge-padded by `ndamp` cells carrying the sin^2 sponge; unit-amplitude Ricker source added
to sxx and szz as stf*dt/2; dict keys t, vx, vz, pr, ett, exx, ezz, exz with (n, nt) float64 arrays.
The time stepping runs in fp32 on the GPU (sponge flavour of libsepfwi: velocity -> stress -> source ->
record, swapped density averages and all -- see SURVEY.md A.7); the reference is a Numba fp64 CPU loop.
All shots are propagated concurrently on the device instead of a multiprocessing pool.
"""
import numpy as np

from . import _lib
from .engine import Propagator, ShotSpec


class elasticSolver(object):

    def __init__(self, nx, nz, ndamp, dx, dz, dt, nt, f0, vp, vs, rho, src_coord, das_coord, geo_coord,
                 das_sensitivity, device=0, max_batch=None):
        self.nx, self.nz, self.ndamp = nx + 2 * ndamp, nz + 2 * ndamp, ndamp
        self.dx, self.dz, self.dt, self.nt, self.f0 = dx, dz, dt, nt, f0
        self.src_coord, self.das_coord, self.geo_coord = (np.asarray(c, np.float64) for c in (src_coord, das_coord, geo_coord))
        self.das_sensitivity = np.asarray(das_sensitivity, np.float64)
        self.x = np.arange(0, nx * dx, dx)
        self.z = np.arange(0, nz * dz, dz)
        self.t = np.arange(0, nt * dt, dt)
        self.save_step = 10
        for c, what in ((self.src_coord, "src_coord"), (self.das_coord, "das_coord"), (self.geo_coord, "geo_coord")):
            if c.ndim != 2 or c.shape[1] != 2:
                raise ValueError("The shape of %s should be (n, 2)" % what)
        grid = lambda c: np.array([np.round(c[:, 0] / dx).astype(int), np.round(c[:, 1] / dz).astype(int)]) + ndamp
        self.src_grid, self.das_grid, self.geo_grid = grid(self.src_coord), grid(self.das_coord), grid(self.geo_coord)
        self.src_num, self.das_num, self.geo_num = self.src_grid.shape[1], self.das_grid.shape[1], self.geo_grid.shape[1]
        for g, what in ((self.src_grid, "source"), (self.das_grid, "das receiver"), (self.geo_grid, "geo receiver")):
            if g.size and (g[0].max() > self.nx or g[1].max() > self.nz or g.min() < 0):
                raise ValueError("The %s coordinates should be within the model" % what)
        if self.das_sensitivity.shape != (self.das_num, 6):
            raise ValueError("The shape of das_sensitivity should be (nchannel, 6) with exx, exy, exz, eyy, eyz, and ezz")
        tt = self.t
        self.stf = (1.0 - 2.0 * np.pi ** 2 * f0 ** 2 * (tt - 1.2 / f0) ** 2) * np.exp(-np.pi ** 2 * f0 ** 2 * (tt - 1.2 / f0) ** 2)
        self._device = device
        nrec = self.geo_num + self.das_num
        batch = max_batch or max(1, min(self.src_num, 16))
        self._prop = Propagator(self.nz, self.nx, ndamp, 0, nt, dz, dx, dt, f0, flavour=_lib.FLAVOUR_SPONGE,
                                max_batch=batch, max_nrec=max(nrec, 1), device=device)
        self.set_model(vp, vs, rho)

    def set_model(self, vp, vs, rho):
        nx, nz = self.nx - 2 * self.ndamp, self.nz - 2 * self.ndamp
        vp, vs, rho = (np.asarray(a, np.float64) for a in (vp, vs, rho))
        if vp.shape != (nx, nz) or vs.shape != (nx, nz) or rho.shape != (nx, nz):
            raise ValueError("The shape of vp, vs and rho should be (nx, nz)")
        self.vp, self.vs, self.rho = (np.pad(a, self.ndamp, 'edge') for a in (vp, vs, rho))
        self.mu = self.rho * self.vs ** 2
        self.lam = self.rho * self.vp ** 2 - 2 * self.mu
        # the library is [z][x]; Pa for the sponge flavour
        self._prop.set_model(np.ascontiguousarray(self.lam.T, np.float32), np.ascontiguousarray(self.mu.T, np.float32),
                             np.ascontiguousarray(self.rho.T, np.float32))

    def _shot(self, isrc):
        zrec = np.concatenate([self.geo_grid[1], self.das_grid[1]]).astype(np.int32)
        xrec = np.concatenate([self.geo_grid[0], self.das_grid[0]]).astype(np.int32)
        w = np.zeros((zrec.size, 3), np.float32)
        s = self.das_sensitivity                      # ett = s0*exx + s3*ezz + s1*exz  (elasticSolver.py:276)
        w[self.geo_num:, 0], w[self.geo_num:, 1], w[self.geo_num:, 2] = s[:, 0], s[:, 3], s[:, 1]
        return ShotSpec(self.src_grid[1, isrc], self.src_grid[0, isrc], zrec, xrec, self.stf[:self.nt].astype(np.float32),
                        weights=w)

    def _solu(self, out):
        g = self.geo_num
        f64 = lambda a: np.asarray(a, np.float64)
        return {'t': self.t, 'vx': f64(out['vx'][:g]), 'vz': f64(out['vz'][:g]), 'pr': f64(out['pr'][:g]),
                'ett': f64(out['ett'][g:]), 'exx': f64(out['exx'][g:]), 'ezz': f64(out['ezz'][g:]), 'exz': f64(out['exz'][g:])}

    _COMPS = ("pr", "vx", "vz", "ett", "exx", "ezz", "exz")

    def forward(self, save_wavefield=False):
        if save_wavefield:
            return [self.forward_it(i, True) for i in range(self.src_num)]
        outs = self._prop.forward([self._shot(i) for i in range(self.src_num)], comps=self._COMPS)
        return [self._solu(o) for o in outs]

    def forward_it(self, isrc, save_wavefield=False):
        if not save_wavefield:
            return self._solu(self._prop.forward([self._shot(isrc)], comps=self._COMPS)[0])
        # snapshots every save_step steps of the interior, as (save_num, nx, nz) arrays (elasticSolver.py:231-237,279-303)
        out, snaps = self._prop.forward_snapshots(self._shot(isrc), self.save_step, comps=self._COMPS)
        solu = self._solu(out)
        save_num = self.nt // self.save_step + 1
        for k, name in enumerate(("sxx_wavefield", "szz_wavefield", "vx_wavefield", "vz_wavefield")):
            w = np.zeros((save_num,) + snaps.shape[2:][::-1], np.float64)
            w[:snaps.shape[0]] = np.transpose(snaps[:, k], (0, 2, 1))
            solu[name] = w
        return solu
