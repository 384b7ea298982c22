"""Propagator: thin Python owner of one libsepfwi handle (one GPU).

Mirrors what the reference builds inside every `cufd` call -- Parameter, Model, Cpml,
Bnd, Src_Rec (DAS_Waveform_Inversion/Ops/FWI/Src/libCUFD.cu:47-87) -- but as a persistent
object, so an inversion pays for allocation and CPML set-up once.

Arrays may be numpy arrays, CPU torch tensors (host path: the library copies) or CUDA
torch tensors (zero-copy: the library reads / writes them in place on the current stream).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Params, Shot, check, lib

_COMP = {"pr": _lib.T_PR, "vx": _lib.T_VX, "vz": _lib.T_VZ, "ett": _lib.T_ETT,
         "exx": _lib.T_EXX, "ezz": _lib.T_EZZ, "exz": _lib.T_EXZ}


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _as_f32_host(a):
    if _is_torch(a):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def _as_i32(a):
    if _is_torch(a):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.int32)


class ShotSpec(object):
    """One shot in padded-grid indices (survey index + nPml, Src_Rec.cu:87-115)."""

    def __init__(self, zs, xs, zrec, xrec, stf, src_rxz=1.0, weights=None, win_start=None, win_end=None, trace_weights=None,
                 src_weight=1.0):
        self.zs, self.xs = int(zs), int(xs)
        self.zrec, self.xrec = _as_i32(zrec), _as_i32(xrec)
        if self.zrec.shape != self.xrec.shape or self.zrec.ndim != 1:
            raise ValueError("zrec and xrec must be 1-D arrays of equal length")
        self.stf = _as_f32_host(stf)
        self.src_rxz = float(src_rxz)
        self.weights = None if weights is None else _as_f32_host(weights).reshape(-1, 3)
        if self.weights is not None and self.weights.shape[0] != self.zrec.size:
            raise ValueError("weights must be (nrec, 3)")
        # data-side options (Propagator.set_data_options): per-trace windows [s], trace weights, source weight
        self.win_start = None if win_start is None else _as_f32_host(win_start).reshape(-1)
        self.win_end = None if win_end is None else _as_f32_host(win_end).reshape(-1)
        self.trace_weights = None if trace_weights is None else _as_f32_host(trace_weights).reshape(-1)
        for a in (self.win_start, self.win_end, self.trace_weights):
            if a is not None and a.size != self.zrec.size:
                raise ValueError("win_start / win_end / trace_weights need one entry per receiver")
        self.src_weight = float(src_weight)

    @property
    def nrec(self):
        return int(self.zrec.size)


class Propagator(object):
    def __init__(self, nz, nx, nPml, nPad, nSteps, dz, dx, dt, f0, fiber=_lib.FIBER_EXX,
                 flavour=_lib.FLAVOUR_CPML, max_batch=1, max_nrec=1, with_adjoint=False, device=0, kernels=0,
                 ref_race_compat=False):
        self.params = Params(int(nz), int(nx), int(nPml), int(nPad), int(nSteps), float(dz), float(dx), float(dt),
                             float(f0), int(fiber), int(flavour), int(max_batch), int(max_nrec),
                             1 if with_adjoint else 0, int(kernels), 1 if ref_race_compat else 0)
        self.device = int(device)
        self.nz, self.nx, self.nSteps = int(nz), int(nx), int(nSteps)
        self._h = C.c_void_p()
        check(lib().sepfwi_create(C.byref(self.params), self.device, C.byref(self._h)))

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().sepfwi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- helpers ----------------------------------------------------------------------------
    def _stream(self):
        try:
            import torch
            if torch.cuda.is_available():
                return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        except ImportError:
            pass
        return C.c_void_p(0)

    def _ptr(self, a, keep):
        """(pointer, is_device) of a float32 [..] array; host arrays are made contiguous and kept alive."""
        if _is_torch(a) and a.is_cuda:
            import torch
            if a.dtype != torch.float32 or not a.is_contiguous():
                a = a.detach().to(torch.float32).contiguous()
            keep.append(a)
            if a.device.index != self.device:
                raise ValueError("tensor lives on cuda:%s, propagator on cuda:%d" % (a.device.index, self.device))
            return a.data_ptr(), True
        h = _as_f32_host(a)
        keep.append(h)
        return h.ctypes.data, False

    # -- model --------------------------------------------------------------------------------
    def set_model(self, lam, mu, rho):
        """lam, mu [MPa] (sponge flavour: Pa), rho [kg/m^3]; (nz, nx) row-major float32."""
        keep = []
        ptrs = [self._ptr(a, keep) for a in (lam, mu, rho)]
        for a in keep:
            if tuple(a.shape) != (self.nz, self.nx):
                raise ValueError("model arrays must have shape (%d, %d), got %s" % (self.nz, self.nx, tuple(a.shape)))
        dev = [p[1] for p in ptrs]
        if any(dev) and not all(dev):
            raise ValueError("lam, mu, rho must all be on the host or all on the device")
        check(lib().sepfwi_set_model(self._h, ptrs[0][0], ptrs[1][0], ptrs[2][0],
                                     _lib.MEM_DEVICE if dev[0] else _lib.MEM_HOST, self._stream()))

    def set_data_options(self, if_win=False, win_ratio=0.0, filter=None, if_cross_misfit=False, if_src_update=False):
        """Data-side operators of later gradient() calls -- the switches of para_file.json (`if_win`, `filter` = [f0, f1, f2, f3] Hz,
        `if_cross_misfit`, `if_src_update`; Src/Parameter.cpp:139-176).  All off restores residual = obs - syn."""
        o = _lib.DataOptions()
        o.if_win, o.win_ratio = (1 if if_win else 0), float(win_ratio)
        if filter is not None:
            o.if_filter = 1
            for k in range(4):
                o.filter[k] = float(filter[k])
        o.if_cross_misfit, o.if_src_update = (1 if if_cross_misfit else 0), (1 if if_src_update else 0)
        check(lib().sepfwi_set_data_options(self._h, C.byref(o)))
        self._src_update = bool(if_src_update)

    @property
    def courant(self):
        c = C.c_float()
        check(lib().sepfwi_courant(self._h, C.byref(c)))
        return c.value

    def cpml(self, axis):
        n = (self.nz - self.params.nPad) if axis == 0 else self.nx
        out = np.zeros((6, n), np.float32)
        check(lib().sepfwi_get_cpml(self._h, axis, out.ctypes.data))
        return dict(zip(["K", "a", "b", "K_half", "a_half", "b_half"], out))

    @property
    def launches(self):
        return int(lib().sepfwi_launch_count(self._h))

    @property
    def resident_launches(self):
        """Cooperative launches of the shared-memory-resident forward loop so far (0: the streaming kernels ran)."""
        return int(lib().sepfwi_resident_launches(self._h))

    def bytes_per_slot(self):
        """Device bytes per unit of max_batch for this handle's parameters."""
        return int(lib().sepfwi_bytes_per_slot(C.byref(self.params)))

    def last_timing(self):
        f, b = C.c_float(), C.c_float()
        check(lib().sepfwi_last_timing(self._h, C.byref(f), C.byref(b)))
        return f.value, b.value

    def set_profile(self, nsteps):
        """Bracket every launch of the first `nsteps` steps of each time loop with CUDA events."""
        check(lib().sepfwi_set_profile(self._h, int(nsteps)))

    def profile(self):
        """{kernel name: (total ms, launches)} accumulated since set_profile."""
        ms = (C.c_double * _lib.NKERNEL)()
        n = (C.c_longlong * _lib.NKERNEL)()
        check(lib().sepfwi_get_profile(self._h, ms, n))
        return {lib().sepfwi_kernel_name(k).decode(): (ms[k], int(n[k])) for k in range(_lib.NKERNEL) if n[k] > 0}

    def ring_len(self):
        return int(lib().sepfwi_ring_len(C.byref(self.params)))

    def ring_save(self, field):
        f = _as_f32_host(field)
        out = np.zeros(self.ring_len(), np.float32)
        check(lib().sepfwi_ring_save(self._h, f.ctypes.data, out.ctypes.data))
        return out

    def ring_restore(self, field, bnd):
        f = _as_f32_host(field).copy()
        b = _as_f32_host(bnd)
        check(lib().sepfwi_ring_restore(self._h, f.ctypes.data, b.ctypes.data))
        return f

    # -- shots ----------------------------------------------------------------------------------
    def _shot_array(self, shots, keep):
        arr = (Shot * len(shots))()
        for k, s in enumerate(shots):
            c = arr[k]
            c.zs, c.xs, c.nrec = s.zs, s.xs, s.nrec
            c.zrec, c.xrec, c.stf = s.zrec.ctypes.data, s.xrec.ctypes.data, s.stf.ctypes.data
            if s.stf.size != self.nSteps:
                raise ValueError("stf must have nSteps=%d samples, got %d" % (self.nSteps, s.stf.size))
            c.src_rxz = s.src_rxz
            if s.weights is not None:
                c.weights = s.weights.ctypes.data
            for name in ("win_start", "win_end", "trace_weights"):
                a = getattr(s, name, None)
                if a is not None:
                    setattr(c, name, a.ctypes.data)
            c.src_weight = getattr(s, "src_weight", 1.0)
            keep.append(s)
        return arr

    def forward(self, shots, comps=("pr", "vx", "vz", "ett"), device_out=False, out=None):
        """Forward modelling.  Returns a list (one per shot) of {component: [nrec, nSteps] array}.
        device_out=True returns CUDA torch tensors (no device->host copy).  `out` may supply
        preallocated (e.g. pinned) host arrays: list of dicts like the return value."""
        keep = []
        arr = self._shot_array(shots, keep)
        res = []
        for k, s in enumerate(shots):
            d = {}
            for name in comps:
                if out is not None:
                    buf = out[k][name]
                    ptr = buf.data_ptr() if _is_torch(buf) else buf.ctypes.data
                elif device_out:
                    import torch
                    buf = torch.empty((s.nrec, self.nSteps), dtype=torch.float32, device="cuda:%d" % self.device)
                    ptr = buf.data_ptr()
                else:
                    buf = np.empty((s.nrec, self.nSteps), np.float32)
                    ptr = buf.ctypes.data
                arr[k].out[_COMP[name]] = ptr
                d[name] = buf
            res.append(d)
        check(lib().sepfwi_forward(self._h, len(shots), arr, _lib.MEM_DEVICE if device_out else _lib.MEM_HOST,
                                   self._stream()))
        return res

    def forward_snapshots(self, shot, save_step, comps=("pr", "vx", "vz", "ett")):
        """Sponge flavour: forward modelling of one shot plus the interior of sxx, szz, vx, vz after every time step that is a
        multiple of `save_step` (elasticSolver.forward_it(isrc, True)).  Returns (traces dict, snapshots) with snapshots
        shaped [(nSteps-1)//save_step + 1, 4, nz - 2 nPml, nx - 2 nPml], fields in the order sxx, szz, vx, vz."""
        keep = []
        arr = self._shot_array([shot], keep)
        d = {}
        for name in comps:
            buf = np.empty((shot.nrec, self.nSteps), np.float32)
            arr[0].out[_COMP[name]] = buf.ctypes.data
            d[name] = buf
        p = self.params
        snaps = np.zeros(((self.nSteps - 1) // int(save_step) + 1, 4, p.nz - p.nPad - 2 * p.nPml, p.nx - 2 * p.nPml), np.float32)
        check(lib().sepfwi_forward_snapshots(self._h, arr, int(save_step), snaps.ctypes.data, _lib.MEM_HOST, self._stream()))
        return d, snaps

    def condition(self, shot, obs, syn):
        """The data-side chain alone (no propagation) on host traces [nrec, nSteps] of one shot.
        Returns dict(res, syn, misfit, src_updated)."""
        keep = []
        arr = self._shot_array([shot], keep)
        o, s = _as_f32_host(obs), _as_f32_host(syn)
        if o.shape != (shot.nrec, self.nSteps) or s.shape != o.shape:
            raise ValueError("obs and syn must have shape (%d, %d)" % (shot.nrec, self.nSteps))
        res, out = np.zeros_like(o), np.zeros_like(o)
        upd = np.zeros(self.nSteps, np.float32)
        arr[0].src_updated = upd.ctypes.data
        J = C.c_double(0.0)
        check(lib().sepfwi_condition(self._h, arr, o.ctypes.data, s.ctypes.data, res.ctypes.data, out.ctypes.data, C.byref(J),
                                     _lib.MEM_HOST, self._stream()))
        return dict(res=res, syn=out, misfit=J.value, src_updated=upd)

    def gradient(self, shots, obs, with_adj=True, device=False, want_syn=False, grad_out=None):
        """Misfit and gradient of `shots` against observed DAS data `obs` (list of [nrec, nSteps]).
        device=True: obs are CUDA tensors and the gradients are returned as CUDA tensors.
        grad_out: optional (glam, gmu, grho) float32 (nz, nx) contiguous CUDA tensors (device=True) the library writes
        into -- e.g. views of a persistent packed all-reduce buffer (sepfwi.dist.PackedGradients).
        Returns dict(misfit, misfit64, glam, gmu, grho, gstf[list], syn[list])."""
        keep = []
        arr = self._shot_array(shots, keep)
        gstf, syn, src_upd = [], [], []
        for k, s in enumerate(shots):
            p, is_dev = self._ptr(obs[k], keep)
            if is_dev != bool(device):
                raise ValueError("obs must live in the space selected by `device`")
            if tuple(keep[-1].shape) != (s.nrec, self.nSteps):
                raise ValueError("obs[%d] must have shape (%d, %d)" % (k, s.nrec, self.nSteps))
            arr[k].obs_ett = p
            if with_adj:
                g = np.zeros(self.nSteps, np.float32)
                arr[k].gstf = g.ctypes.data
                gstf.append(g)
            if getattr(self, "_src_update", False):
                u = np.zeros(self.nSteps, np.float32)
                arr[k].src_updated = u.ctypes.data
                src_upd.append(u)
            if want_syn:
                if device:
                    import torch
                    b = torch.empty((s.nrec, self.nSteps), dtype=torch.float32, device="cuda:%d" % self.device)
                    arr[k].out[_lib.T_ETT] = b.data_ptr()
                else:
                    b = np.empty((s.nrec, self.nSteps), np.float32)
                    arr[k].out[_lib.T_ETT] = b.ctypes.data
                syn.append(b)
        misfit = C.c_float(0.0)
        g3 = [None, None, None]
        ptrs = [None, None, None]
        if with_adj:
            for k in range(3):
                if device and grad_out is not None:
                    g = grad_out[k]
                    if (not g.is_cuda or g.device.index != self.device or not g.is_contiguous()
                            or tuple(g.shape) != (self.nz, self.nx) or str(g.dtype) != "torch.float32"):
                        raise ValueError("grad_out[%d] must be a contiguous float32 (nz, nx) tensor on cuda:%d" % (k, self.device))
                    g3[k] = g
                    ptrs[k] = g.data_ptr()
                elif device:
                    import torch
                    g3[k] = torch.empty((self.nz, self.nx), dtype=torch.float32, device="cuda:%d" % self.device)
                    ptrs[k] = g3[k].data_ptr()
                else:
                    g3[k] = np.empty((self.nz, self.nx), np.float32)
                    ptrs[k] = g3[k].ctypes.data
        check(lib().sepfwi_gradient(self._h, len(shots), arr, 1 if with_adj else 0, C.byref(misfit),
                                    ptrs[0], ptrs[1], ptrs[2], _lib.MEM_DEVICE if device else _lib.MEM_HOST,
                                    self._stream()))
        m64 = C.c_double(0.0)
        check(lib().sepfwi_last_misfit(self._h, C.byref(m64)))
        return dict(misfit=misfit.value, misfit64=m64.value, glam=g3[0], gmu=g3[1], grho=g3[2], gstf=gstf, syn=syn,
                    src_updated=src_upd)
