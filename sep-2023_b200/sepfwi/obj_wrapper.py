"""scipy.optimize front-end of an FWI module, same interface as DAS_Waveform_Inversion/Ops/FWI/obj_wrapper.py:10-97
(`PyTorchObjective(obj, loss)` with `.x0`, `.bounds`, `.fun(x)`, `.jac(x)`), as the reference drivers use it
(notebooks/Main-001-FWI-Anomaly-Vp-Vs-Den.py:118-168):

    obj = PyTorchObjective(fwi, lambda: fwi(Shot_ids, ngpu=ngpu))
    optimize.minimize(obj.fun, obj.x0, method='L-BFGS-B', jac=obj.jac, bounds=obj.bounds, ...)

scipy works on one flat float64 vector; this class scatters it into the module's parameters (which may live on
the GPU), evaluates misfit + gradient ONCE per new x and hands both back.  Differences from the reference, all
internal: parameters are updated in place instead of through load_state_dict (no buffer round trip), and the
"same x as last time" test is exact-array equality first, the reference's 1e-8 max-norm test second.
"""
from collections import OrderedDict

import numpy as np
import torch


class PyTorchObjective(object):
    def __init__(self, obj, loss):
        self.obj = obj            # torch module with the unknowns as nn.Parameters and a `Bounds` dict
        self.loss = loss          # callable -> scalar misfit tensor
        self._params = OrderedDict(obj.named_parameters())
        self.param_shapes = OrderedDict((n, tuple(p.shape)) for n, p in self._params.items())
        self.x0 = np.concatenate([p.detach().cpu().numpy().ravel() for p in self._params.values()]).astype(np.float64)
        self.bounds = self.pack_bounds() if getattr(obj, 'Bounds', {}) else None
        self.nfev = 0
        self.cached_x = None
        self.f = None
        self._jac = None

    # ---- packing ---------------------------------------------------------------------------
    def unpack_parameters(self, x):
        """flat vector -> OrderedDict name -> torch tensor of the parameter's shape (float64, CPU)."""
        out, i = OrderedDict(), 0
        for n, shp in self.param_shapes.items():
            size = int(np.prod(shp)) if len(shp) else 1
            out[n] = torch.from_numpy(np.asarray(x[i:i + size], dtype=np.float64).reshape(shp))
            i += size
        return out

    def pack_grads(self):
        return np.concatenate([p.grad.detach().cpu().numpy().ravel() for p in self._params.values()]).astype(np.float64)

    def pack_bounds(self):
        """L-BFGS-B box: obj.Bounds[name] = (lower, upper), arrays of the parameter's shape or scalars."""
        from scipy import optimize
        lo, hi = [], []
        for n, shp in self.param_shapes.items():
            size = int(np.prod(shp)) if len(shp) else 1
            if n in self.obj.Bounds:
                l, u = self.obj.Bounds[n][0], self.obj.Bounds[n][1]
                l = l.detach().cpu().numpy() if torch.is_tensor(l) else np.asarray(l)
                u = u.detach().cpu().numpy() if torch.is_tensor(u) else np.asarray(u)
                lo.append(np.broadcast_to(l, shp).ravel() if l.ndim else np.full(size, float(l)))
                hi.append(np.broadcast_to(u, shp).ravel() if u.ndim else np.full(size, float(u)))
            else:
                lo.append(np.full(size, -np.inf)); hi.append(np.full(size, np.inf))
        return optimize.Bounds(np.concatenate(lo).astype(np.float64), np.concatenate(hi).astype(np.float64))

    # ---- evaluation ------------------------------------------------------------------------
    def is_new(self, x):
        if self.cached_x is None:
            return True
        x = np.asarray(x)
        return not np.array_equal(x, self.cached_x) and float(np.abs(x - self.cached_x).max()) > 1e-8

    def cache(self, x):
        with torch.no_grad():
            for n, t in self.unpack_parameters(x).items():
                p = self._params[n]
                p.copy_(t.to(device=p.device, dtype=p.dtype))
        self.cached_x = np.array(x, dtype=np.float64, copy=True)
        self.obj.zero_grad()
        f = self.loss()
        self.f = float(f.item())
        f.backward()
        self._jac = self.pack_grads()
        self.nfev += 1

    def fun(self, x):
        if self.is_new(x):
            self.cache(x)
        return self.f

    def jac(self, x):
        if self.is_new(x):
            self.cache(x)
        return self._jac
