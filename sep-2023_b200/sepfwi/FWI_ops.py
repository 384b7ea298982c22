"""torch front-end with the names and call signatures of DAS_Waveform_Inversion/Ops/FWI/FWI_ops.py:

    FWIFunction(Lambda, Mu, Den, Stf, ngpu, Shot_ids, para_fname) -> misfit      (:46-63)
    FWI(Vp, Vs, Den, Stf, opt, Mask, Vp_bounds, Vs_bounds, Den_bounds)           (:66-127)
    FWI_obscalc(Vp, Vs, Den, Stf, para_fname)                                    (:131-141)
    FWI_Lame_Den / FWI_IP_IS_Den / FWI_Vp_Vs_IP / FWI_Vp_Vs_IS                   (:146-395)
    FWI_Rock_Physics_VRH / FWI_Rock_Physics_gassmann (PHI, CC, SW)               (:401-619)

Every parameterisation is the same module -- pad three (nz_orig, nx_orig) fields to the padded grid, blend
with the frozen reference model through `Mask`, map to (Lambda [MPa], Mu [MPa], Den) and call the op -- so it
is written once (`_ThreeParameterFWI`) and specialised by the map.  Unlike the reference, parameters and the
mask may live on the GPU; the op then runs without any host copy of the model or the gradients.
"""
import torch
import torch.nn as nn

from . import fwi_ops
from . import fwi_utils as ft


class FWIFunction(torch.autograd.Function):
    """forward() computes misfit AND gradients in one sweep and stashes the gradients; backward() returns them.
    `scale_by_grad_output=False` reproduces the reference, which ignores grad_misfit (FWI_ops.py:54-63)."""
    scale_by_grad_output = False

    @staticmethod
    def forward(ctx, Lambda, Mu, Den, Stf, ngpu, Shot_ids, para_fname):
        out = fwi_ops.backward(Lambda, Mu, Den, Stf, ngpu, Shot_ids, para_fname)
        ctx.grads = out[1:]
        return out[0].to(Lambda.device)

    @staticmethod
    def backward(ctx, grad_misfit):
        gl, gm, gd, gs = ctx.grads
        if FWIFunction.scale_by_grad_output:
            gl, gm, gd, gs = (g * grad_misfit.to(g.device) for g in (gl, gm, gd, gs))
        return gl, gm, gd, gs, None, None, None


class _ThreeParameterFWI(nn.Module):
    NAMES = ("A", "B", "C")

    def __init__(self, p0, p1, p2, Stf, opt, Mask=None, b0=None, b1=None, b2=None):
        super().__init__()
        self.nz, self.nx = opt['nz'], opt['nx']
        self.nz_orig, self.nx_orig = opt['nz_orig'], opt['nx_orig']
        self.nPml, self.nPad = opt['nPml'], opt['nPad']
        self.Bounds = {}
        pads = ft.padding(p0, p1, p2, self.nz_orig, self.nx_orig, self.nz, self.nx, self.nPml, self.nPad)
        for name, t, pad, bounds in zip(self.NAMES, (p0, p1, p2), pads, (b0, b1, b2)):
            self.register_buffer(name + '_ref', pad.clone().detach())
            if t.requires_grad:
                setattr(self, name, nn.Parameter(t))
                if bounds is not None:
                    self.Bounds[name] = bounds
            else:
                setattr(self, name, t)
        if Mask is None:
            Mask = torch.ones((self.nz + 2 * self.nPml + self.nPad, self.nx + 2 * self.nPml), dtype=torch.float32,
                              device=p0.device)
        self.Mask = Mask
        self.Stf = Stf
        self.para_fname = opt['para_fname']

    def _masked(self):
        fields = [getattr(self, n) for n in self.NAMES]
        pads = ft.padding(*fields, self.nz_orig, self.nx_orig, self.nz, self.nx, self.nPml, self.nPad)
        return [self.Mask * p + (1.0 - self.Mask) * getattr(self, n + '_ref') for n, p in zip(self.NAMES, pads)]

    @staticmethod
    def to_lame(a, b, c):
        raise NotImplementedError

    def forward(self, Shot_ids, ngpu=1):
        Lambda, Mu, Den = self.to_lame(*self._masked())
        return FWIFunction.apply(Lambda, Mu, Den, self.Stf, ngpu, Shot_ids, self.para_fname)


class FWI(_ThreeParameterFWI):
    """Vp, Vs [m/s], Den [kg/m^3]  (FWI_ops.py:116-127)."""
    NAMES = ("Vp", "Vs", "Den")

    def __init__(self, Vp, Vs, Den, Stf, opt, Mask=None, Vp_bounds=None, Vs_bounds=None, Den_bounds=None):
        super().__init__(Vp, Vs, Den, Stf, opt, Mask, Vp_bounds, Vs_bounds, Den_bounds)

    @staticmethod
    def to_lame(vp, vs, den):
        return (vp ** 2 - 2.0 * vs ** 2) * den / 1e6, vs ** 2 * den / 1e6, den


class FWI_Lame_Den(_ThreeParameterFWI):
    """Lambda, Mu [MPa], Den  (FWI_ops.py:195-204)."""
    NAMES = ("Lam", "Mu", "Den")

    def __init__(self, Lam, Mu, Den, Stf, opt, Mask=None, Lam_bounds=None, Mu_bounds=None, Den_bounds=None):
        super().__init__(Lam, Mu, Den, Stf, opt, Mask, Lam_bounds, Mu_bounds, Den_bounds)

    @staticmethod
    def to_lame(lam, mu, den):
        return lam, mu, den


class FWI_IP_IS_Den(_ThreeParameterFWI):
    """P and S impedances (already scaled so that IP^2/Den is in MPa, as in the reference) and Den (FWI_ops.py:256-267)."""
    NAMES = ("IP", "IS", "Den")

    def __init__(self, IP, IS, Den, Stf, opt, Mask=None, IP_bounds=None, IS_bounds=None, Den_bounds=None):
        super().__init__(IP, IS, Den, Stf, opt, Mask, IP_bounds, IS_bounds, Den_bounds)

    @staticmethod
    def to_lame(ip, is_, den):
        return (ip ** 2 - 2.0 * is_ ** 2) / den, is_ ** 2 / den, den


class FWI_Vp_Vs_IP(_ThreeParameterFWI):
    """Vp, Vs and P impedance  (FWI_ops.py:318-330)."""
    NAMES = ("Vp", "Vs", "IP")

    def __init__(self, Vp, Vs, IP, Stf, opt, Mask=None, Vp_bounds=None, Vs_bounds=None, IP_bounds=None):
        super().__init__(Vp, Vs, IP, Stf, opt, Mask, Vp_bounds, Vs_bounds, IP_bounds)

    @staticmethod
    def to_lame(vp, vs, ip):
        den = ip / vp
        mu = den * vs ** 2
        return ip * vp - 2.0 * mu, mu, den


class FWI_Vp_Vs_IS(_ThreeParameterFWI):
    """Vp, Vs and S impedance  (FWI_ops.py:381-395)."""
    NAMES = ("Vp", "Vs", "IS")

    def __init__(self, Vp, Vs, IS, Stf, opt, Mask=None, Vp_bounds=None, Vs_bounds=None, IS_bounds=None):
        super().__init__(Vp, Vs, IS, Stf, opt, Mask, Vp_bounds, Vs_bounds, IS_bounds)

    @staticmethod
    def to_lame(vp, vs, is_):
        den = is_ / vs
        return den * vp ** 2 - 2.0 * is_ * vs, is_ * vs, den


# Mineral / fluid constants of the reference's rock-physics maps (FWI_ops.py:459-471, 575-587): quartz, clay, water,
# hydrocarbon bulk moduli [Pa], quartz / clay shear moduli [Pa], densities [kg/m^3], consolidation parameter.
ROCK = dict(k_q=37.00e9, k_c=21.00e9, k_w=2.25e9, k_h=0.04e9, mu_q=44.00e9, mu_c=10.00e9,
            rho_q=2.65e3, rho_c=2.55e3, rho_w=1.00e3, rho_h=0.10e3, cs=20.0)


def _density(phi, cc, sw):
    """Volume average of fluid (water / hydrocarbon) and skeleton (clay / quartz) densities."""
    R = ROCK
    rho_f = R['rho_w'] * sw + R['rho_h'] * (1 - sw)
    rho_s = R['rho_c'] * cc + R['rho_q'] * (1 - cc)
    return rho_f * phi + rho_s * (1 - phi)


class FWI_Rock_Physics_VRH(_ThreeParameterFWI):
    """Porosity, clay content, water saturation -> Voigt-Reuss-Hill average moduli (FWI_ops.py:451-507):
    bulk modulus = mean of the Voigt (arithmetic) and Reuss (harmonic) averages over the four constituents,
    shear modulus = half the Voigt average of the solid part (the Reuss shear modulus of a fluid-bearing mix is 0)."""
    NAMES = ("PHI", "CC", "SW")

    def __init__(self, PHI, CC, SW, Stf, opt, Mask=None, PHI_bounds=None, CC_bounds=None, SW_bounds=None):
        super().__init__(PHI, CC, SW, Stf, opt, Mask, PHI_bounds, CC_bounds, SW_bounds)

    @staticmethod
    def to_lame(phi, cc, sw):
        R = ROCK
        k_voigt = (1 - phi) * (R['k_c'] * cc + R['k_q'] * (1 - cc)) + phi * (R['k_w'] * sw + R['k_h'] * (1 - sw))
        k_reuss = 1 / ((1 - phi) * (cc / R['k_c'] + (1 - cc) / R['k_q']) + phi * (sw / R['k_w'] + (1 - sw) / R['k_h']))
        k = 0.5 * (k_voigt + k_reuss)
        mu = 0.5 * ((1 - phi) * (R['mu_c'] * cc + R['mu_q'] * (1 - cc)) + 0)
        return (k - 2. / 3. * mu) / 1e6, mu / 1e6, _density(phi, cc, sw)


class FWI_Rock_Physics_gassmann(_ThreeParameterFWI):
    """Porosity, clay content, water saturation -> Gassmann fluid substitution on a consolidation-parameter dry frame
    (FWI_ops.py:567-619, after PyFWI).  The P modulus uses 0.75 mu_d exactly as the reference does (:606)."""
    NAMES = ("PHI", "CC", "SW")

    def __init__(self, PHI, CC, SW, Stf, opt, Mask=None, PHI_bounds=None, CC_bounds=None, SW_bounds=None):
        super().__init__(PHI, CC, SW, Stf, opt, Mask, PHI_bounds, CC_bounds, SW_bounds)

    @staticmethod
    def to_lame(phi, cc, sw):
        R = ROCK
        k_f = R['k_w'] * sw + R['k_h'] * (1 - sw)
        k_s = R['k_c'] * cc + R['k_q'] * (1 - cc)
        mu_s = R['mu_c'] * cc + R['mu_q'] * (1 - cc)
        k_d = k_s * ((1 - phi) / (1 + R['cs'] * phi))
        mu_d = mu_s * ((1 - phi) / (1 + 1.5 * R['cs'] * phi))
        delta = ((1 - phi) / phi) * (k_f / k_s) * (1 - (k_d / (k_s - k_s * phi)))
        k_u = (phi * k_d + (1 - (1 + phi) * (k_d / k_s)) * k_f) / (phi * (1 + delta))
        rho = _density(phi, cc, sw)
        vp = torch.sqrt((k_u + 0.75 * mu_d) / rho)
        vs = torch.sqrt(mu_d / rho)
        return rho * (vp ** 2 - 2 * vs ** 2) / 1e6, rho * vs ** 2 / 1e6, rho


class FWI_obscalc(nn.Module):
    """Observed-data generation: forward-models the shots and writes Shot_*.bin (FWI_ops.py:131-141)."""

    def __init__(self, Vp, Vs, Den, Stf, para_fname):
        super().__init__()
        self.Lambda = (Vp ** 2 - 2.0 * Vs ** 2) * Den / 1e6
        self.Mu = Vs ** 2 * Den / 1e6
        self.Den = Den
        self.Stf = Stf
        self.para_fname = para_fname

    def forward(self, Shot_ids, ngpu=1):
        fwi_ops.obscalc(self.Lambda, self.Mu, self.Den, self.Stf, ngpu, Shot_ids, self.para_fname)
