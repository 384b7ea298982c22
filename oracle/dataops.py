"""CPU restatement (numpy, float64 FFTs) of the reference's data-side operators -- TEST INFRASTRUCTURE ONLY.

These are the operators `para_file.json` parameterises (`if_win`, `filter`, `if_cross_misfit`, `if_src_update`) and that act
between the forward and the reverse-time loop of every shot.  In the reference their call sites are written out but
commented (DAS_Waveform_Inversion/Ops/FWI/Src/libCUFD.cu:353-457); the operators themselves are live code in
Src/utilities.cu.  Each function below cites the lines it follows; `condition()` strings them together in the order of the
commented call sites, applied to the DAS component (the only component that enters the objective, libCUFD.cu:427).

Pinned by execution: oracle/ref_dataops_shim.cu (compiled into oracle/_ref/libcufd_ref.so with the reference's sources) runs the
reference's own kernels in the order and with the launch configurations of the commented call sites; its outputs on seeded traces are
committed as tests/golden/dataops_ref.npz (tests/golden/make_dataops_golden.py, run on the B200) and tests/test_oracle.py holds this
file to them, next to self-checks of the formulas (band-pass of in-band / out-of-band sinusoids, cross-misfit value and gradient by
finite differences, Wiener update recovering a known filter).  Two quirks of the reference are NOT restated (both are launch-grid
slips, not intent): its padded scratch rows are zero-filled over the first nt columns only (utilities.cu:1133 -- the golden run
parks zeroed blocks in the allocator so that the rest is zero too), and its band-pass / coefficient kernels are launched over
ceil(nt / 32) * 32 columns, so the Nyquist bin nt of the 2 nt transform stays unfiltered when nt is a multiple of 32.
"""
import numpy as np

DIVCONST = 1e-9          # Src/utilities.h:24


def taper_amp(t, t0, t3, offset):
    """sin / cos end taper shared by cuda_window (utilities.cu:823-833, 868-878): amplitude, to be squared by the caller."""
    t1, t2 = t0 + offset, t3 - offset
    a = np.zeros_like(t)
    m = (t >= t0) & (t < t1)
    a[m] = np.sin(np.pi / 2.0 * (t[m] - t0) / (t1 - t0))
    a[(t >= t1) & (t < t2)] = 1.0
    m = (t >= t2) & (t < t3)
    a[m] = np.cos(np.pi / 2.0 * (t[m] - t2) / (t3 - t2))
    return a


def window_traces(data, dt, win_start, win_end, weights, src_weight, ratio):
    """cuda_window with per-trace windows and weights, utilities.cu:790-842.  data [nrec][nt] (copy returned)."""
    nrec, nt = data.shape
    t = np.float32(dt) * np.arange(nt, dtype=np.float32)
    tmax = np.float32(nt * dt)
    out = np.array(data, np.float64)
    for r in range(nrec):
        t0 = min(max(np.float32(win_start[r]), 0.0), tmax)
        t3 = min(max(np.float32(win_end[r]), 0.0), tmax)
        off = (t3 - t0) * ratio
        if off <= 0.0:
            continue                                  # "Window error 1": the trace is left untouched (:815-819)
        a = taper_amp(t.astype(np.float64), float(t0), float(t3), float(off))
        out[r] *= a * a * float(weights[r]) * float(src_weight)
    return out


def window_simple(data, dt, ratio):
    """cuda_window without windows, utilities.cu:844-884: taper of `ratio` x record length at both ends."""
    nt = data.shape[-1]
    t = dt * np.arange(nt, dtype=np.float64)
    t3 = nt * dt
    off = nt * dt * ratio
    if 2.0 * off >= t3:
        return np.array(data, np.float64)
    a = taper_amp(t, 0.0, t3, off)
    return np.array(data, np.float64) * (a * a)


def bp_gain(nt2, dt, f):
    """cuda_bp_filter1d, utilities.cu:733-765: per-bin gain (amp squared) of the padded transform of length nt2."""
    nf = nt2 // 2 + 1
    freq = np.arange(nf) * (1.0 / dt / nt2)
    f0, f1, f2, f3 = [float(v) for v in f]
    a = np.zeros(nf)
    m = (freq >= f0) & (freq < f1)
    a[m] = np.sin(np.pi / 2.0 * (freq[m] - f0) / (f1 - f0))
    a[(freq >= f1) & (freq < f2)] = 1.0
    m = (freq >= f2) & (freq < f3)
    a[m] = np.cos(np.pi / 2.0 * (freq[m] - f2) / (f3 - f2))
    return a * a


def bp_filter(data, dt, f):
    """bp_filter1d, utilities.cu:1115-1168: zero-pad to 2 nt, real FFT, gain, inverse FFT, crop, 1 / (2 nt)."""
    nt = data.shape[-1]
    F = np.fft.rfft(np.asarray(data, np.float64), n=2 * nt, axis=-1)
    return np.fft.irfft(F * bp_gain(2 * nt, dt, f), n=2 * nt, axis=-1)[..., :nt]


def normfact(a, b):
    """cuda_find_normfact, utilities.cu:1011-1040: per-trace dot product + DIVCONST."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64)).sum(axis=-1) + DIVCONST


def cross_misfit(cross, obsn, caln, weights, src_weight):
    """cuda_normal_misfit, utilities.cu:1058-1087 (the factor -2 anticipates the final 0.5, libCUFD.cu:776)."""
    return -2.0 * float((cross / (np.sqrt(obsn) * np.sqrt(caln)) * weights * src_weight).sum())


def cross_adjoint_source(obs, cal, cross, obsn, caln, weights, src_weight):
    """cuda_normal_adjoint_source, utilities.cu:1090-1113."""
    return ((obs - (cross / caln)[:, None] * cal) / (np.sqrt(obsn) * np.sqrt(caln))[:, None]) * (weights * src_weight)[:, None]


def source_update(obs, cal, src, dt):
    """source_update, utilities.cu:1170-1276 (+ cuda_spectrum_update :904-975, amp_ratio_comp :1328-1356): least-squares
    source-signature filter per frequency, applied to the calculated data and to the source.
    Returns (cal_new, src_new, coef [nt + 1] complex, amp_ratio)."""
    nrec, nt = obs.shape
    n2 = 2 * nt
    pad = lambda a: np.concatenate([np.asarray(a, np.float64), np.zeros(a.shape[:-1] + (nt,))], axis=-1)
    Fo = np.fft.rfft(window_simple(pad(obs), dt, 0.01), axis=-1)      # taper on the PADDED length (:1197-1200)
    Fc = np.fft.rfft(window_simple(pad(cal), dt, 0.01), axis=-1)
    Fs = np.fft.rfft(pad(np.asarray(src, np.float64)))                # the source is not tapered (:1202-1203)
    num = (np.conj(Fc) * Fo).sum(axis=0)
    den = (np.conj(Fc) * Fc).sum(axis=0) + 1e-6                       # lambda, :913,955
    coef = num / den
    cal_new = np.fft.irfft(Fc * coef, n=n2, axis=-1)[:, :nt]          # numpy's 1 / n2 = cuda_normalize(1 / nSteps_pad) :1247
    src_new = np.fft.irfft(Fs * coef, n=n2)[:nt]
    mo, mc = np.abs(obs).max(), np.abs(cal_new).max()
    return cal_new, src_new, coef, (float(mo / mc) if mc != 0.0 else 0.0)


def source_update_adj(res, amp_ratio, coef, dt):
    """source_update_adj, utilities.cu:1280-1326: the residual through the same per-frequency filter (multiplied by coef, not by
    its conjugate, as the reference's cuda_filter1d does), scaled by amp_ratio."""
    nrec, nt = res.shape
    n2 = 2 * nt
    padded = np.concatenate([np.asarray(res, np.float64), np.zeros((nrec, nt))], axis=-1)
    F = np.fft.rfft(window_simple(padded, dt, 0.01), axis=-1)
    return amp_ratio * np.fft.irfft(F * coef, n=n2, axis=-1)[:, :nt]


def condition(obs, cal, src, dt, if_win=False, win_start=None, win_end=None, weights=None, src_weight=1.0, win_ratio=0.005,
              filt=None, if_cross_misfit=False, if_src_update=False):
    """The whole data-side chain of one shot in the order of libCUFD.cu:353-457.
    obs, cal [nrec][nt]; returns dict(res, misfit (before the final 0.5), cal, src)."""
    nrec, nt = obs.shape
    w = np.ones(nrec) if weights is None else np.asarray(weights, np.float64)
    obs, cal = np.asarray(obs, np.float64), np.asarray(cal, np.float64)
    src_new = np.asarray(src, np.float64)
    if if_win:                                                        # :353-363
        obs = window_traces(obs, dt, win_start, win_end, w, src_weight, win_ratio)
        cal = window_traces(cal, dt, win_start, win_end, w, src_weight, win_ratio)
    else:                                                             # :363-367: end tapers of win_ratio x record length
        obs, cal = window_simple(obs, dt, win_ratio), window_simple(cal, dt, win_ratio)
    if filt is not None:                                              # :370-373
        obs, cal = bp_filter(obs, dt, filt), bp_filter(cal, dt, filt)
    if if_cross_misfit:                                               # :376-384
        obsn, caln, cross = normfact(obs, obs), normfact(cal, cal), normfact(obs, cal)
    amp_ratio, coef = 1.0, None
    if if_src_update:                                                 # :387-394
        cal, src_new, coef, amp_ratio = source_update(obs, cal, src_new, dt)
    if not if_cross_misfit:                                           # :397-400 / gpuMinus :154-167 + cuda_cal_objective
        res = obs - cal
        res[:, 0] = 0.0
        J = float((res * res).sum())
    else:                                                             # :401-407
        res = np.zeros_like(obs)
        J = cross_misfit(cross, obsn, caln, w, src_weight)
    if if_src_update:                                                 # :430-433
        res = source_update_adj(res, amp_ratio, coef, dt)
    if if_cross_misfit:                                               # :436-443
        res = cross_adjoint_source(obs, cal, cross, obsn, caln, w, src_weight)
    if filt is not None:                                              # :446-448
        res = bp_filter(res, dt, filt)
    if if_win:                                                        # :450-457
        res = window_traces(res, dt, win_start, win_end, w, src_weight, win_ratio)
    else:
        res = window_simple(res, dt, win_ratio)
    return dict(res=res, misfit=J, cal=cal, src=src_new, obs=obs)
