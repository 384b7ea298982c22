"""File formats and model padding of the TorchFWI front-end, same call signatures as
DAS_Waveform_Inversion/Ops/FWI/fwi_utils.py (paraGen :46-83, surveyGen :87-124,
sourceGene :127-140, padding :31-44, padding_numpy_array :11-27).

Both JSON files are written on ONE line: the reference's C++ readers only `getline`
the first line (Src/Parameter.cpp:29, Src/Src_Rec.cu:32) and so does sepfwi_cufd.
"""
import json
import os

import numpy as np


def padded_shape(nz, nx, nPml):
    """nPad rule of the reference drivers (notebooks/Main-001-FWI-Anomaly-Vp-Vs-Den.py:35):
    pad the bottom so that nz + 2 nPml + nPad is a multiple of 32 (a full 32 when already aligned)."""
    nPad = int(32 - np.mod(nz + 2 * nPml, 32))
    return nz + 2 * nPml + nPad, nx + 2 * nPml, nPad


def paraGen(nz, nx, dz, dx, nSteps, dt, f0, nPml, nPad, para_fname, survey_fname, data_dir_name,
            if_win=False, filter_para=None, if_src_update=False, scratch_dir_name='', if_cross_misfit=False,
            das_component=None, max_batch=None, ref_race_compat=None):
    """Write para_file.json.  `das_component` ('exx' | 'ezz'), `max_batch` and `ref_race_compat` are extensions of
    this implementation (the reference selects the fiber direction by editing libCUFD.cu:327-332; ref_race_compat
    reproduces the lost update of its racy residual injection, see include/sepfwi.h)."""
    para = {'nz': nz, 'nx': nx, 'dz': dz, 'dx': dx, 'nSteps': nSteps, 'dt': dt, 'f0': f0,
            'nPoints_pml': nPml, 'nPad': nPad}
    if if_win:
        para['if_win'] = True
    if filter_para is not None:
        para['filter'] = filter_para
    if if_src_update:
        para['if_src_update'] = True
    para['survey_fname'] = survey_fname
    para['data_dir_name'] = data_dir_name
    os.makedirs(data_dir_name, exist_ok=True)
    if if_cross_misfit:
        para['if_cross_misfit'] = True
    if scratch_dir_name != '':
        para['scratch_dir_name'] = scratch_dir_name
        os.makedirs(scratch_dir_name, exist_ok=True)
    if das_component is not None:
        para['das_component'] = das_component
    if max_batch is not None:
        para['max_batch'] = int(max_batch)
    if ref_race_compat:
        para['ref_race_compat'] = True
    with open(para_fname, 'w') as fp:
        json.dump(para, fp)
    return para


def surveyGen(z_src, x_src, z_rec, x_rec, survey_fname, Windows=None, Weights=None, Src_Weights=None,
              Src_rxz=None, Rec_rxz=None, Das_sensitivity=None):
    """Write survey_file.json; every shot shares the receiver line (interior grid indices).
    `Das_sensitivity` (extension): (nrec, 3) weights of (exx, ezz, exz) per channel -- an arbitrarily oriented
    fiber as in DAS_Waveform_Modeling/src/elasticSolver.py:270-276 instead of the reference op's straight horizontal /
    vertical fiber (libCUFD.cu:327-332); recorded as ett and injected by the adjoint with the same weights."""
    z_src, x_src = np.asarray(z_src).tolist(), np.asarray(x_src).tolist()
    z_rec, x_rec = np.asarray(z_rec).tolist(), np.asarray(x_rec).tolist()
    survey = {'nShots': len(x_src)}
    for i in range(len(x_src)):
        shot = {'z_src': z_src[i], 'x_src': x_src[i], 'nrec': len(x_rec), 'z_rec': z_rec, 'x_rec': x_rec}
        if Windows is not None:
            shot['win_start'] = Windows['shot' + str(i)]['start']
            shot['win_end'] = Windows['shot' + str(i)]['end']
        if Weights is not None:
            shot['weights'] = Weights['shot' + str(i)]['weights']
        if Src_Weights is not None:
            shot['src_weight'] = Src_Weights[i]
        if Src_rxz is not None:
            shot['src_rxz'] = Src_rxz[i]
        if Rec_rxz is not None:
            shot['rec_rxz'] = np.asarray(Rec_rxz).tolist()
        if Das_sensitivity is not None:
            w = np.asarray(Das_sensitivity, np.float64).reshape(len(x_rec), 3)
            shot['das_sensitivity'] = w.tolist()
        survey['shot' + str(i)] = shot
    with open(survey_fname, 'w') as fp:
        json.dump(survey, fp)
    return survey


def sourceGene(f, nStep, delta_t):
    """Ricker wavelet, delay 1.2/f, amplitude 1e7 (fwi_utils.py:127-140), float64."""
    e = np.pi * np.pi * f * f
    t = delta_t * np.arange(nStep) - 1.2 / f
    return (1 - 2 * e * t ** 2) * np.exp(-e * t ** 2) * 1.0e7


def padding(cp, cs, den, nz_orig, nx_orig, nz, nx, nPml, nPad):
    """Bilinear resize (nz_orig,nx_orig)->(nz,nx) then replicate-pad by nPml (bottom nPml+nPad);
    differentiable torch op, fwi_utils.py:31-44."""
    import torch.nn.functional as F
    out = []
    for t in (cp, cs, den):
        t = t.view(1, 1, nz_orig, nx_orig)
        t = F.interpolate(t, size=(nz, nx), mode='bilinear', align_corners=False)
        t = F.pad(t, pad=(nPml, nPml, nPml, nPml + nPad), mode='replicate')
        out.append(t.view(nz + 2 * nPml + nPad, nx + 2 * nPml))
    return tuple(out)


def padding_numpy_array(arr, npml, npad):
    """Replicate-pad a (nz, nx) numpy array like `padding` does (fwi_utils.py:11-27)."""
    return np.pad(arr, ((npml, npml + npad), (npml, npml)), mode='edge')


def read_json_first_line(fname):
    with open(fname, 'r') as fp:
        return json.loads(fp.readline())


def load_survey(survey_fname, shot_ids, nPml):
    """Parse survey_file.json for `shot_ids` the way Src_Rec does (Src/Src_Rec.cu:74-120):
    +nPml on every index.  Returns a list of dict(zs, xs, zrec, xrec, src_rxz, weights, win_start, win_end, trace_weights, src_weight)."""
    js = read_json_first_line(survey_fname)
    out = []
    for sid in shot_ids:
        s = js['shot%d' % int(sid)]
        n = int(s['nrec'])
        out.append(dict(zs=int(s['z_src']) + nPml, xs=int(s['x_src']) + nPml,
                        zrec=np.asarray(s['z_rec'][:n], np.int32) + nPml,
                        xrec=np.asarray(s['x_rec'][:n], np.int32) + nPml,
                        src_rxz=float(s.get('src_rxz', 1.0)),
                        weights=(np.asarray(s['das_sensitivity'], np.float32).reshape(n, 3)
                                 if 'das_sensitivity' in s else None),
                        # data-side options (Src_Rec.cu:145-201): windows in seconds, trace weights, source weight
                        win_start=(np.asarray(s['win_start'][:n], np.float32) if 'win_start' in s else None),
                        win_end=(np.asarray(s['win_end'][:n], np.float32) if 'win_end' in s else None),
                        trace_weights=(np.asarray(s['weights'][:n], np.float32) if 'weights' in s else None),
                        src_weight=float(s.get('src_weight', 1.0))))
    return out
