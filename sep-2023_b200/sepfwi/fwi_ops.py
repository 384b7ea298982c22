"""Drop-in for the reference's pybind module `fwi` (DAS_Waveform_Inversion/Ops/FWI/Src/Torch_Fwi.cpp):

    forward (Lambda, Mu, Den, Stf, gpu_id, Shot_ids, para_fname) -> [misfit]                       (:12-36)
    backward(Lambda, Mu, Den, Stf, ngpu,   Shot_ids, para_fname) -> [misfit, gLam, gMu, gDen, gStf] (:38-104)
    obscalc (Lambda, Mu, Den, Stf, ngpu,   Shot_ids, para_fname) -> None, writes Shot_*.bin        (:106-136)

Same argument meaning: Lambda/Mu in MPa and Den in kg/m^3 as (nz_pad, nx_pad) float32 tensors, Stf
(nSrc, nSteps), Shot_ids int32, para_fname the para_file.json written by fwi_utils.paraGen.  Differences,
all deliberate and documented in INTEGRATION.md:
  * tensors may live on the GPU (the reference only takes CPU tensors); results come back on the inputs' device;
  * errors raise RuntimeError instead of exit(1);
  * `ngpu` > 1 in one process drives GPUs 0..ngpu-1 from threads like the reference's OpenMP loop; under
    torch.distributed (one process per GPU) the shots are sharded over the ranks instead and the partial
    gradients are summed by a single NCCL all-reduce (sepfwi.dist);
  * gStf has the row of EVERY processed shot at its shot id (the reference keeps GPU 0's rows at local
    indices only, Torch_Fwi.cpp:103 / libCUFD.cu:671-673).
Handles, CPML tables and observed data are cached between calls (the reference rebuilds them per call).
"""
import os
import threading

import numpy as np

from . import _lib, dist, fwi_utils
from .engine import Propagator, ShotSpec

_PROPS = {}
_OBS = {}
_AUTOB = {}      # (propagator key, nrec, nshots) -> shots per launch: torch.cuda.mem_get_info costs ~50 ms, a third of a small evaluation
_LOCK = threading.Lock()


def clear_cache():
    with _LOCK:
        for p in _PROPS.values():
            p.close()
        _PROPS.clear()
        _OBS.clear()
        _AUTOB.clear()
    dist.clear_packed()


def _torch():
    import torch
    return torch


def auto_batch(nz, nx, nPad, nSteps, nrec, nPml, with_adjoint, nshots, device, params=None):
    """Concurrent shots per launch: enough to give every launch a few million cells, within memory.  The bytes one slot
    costs come from the library (sepfwi_bytes_per_slot: state block, boundary ring, traces, gradients, tables) plus the
    device-cached observed data of the shot; at most 70 % of what is free now, split into equal batches."""
    torch = _torch()
    cells = float((nz - nPad) * nx)
    # measured on B200 (C3 grid, 730 k cells): 5.7 / 6.6 / 7.2 / 7.8 shot-gradients/s at 8 / 16 / 32 / 64 shots per launch -- the tails of
    # every launch and the slow CPML items amortise over the batch; small grids: one launch for the whole survey beats two
    B = min(64, int(5.0e7 / cells) + 1)
    if params is None:
        params = _lib.Params(int(nz), int(nx), int(nPml), int(nPad), int(nSteps), 1.0, 1.0, 1.0, 1.0, 0, 0, 1, max(1, int(nrec)),
                             1 if with_adjoint else 0, 0, 0)
    per = float(_lib.lib().sepfwi_bytes_per_slot(params)) + 4.0 * nrec * nSteps
    free = torch.cuda.mem_get_info(device)[0]
    while B > 1 and per * B > 0.7 * free:
        B -= 1
    B = max(1, min(B, nshots))
    nbatch = -(-nshots // B)
    return -(-nshots // nbatch)          # equal batches: 64 shots at a limit of 59 run as 32 + 32, not 59 + 5


def _prop(para, device, with_adjoint, nrec, nshots):
    fiber = _lib.FIBER_EZZ if para.get("das_component", "exx") == "ezz" else _lib.FIBER_EXX
    race = bool(para.get("ref_race_compat", False))
    key = (device, para["nz"], para["nx"], para["nPoints_pml"], para["nPad"], para["nSteps"], float(para["dz"]),
           float(para["dx"]), float(para["dt"]), float(para["f0"]), fiber, bool(with_adjoint), race)
    B = int(para.get("max_batch", 0))
    if not B:
        with _LOCK:
            B = _AUTOB.get((key, nrec, nshots), 0)
        if not B:
            B = auto_batch(para["nz"], para["nx"], para["nPad"], para["nSteps"], nrec, para["nPoints_pml"], with_adjoint, nshots, device)
            with _LOCK:
                _AUTOB[(key, nrec, nshots)] = B
    with _LOCK:
        p = _PROPS.get(key)
        if p is not None and (p.params.max_nrec < nrec or p.params.max_batch < min(B, nshots)):
            p.close()
            p = None
        while p is None:
            try:
                p = Propagator(para["nz"], para["nx"], para["nPoints_pml"], para["nPad"], para["nSteps"], para["dz"],
                               para["dx"], para["dt"], para["f0"], fiber=fiber, max_batch=B, max_nrec=nrec,
                               with_adjoint=with_adjoint, device=device, ref_race_compat=race)
            except _lib.SepfwiError as e:
                if e.code != -5 or B <= 1:       # SEPFWI_ENOMEM: other tenants of the GPU; halve the batch and retry
                    raise
                B = max(1, B // 2)
                if not int(para.get("max_batch", 0)):
                    _AUTOB[(key, nrec, nshots)] = B      # remember what fitted: the next call must not ask for the larger batch again
        _PROPS[key] = p
    return p


def _shots(para, shot_ids, Stf):
    torch = _torch()
    stf = Stf.detach().cpu().numpy() if isinstance(Stf, torch.Tensor) else np.asarray(Stf)
    stf = np.ascontiguousarray(stf, np.float32)
    sv = fwi_utils.load_survey(para["survey_fname"], shot_ids, para["nPoints_pml"])
    return [ShotSpec(s["zs"], s["xs"], s["zrec"], s["xrec"], stf[int(sid)], s["src_rxz"], weights=s.get("weights"),
                     win_start=s.get("win_start"), win_end=s.get("win_end"), trace_weights=s.get("trace_weights"),
                     src_weight=s.get("src_weight", 1.0))
            for s, sid in zip(sv, shot_ids)], stf.shape


def _data_options(para):
    """The data-side switches of para_file.json (Src/Parameter.cpp:139-176) as Propagator.set_data_options keywords."""
    return dict(if_win=bool(para.get("if_win", False)), filter=para.get("filter"),
                if_cross_misfit=bool(para.get("if_cross_misfit", False)), if_src_update=bool(para.get("if_src_update", False)))


_OBS_CAP_BYTES = 64 << 30      # device bytes the observed-data cache may hold per process (least recently used goes first)


def _obs(para, sid, nrec, device):
    """Shot_ett{id}.bin (libCUFD.cu:221-223), cached on the device.  One entry per (path, device): a regenerated file
    (other mtime / size) REPLACES its entry, and the cache as a whole is bounded (LRU)."""
    torch = _torch()
    path = os.path.join(para["data_dir_name"], "Shot_ett%d.bin" % int(sid))
    try:
        st = os.stat(path)
    except OSError:
        raise RuntimeError("File reading error! Attempted to read %s" % path)
    key, stamp = (path, device), (st.st_mtime_ns, st.st_size)
    with _LOCK:
        ent = _OBS.pop(key, None)
        if ent is not None and ent[0] == stamp:
            _OBS[key] = ent                       # re-inserted last: most recently used
            return ent[1]
    a = np.fromfile(path, np.float32)
    if a.size != nrec * para["nSteps"]:
        raise RuntimeError("%s holds %d floats, expected %d x %d" % (path, a.size, nrec, para["nSteps"]))
    t = torch.from_numpy(a.reshape(nrec, para["nSteps"])).to("cuda:%d" % device)
    with _LOCK:
        _OBS[key] = (stamp, t)
        total = sum(v[1].numel() * 4 for v in _OBS.values())
        for k in list(_OBS):
            if total <= _OBS_CAP_BYTES or k == key:
                break
            total -= _OBS[k][1].numel() * 4
            del _OBS[k]
    return t


def _ids(Shot_ids):
    torch = _torch()
    if isinstance(Shot_ids, torch.Tensor):
        return [int(v) for v in Shot_ids.detach().cpu().reshape(-1).tolist()]
    return [int(v) for v in np.asarray(Shot_ids).reshape(-1).tolist()]


def _dev_of(t, default=None):
    torch = _torch()
    if isinstance(t, torch.Tensor) and t.is_cuda:
        return t.device.index
    if default is not None:
        return default
    if not torch.cuda.is_available():
        raise RuntimeError("sepfwi needs a CUDA device; there is no CPU fallback")
    return torch.cuda.current_device()


def _model_on(device, *tensors):
    torch = _torch()
    out = []
    for t in tensors:
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t))
        out.append(t.detach().to(device="cuda:%d" % device, dtype=torch.float32).contiguous())
    return out


def _gradient_on_device(device, Lambda, Mu, Den, Stf, ids, para, with_adj, packed=False):
    """Misfit (+gradients) of shots `ids` on one GPU.  Returns (misfit, gl, gm, gd, gstf_full) with CUDA tensors; with
    packed=True the gradients are written straight into this device's persistent all-reduce buffer and the
    dist.PackedGradients object is returned instead."""
    torch = _torch()
    with torch.cuda.device(device):
        shots, stf_shape = _shots(para, ids, Stf)
        nrec = max([s.nrec for s in shots] + [1])
        P = _prop(para, device, with_adj, nrec, len(ids))
        lam, mu, den = _model_on(device, Lambda, Mu, Den)
        P.set_model(lam, mu, den)
        opts = _data_options(para)
        if getattr(P, "_data_opts", None) != opts and (any(v for v in opts.values()) or getattr(P, "_data_opts", None) is not None):
            P.set_data_options(**opts)
            P._data_opts = opts
        obs = [_obs(para, sid, s.nrec, device) for sid, s in zip(ids, shots)]
        if packed:
            pk = dist.packed_for((para["nz"], para["nx"]), stf_shape, "cuda:%d" % device)
            r = P.gradient(shots, obs, with_adj=True, device=True, grad_out=pk.views())
            pk.set_local(r["misfit64"], dict(zip(ids, r["gstf"])))
            return pk
        r = P.gradient(shots, obs, with_adj=with_adj, device=True)
        gstf = torch.zeros(stf_shape, dtype=torch.float32, device="cuda:%d" % device)
        if with_adj:
            stage = torch.zeros(stf_shape, dtype=torch.float32)
            for sid, g in zip(ids, r["gstf"]):
                stage[int(sid)] = torch.from_numpy(g)
            gstf.copy_(stage)
        return r["misfit64"], r["glam"], r["gmu"], r["grho"], gstf


def forward(Lambda, Mu, Den, Stf, gpu_id, Shot_ids, para_fname):
    """Misfit only (cufd calc_id = 0) on GPU `gpu_id`."""
    torch = _torch()
    para = fwi_utils.read_json_first_line(para_fname)
    J = _gradient_on_device(int(gpu_id), Lambda, Mu, Den, Stf, _ids(Shot_ids), para, False)[0]
    return [torch.tensor([J], dtype=torch.float32)]


def backward(Lambda, Mu, Den, Stf, ngpu, Shot_ids, para_fname):
    torch = _torch()
    para = fwi_utils.read_json_first_line(para_fname)
    ids = _ids(Shot_ids)
    out_dev = Lambda.device if isinstance(Lambda, torch.Tensor) else torch.device("cpu")
    if dist.is_distributed():
        ws, rank = dist.world()
        dev = _dev_of(Lambda)
        pk = _gradient_on_device(dev, Lambda, Mu, Den, Stf, dist.shard(ids, ws, rank), para, True, packed=True)
        J, gl, gm, gd, gs = pk.allreduce()
        gl, gm, gd, gs = gl.clone(), gm.clone(), gd.clone(), gs.clone()      # the packed buffer is reused by the next call
    elif int(ngpu) <= 1:
        J, gl, gm, gd, gs = _gradient_on_device(_dev_of(Lambda), Lambda, Mu, Den, Stf, ids, para, True)
    else:
        ngpu = int(ngpu)
        if ngpu > len(ids):
            raise RuntimeError("The number of GPUs should be smaller than the number of shots!")
        if ngpu > torch.cuda.device_count():
            raise RuntimeError("ngpu=%d but only %d CUDA devices are visible" % (ngpu, torch.cuda.device_count()))
        res = [None] * ngpu
        err = []

        def work(i):
            try:
                res[i] = _gradient_on_device(i, Lambda, Mu, Den, Stf, dist.shard(ids, ngpu, i), para, True)
            except Exception as e:   # surfaced below: a failing GPU must not be silent
                err.append(e)

        th = [threading.Thread(target=work, args=(i,)) for i in range(ngpu)]
        [t.start() for t in th]
        [t.join() for t in th]
        if err:
            raise err[0]
        J = sum(r[0] for r in res)
        gl, gm, gd, gs = (sum(r[k].to("cuda:0") for r in res) for k in range(1, 5))
    return [torch.tensor([J], dtype=torch.float32).to(out_dev), gl.to(out_dev), gm.to(out_dev), gd.to(out_dev),
            gs.to(out_dev)]


def obscalc(Lambda, Mu, Den, Stf, ngpu, Shot_ids, para_fname):
    """Forward-model the shots and write Shot_{pr,vx,vz,ett}{id}.bin into data_dir_name (libCUFD.cu:755-769)."""
    torch = _torch()
    para = fwi_utils.read_json_first_line(para_fname)
    ids = _ids(Shot_ids)
    if dist.is_distributed():
        ws, rank = dist.world()
        groups = [(_dev_of(Lambda), dist.shard(ids, ws, rank))]
    elif int(ngpu) <= 1:
        groups = [(_dev_of(Lambda), ids)]
    else:
        if int(ngpu) > len(ids):
            raise RuntimeError("The number of GPUs should be smaller than the number of shots!")
        groups = [(i, dist.shard(ids, int(ngpu), i)) for i in range(int(ngpu))]
    os.makedirs(para["data_dir_name"], exist_ok=True)
    err = []

    def work(dev, sub):
        try:
            with torch.cuda.device(dev):
                shots, _ = _shots(para, sub, Stf)
                nrec = max([s.nrec for s in shots] + [1])
                P = _prop(para, dev, False, nrec, len(sub))
                P.set_model(*_model_on(dev, Lambda, Mu, Den))
                out = P.forward(shots)
                for sid, o in zip(sub, out):
                    for c in ("pr", "vx", "vz", "ett"):
                        o[c].tofile(os.path.join(para["data_dir_name"], "Shot_%s%d.bin" % (c, sid)))
        except Exception as e:
            err.append(e)

    th = [threading.Thread(target=work, args=g) for g in groups]
    [t.start() for t in th]
    [t.join() for t in th]
    if err:
        raise err[0]
