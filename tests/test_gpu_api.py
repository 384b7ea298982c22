"""GPU tests of the reference-facing Python API (run with -m gpu): the GPU-backed elasticSolver against
the Numba golden vectors, the fwi_ops / FWI autograd front-end with its file side channels, and the
reference's own published L-BFGS log values."""
import os

import numpy as np
import pytest

import problems
from util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["numba_homog", "numba_hetero", "numba_wavefield"])
def test_elastic_solver_matches_numba_reference(golden_dir, name):
    """north_star: DAS seismograms within 1e-4 relative L2 of the Numba CPU modelling, computing in fp32."""
    from sepfwi.elasticSolver import elasticSolver
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    kw = {k[3:]: g[k] for k in g.files if k.startswith("in_")}
    args = {k: (v.item() if v.ndim == 0 else v) for k, v in kw.items()}
    solver = elasticSolver(**args)
    sol = solver.forward()
    assert len(sol) == kw["src_coord"].shape[0]
    for isrc, s in enumerate(sol):
        for k in ("vx", "vz", "pr", "ett", "exx", "ezz", "exz"):
            ref = g["out%d_%s" % (isrc, k)]
            assert s[k].shape == ref.shape and s[k].dtype == np.float64
            assert rel_l2(s[k], ref) < 1e-4, (name, isrc, k, rel_l2(s[k], ref))
    one = solver.forward_it(0, False)
    assert np.array_equal(one["ett"], sol[0]["ett"])
    if name == "numba_wavefield":
        # forward_it(isrc, save_wavefield=True): interior snapshots every save_step = 10 steps, (save_num, nx, nz) arrays
        snap = solver.forward_it(0, True)
        assert np.array_equal(snap["ett"], one["ett"])
        for k in ("sxx_wavefield", "szz_wavefield", "vx_wavefield", "vz_wavefield"):
            ref = g["out0_" + k]
            assert snap[k].shape == ref.shape == (95 // 10 + 1, 50, 40) and snap[k].dtype == np.float64
            assert np.abs(ref[-1]).max() > 0 and rel_l2(snap[k], ref) < 1e-4, (k, rel_l2(snap[k], ref))
        both = solver.forward(save_wavefield=True)
        assert np.array_equal(both[0]["vx_wavefield"], snap["vx_wavefield"])
    with pytest.raises(ValueError):
        solver.set_model(kw["vp"][:-1], kw["vs"][:-1], kw["rho"][:-1])


def test_c1_gpu_solver_against_numba_and_aki_richards(golden_dir):
    """BASELINE.json configs[0] on the GPU: 2-D homogeneous 201 x 201 grid, explosive source -- within 1e-4 relative L2 of the
    reference's Numba modelling (fp32 vs fp64) and through the reference's independent Aki & Richards check."""
    from sepfwi.elasticSolver import elasticSolver
    from util import analytic_agreement, assert_analytic_agreement
    g = np.load(os.path.join(golden_dir, "analytic_c1.npz"))
    kw = {k[3:]: (g[k].item() if g[k].ndim == 0 else g[k]) for k in g.files if k.startswith("in_")}
    sol = elasticSolver(**kw).forward()[0]
    for k in ("vx", "vz", "pr", "ett", "exx", "ezz", "exz"):
        ref = g["numba_" + k]
        assert sol[k].shape == ref.shape
        if np.abs(ref).max() > 0:
            # rows that vanish by symmetry (vz, exz of the receiver on the source's z level) carry rounding noise only
            live = np.abs(ref).max(axis=1) > 1e-9 * np.abs(ref).max()
            assert rel_l2(sol[k][live], ref[live]) < 1e-4, (k, rel_l2(sol[k][live], ref[live]))
    assert_analytic_agreement(analytic_agreement(sol, g))


def _write_problem(prob, tmp, **para_kw):
    from sepfwi import fwi_utils as ft
    para, survey, data = os.path.join(tmp, "para.json"), os.path.join(tmp, "survey.json"), os.path.join(tmp, "Data")
    ft.paraGen(prob.nz, prob.nx, prob.dz, prob.dx, prob.nSteps, prob.dt, prob.f0, prob.nPml, prob.nPad, para, survey, data, **para_kw)
    ft.surveyGen(prob.z_src, prob.x_src, prob.z_rec, prob.x_rec, survey)
    return para, data


def test_fwi_ops_matches_oracle_through_files(tmp_path):
    """obscalc -> Shot_*.bin -> backward / forward, CPU tensors in, CPU tensors out (the reference's calling convention)."""
    import torch
    from oracle import oracle as O
    from sepfwi import fwi_ops
    from util import oracle_par
    prob = problems.tiny()
    para, data = _write_problem(prob, str(tmp_path))
    ids = torch.arange(prob.nshots, dtype=torch.int32)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    stf = T(prob.stf)
    fwi_ops.obscalc(*map(T, prob.true), stf, 1, ids, para)
    par = oracle_par(O, prob)
    obs = {}
    for sid, (zs, xs, zr, xr) in prob.survey().items():
        ref = O.forward(par, *prob.true, prob.stf[sid], zs, xs, zr, xr)
        for c in ("pr", "vx", "vz", "ett"):
            got = np.fromfile(os.path.join(data, "Shot_%s%d.bin" % (c, sid)), np.float32).reshape(len(zr), prob.nSteps)
            assert rel_l2(got, ref[c]) < 2e-5
        obs[sid] = np.fromfile(os.path.join(data, "Shot_ett%d.bin" % sid), np.float32).reshape(len(zr), prob.nSteps)
    J, gl, gm, gd, gs = O.fwi_backward(par, *prob.start, prob.stf, 1, np.arange(prob.nshots), prob.survey(), obs)
    out = fwi_ops.backward(*map(T, prob.start), stf, 1, ids, para)
    assert not out[1].is_cuda and out[0].shape == (1,)
    assert abs(out[0].item() - J) <= 2e-5 * J
    for got, ref in zip(out[1:], (gl, gm, gd, gs)):
        assert rel_l2(got.numpy(), ref) < 2e-4
    assert abs(fwi_ops.forward(*map(T, prob.start), stf, 0, ids, para)[0].item() - J) <= 2e-5 * J
    # CUDA tensors in -> CUDA tensors out, same numbers
    outc = fwi_ops.backward(*[T(a).cuda() for a in prob.start], stf, 1, ids, para)
    assert outc[1].is_cuda and torch.equal(outc[1].cpu(), out[1])
    # a subset of the shots, in a different order, fills exactly those stf-gradient rows
    out1 = fwi_ops.backward(*map(T, prob.start), stf, 1, torch.tensor([1], dtype=torch.int32), para)
    assert torch.all(out1[4][0] == 0) and torch.equal(out1[4][1], out[4][1])
    with pytest.raises(RuntimeError):
        fwi_ops.backward(*map(T, prob.start), stf, 3, ids, para)       # more GPUs than shots
    fwi_ops.clear_cache()


def test_multi_gpu_in_process_equals_single_gpu(tmp_path):
    """ngpu = 2 (threads over GPUs 0,1, like the reference's OpenMP loop) gives the single-GPU sums."""
    import torch
    from sepfwi import fwi_ops
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    prob = problems.small()
    para, data = _write_problem(prob, str(tmp_path))
    ids = torch.arange(prob.nshots, dtype=torch.int32)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    stf = T(prob.stf)
    fwi_ops.obscalc(*map(T, prob.true), stf, 2, ids, para)
    a = fwi_ops.backward(*map(T, prob.start), stf, 1, ids, para)
    b = fwi_ops.backward(*map(T, prob.start), stf, 2, ids, para)
    assert abs(a[0].item() - b[0].item()) <= 1e-6 * abs(a[0].item())
    for x, y in zip(a[1:], b[1:]):
        assert rel_l2(y.numpy(), x.numpy()) < 1e-6
    fwi_ops.clear_cache()


def test_reference_notebook_lbfgs_iterate0(tmp_path):
    """The reference's own published numbers (notebooks/001-FWI-Anomaly-Vp-Vs-Den.ipynb cell 7 output):
        At iterate 0  f = 1.51116D+04  |proj g| = 2.14289D+00
    for the fully in-notebook-specified anomaly problem (cell 3; Main-001-...py:28-135): 101 x 201, 19 shots,
    181 ADJACENT receivers, nt = 1501, FWI module with the top-4-rows mask, homogeneous start.
    f(x0) only involves the forward path.  |proj g| is max |dJ/d(Vp,Vs,Den)| and carries the reference's
    residual-injection race (adjacent receivers straddle five 32-thread block seams), so it is compared with
    ref_race_compat on; the race-free gradient is reported alongside."""
    import torch
    from sepfwi import FWI_ops as F, fwi_ops, fwi_utils as ft
    nz, nx, nPml, dz, dx, dt, nt, f0 = 101, 201, 32, 20.0, 20.0, 0.002, 1501, 10.0
    nz_pad, nx_pad, nPad = ft.padded_shape(nz, nx, nPml)
    vp = np.ones((nx, nz)) * 4000.0
    vs = np.ones((nx, nz)) * 4000.0 / 1.732
    rho = np.ones((nx, nz)) * 2500.0
    vp[42:58, 42:58] += 80.0
    vs[92:108, 42:58] -= 80.0 / 1.732
    rho[142:158, 42:58] += 40
    true = [a.T.astype("float32") for a in (vp, vs, rho)]              # saved transposed, loaded as float32
    init = [np.full((nz, nx), v, "float32") for v in (4000.0, 4000.0 / 1.732, 2500.0)]
    x_src = np.arange(10, nx - 10, 10).astype(int)
    z_src = np.ones(len(x_src), int)
    x_rec = np.arange(10, nx - 10).astype(int)
    z_rec = 95 * np.ones(len(x_rec), int)
    Mask = np.zeros((nz_pad, nx_pad))
    Mask[nPml:nPml + nz, nPml:nPml + nx] = 1.0
    Mask[nPml:nPml + 4, :] = 0.0
    stf = torch.tensor(ft.sourceGene(f0, nt, dt), dtype=torch.float32).repeat(len(x_src), 1)
    ids = torch.tensor(np.arange(len(x_src)), dtype=torch.int32)
    res = {}
    for compat in (True, False):
        work = str(tmp_path / ("c%d" % compat))
        os.makedirs(work)
        para, survey, data = work + "/para_file.json", work + "/survey_file.json", work + "/Data"
        ft.paraGen(nz_pad, nx_pad, dz, dx, nt, dt, f0, nPml, nPad, para, survey, data, ref_race_compat=compat)
        ft.surveyGen(z_src, x_src, z_rec, x_rec, survey)
        pads = [torch.tensor(ft.padding_numpy_array(a, nPml, nPad), dtype=torch.float32) for a in true]
        F.FWI_obscalc(*pads, stf, para)(ids, ngpu=1)
        opt = dict(nz=nz, nx=nx, nz_orig=nz, nx_orig=nx, nPml=nPml, nPad=nPad, para_fname=para)
        th = [torch.tensor(a, dtype=torch.float32, requires_grad=True) for a in init]
        fwi = F.FWI(*th, stf, opt, Mask=torch.tensor(Mask, dtype=torch.float32))
        loss = fwi(ids, ngpu=1)
        loss.backward()
        g = np.concatenate([p.grad.numpy().astype(np.float64).ravel() for p in (fwi.Vp, fwi.Vs, fwi.Den)])
        assert g.size == 60903                                         # "N = 60903" in the log
        res[compat] = (loss.item(), np.abs(g).max())
    msg = ("f(x0) = %.6e (reference log 1.51116e+04); |proj g| = %.6f with the seam update lost as on B200, %.6f race-free "
           "(reference log 2.14289, 4 GPUs of another generation)" % (res[True][0], res[True][1], res[False][1]))
    print(msg)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        open(os.path.join(out, "notebook_golden.txt"), "w").write(msg + "\n")
    assert abs(res[True][0] - 1.51116e4) <= 1e-5 * 1.51116e4 * 2      # printed with 6 significant digits
    assert res[False][0] == res[True][0]                               # the misfit does not involve the adjoint
    # which update the reference's race loses depends on its block scheduling (hardware / driver): the logged value
    # cannot be pinned to 6 digits, only bracketed by the two deterministic gradients
    assert min(abs(res[True][1] - 2.14289), abs(res[False][1] - 2.14289)) <= 0.02 * 2.14289
    fwi_ops.clear_cache()


def test_lbfgs_driver_follows_the_reference_notebook_trace(tmp_path):
    """PyTorchObjective + scipy L-BFGS-B with the reference's options on the reference's own experiment 001
    (sepfwi.drivers mirrors Main-001-FWI-Anomaly-Vp-Vs-Den.py).  The notebook logs
        iterate 0  f = 1.51116D+04,  iterate 1  f = 1.13748D+04,  iterate 2  f = 3.05521D+03,  iterate 3  f = 2.11215D+03
    The first line search only depends on f and g at x0, so iterates 1-2 test the whole gradient (all three parameter
    classes through the autograd chain), not just its max-norm.  The reference's gradient carries its residual-injection
    race (see test_reference_notebook_lbfgs_iterate0), so the trace is followed with ref_race_compat on and only to the
    accuracy the race's hardware dependence allows; the race-free path must decrease the misfit at least as fast."""
    from sepfwi import drivers, fwi_ops
    prob = drivers.anomaly_problem("001")
    notebook = [1.51116e4, 1.13748e4, 3.05521e3, 2.11215e3]
    traces = {}
    for compat in (True, False):
        files = drivers.write_files(prob, str(tmp_path / ("c%d" % compat)), ref_race_compat=compat or None)
        drivers.generate_data(prob, files)
        fwi, obj, log = drivers.invert(prob, files, nIter=3)
        traces[compat] = [f for _, f, _, _ in log]
        assert obj.nfev <= 8 and len(log) == 4
        assert fwi.Vp.grad is not None and fwi.Vp.grad.shape == (prob.nz, prob.nx)
    msg = "L-BFGS f trace: notebook %s | race-compat %s | race-free %s" % (
        ["%.5e" % v for v in notebook], ["%.5e" % v for v in traces[True]], ["%.5e" % v for v in traces[False]])
    print(msg)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        open(os.path.join(out, "notebook_lbfgs_trace.txt"), "w").write(msg + "\n")
    assert abs(traces[True][0] - notebook[0]) <= 2e-5 * notebook[0]
    assert abs(traces[True][1] - notebook[1]) <= 0.02 * notebook[1]          # first step: direction = -g(x0)
    for tr in traces.values():
        assert tr[1] < tr[0] and tr[2] < 0.5 * tr[1] and tr[3] < tr[2]       # same order of decrease as the notebook
        assert tr[3] < 0.2 * tr[0]
    fwi_ops.clear_cache()


def test_fwi_op_with_per_channel_das_sensitivity(tmp_path):
    """arbitrarily oriented fiber through the reference-facing op.  `das_sensitivity` rows (exx, ezz, exz) in
    survey_file.json: (1, 0, 0) must reproduce the straight horizontal fiber of the stock op bit-exactly (traces, misfit and
    gradients); a random orientation per channel must match the engine called directly with the same weights, and its
    gradient must pass a directional finite-difference check of the misfit (the adjoint injects with the same weights)."""
    import torch
    from sepfwi import fwi_ops, fwi_utils as ft
    prob = problems.small()
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    ids = torch.arange(prob.nshots, dtype=torch.int32)
    stf = T(prob.stf)
    nrec = len(prob.x_rec)

    def run(tag, sens):
        work = str(tmp_path / tag)
        os.makedirs(work)
        para, survey, data = work + "/para.json", work + "/survey.json", work + "/Data"
        ft.paraGen(prob.nz, prob.nx, prob.dz, prob.dx, prob.nSteps, prob.dt, prob.f0, prob.nPml, prob.nPad, para, survey, data)
        ft.surveyGen(prob.z_src, prob.x_src, prob.z_rec, prob.x_rec, survey, Das_sensitivity=sens)
        fwi_ops.obscalc(*map(T, prob.true), stf, 1, ids, para)
        ett = [np.fromfile(os.path.join(data, "Shot_ett%d.bin" % s), np.float32).reshape(nrec, prob.nSteps) for s in range(prob.nshots)]
        return para, ett, fwi_ops.backward(*map(T, prob.start), stf, 1, ids, para)

    para0, ett0, out0 = run("stock", None)
    _, ett1, out1 = run("exx", np.tile([1.0, 0.0, 0.0], (nrec, 1)))
    for a, b in zip(ett0, ett1):
        assert np.array_equal(a, b)
    for a, b in zip(out0, out1):
        assert torch.equal(a, b)
    rng = np.random.default_rng(3)
    ang = rng.uniform(0, np.pi, nrec)                                  # fiber direction per channel
    sens = np.stack([np.cos(ang) ** 2, np.sin(ang) ** 2, 2 * np.sin(ang) * np.cos(ang)], 1)
    para, ett2, out2 = run("oriented", sens)
    assert rel_l2(ett2[0], ett0[0]) > 0.1                              # it really is a different measurement
    # Directional derivative of the misfit along a smooth Mu perturbation.  The reference's gradient is a continuous-adjoint
    # gradient (zeroed first sample, tapered stf, boundary-saving reconstruction), not the exact transpose of the discrete
    # forward map: for the stock fiber it agrees with central differences to ~10 %.  The oriented fiber must be as consistent
    # as the stock one -- wrong or missing weights in the adjoint injection would show up as a different ratio.
    lam, mu, den = map(T, prob.start)
    zz, xx = np.meshgrid(np.arange(prob.nz), np.arange(prob.nx), indexing="ij")
    dmu = T((np.exp(-((zz - prob.nz / 2) ** 2 + (xx - prob.nx / 2) ** 2) / (2 * 12.0 ** 2)) * 0.01 * prob.start[1].mean()).astype(np.float32))
    ratio = {}
    for tag, p_, out in (("stock", para0, out0), ("oriented", para, out2)):
        Jp = fwi_ops.forward(lam, mu + dmu, den, stf, 0, ids, p_)[0].item()
        Jm = fwi_ops.forward(lam, mu - dmu, den, stf, 0, ids, p_)[0].item()
        ratio[tag] = float((out[2].double() * dmu.double()).sum()) / ((Jp - Jm) / 2.0)
    assert abs(ratio["stock"] - 1.0) < 0.2 and abs(ratio["oriented"] - ratio["stock"]) < 0.05, ratio
    fwi_ops.clear_cache()


def test_one_process_per_gpu_nccl_matches_single_process(tmp_path):
    """torchrun, one process per GPU, shots sharded like Torch_Fwi.cpp:59-80, ONE NCCL all-reduce of the packed
    [gradients | gstf | misfit] buffer -- the L-BFGS trace of the reference's experiment 001 must be the single-process one."""
    import re
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tool = os.path.join(root, "tools", "main_fwi_anomaly.py")

    def run(prefix, exp):
        for extra in (["--generate_data"], ["--nIter", "2"]):
            p = subprocess.run(prefix + [tool, "--problem", "001", "--exp_name", exp] + extra, capture_output=True, text=True, timeout=600)
            assert p.returncode == 0, p.stderr[-2000:]
        return [float(m) for m in re.findall(r"At iterate\s+\d+\s+f= ([0-9.E+-]+)", p.stdout)]

    one = run([sys.executable], str(tmp_path / "one"))
    two = run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", "29533"], str(tmp_path / "two"))
    assert len(one) == 3 and len(two) == 3
    # the sum over shots is associated differently (per-rank partial sums), so agreement is to rounding, not bit-exact
    assert np.allclose(one, two, rtol=2e-5), (one, two)


# ---------------------------------------------------------------------------------------------
# every parameterisation module, driven through the op with CUDA tensors
@pytest.mark.parametrize("kind,f0,g0", [("002", 1.08179e4, 1.68776), ("003", 3.52002e4, 4.02957)])
def test_reference_notebook_iterate0_lame_and_impedance_modules_on_cuda(tmp_path, kind, f0, g0):
    """The reference's other two published logs: notebooks 002 (FWI_Lame_Den) and 003 (FWI_IP_IS_Den),
        At iterate 0  f = 1.08179D+04 |proj g| = 1.68776   /   f = 3.52002D+04 |proj g| = 4.02957
    with the model, the source and the shot ids living on the GPU: f(x0) to the 6 printed digits; |proj g| carries the
    reference's residual-injection race (181 adjacent receivers) and is only bracketed."""
    import torch
    from sepfwi import drivers, fwi_ops
    prob = drivers.anomaly_problem(kind)
    prob.stf = prob.stf.cuda()
    vals = {}
    for compat in (True, False):
        files = drivers.write_files(prob, str(tmp_path / ("c%d" % compat)), ref_race_compat=compat or None)
        drivers.generate_data(prob, files, device="cuda")
        fwi = drivers.build_fwi(prob, files, device="cuda")
        loss = fwi(prob.shot_ids, ngpu=1)
        loss.backward()
        grads = [getattr(fwi, n).grad for n in fwi.NAMES]
        assert loss.is_cuda and all(g.is_cuda and g.shape == (prob.nz, prob.nx) for g in grads)
        vals[compat] = (loss.item(), max(float(g.abs().max()) for g in grads))
    print("experiment %s: f(x0) = %.6e (log %.5e); |proj g| = %.5f race-compat, %.5f race-free (log %.5f)"
          % (kind, vals[True][0], f0, vals[True][1], vals[False][1], g0))
    assert abs(vals[True][0] - f0) <= 2e-5 * f0 and vals[False][0] == vals[True][0]
    assert min(abs(vals[True][1] - g0), abs(vals[False][1] - g0)) <= 0.03 * g0
    fwi_ops.clear_cache()


def test_every_parameterisation_module_on_cuda_tensors(tmp_path):
    """FWI, FWI_Lame_Den, FWI_IP_IS_Den, FWI_Vp_Vs_IP, FWI_Vp_Vs_IS, FWI_Rock_Physics_VRH / _gassmann through the op on the
    GPU: equivalent parameters give the same misfit, and every module's gradient is the chain rule of its to_lame map applied
    to the (lambda, mu, rho) gradient the op returns (vector-Jacobian product by autograd, checked against the module)."""
    import torch
    from sepfwi import FWI_ops as F, fwi_ops, fwi_utils as ft
    prob = problems.tiny()
    work = str(tmp_path)
    para, survey, data = work + "/para.json", work + "/survey.json", work + "/d"
    ft.paraGen(prob.nz, prob.nx, prob.dz, prob.dx, prob.nSteps, prob.dt, prob.f0, prob.nPml, prob.nPad, para, survey, data)
    ft.surveyGen(prob.z_src, prob.x_src, prob.z_rec, prob.x_rec, survey)
    dev = "cuda"
    T = lambda a: torch.tensor(np.ascontiguousarray(a, np.float32), device=dev)
    stf, ids = T(prob.stf), torch.arange(prob.nshots, dtype=torch.int32, device=dev)
    fwi_ops.obscalc(*map(T, prob.true), stf, 1, ids, para)
    nzo, nxo, P0 = prob.nz_orig, prob.nx_orig, prob.nPml
    crop = lambda a: np.ascontiguousarray(a[P0:P0 + nzo, P0:P0 + nxo])
    lam, mu, den = (crop(a).astype(np.float64) for a in prob.start)
    vp, vs = np.sqrt((lam + 2 * mu) * 1e6 / den), np.sqrt(mu * 1e6 / den)
    opt = dict(nz=nzo, nx=nxo, nz_orig=nzo, nx_orig=nxo, nPml=P0, nPad=prob.nPad, para_fname=para)
    cases = {"FWI": (F.FWI, (vp, vs, den)), "FWI_Lame_Den": (F.FWI_Lame_Den, (lam, mu, den)),
             "FWI_IP_IS_Den": (F.FWI_IP_IS_Den, (vp * den / 1e3, vs * den / 1e3, den)),
             # these two maps carry no 1e6 (FWI_ops.py:324-328, 390-392, reproduced): velocities in km/s give lambda, mu in MPa
             "FWI_Vp_Vs_IP": (F.FWI_Vp_Vs_IP, (vp / 1e3, vs / 1e3, vp / 1e3 * den)),
             "FWI_Vp_Vs_IS": (F.FWI_Vp_Vs_IS, (vp / 1e3, vs / 1e3, vs / 1e3 * den))}
    losses, base = {}, None
    for name, (cls, fields) in cases.items():
        th = [torch.tensor(a, dtype=torch.float32, device=dev, requires_grad=True) for a in fields]
        m = cls(*th, stf, opt)
        loss = m(ids, ngpu=1)
        loss.backward()
        losses[name] = loss.item()
        grads = [getattr(m, n).grad for n in m.NAMES]
        assert all(g is not None and g.is_cuda and float(g.abs().max()) > 0 for g in grads), name
        # chain rule: the op's (lambda, mu, rho) gradient pulled back through to_lame by autograd
        with torch.no_grad():
            out = fwi_ops.backward(*cls.to_lame(*m._masked()), stf, 1, ids, para)
        x = [t.detach().clone().requires_grad_(True) for t in th]
        pads = ft.padding(*x, nzo, nxo, nzo, nxo, P0, prob.nPad)
        L = cls.to_lame(*pads)
        torch.autograd.backward(L, [g.to(dev) for g in out[1:4]])
        for a, b in zip(grads, x):
            assert rel_l2(a.cpu().numpy(), b.grad.cpu().numpy()) < 1e-5, name
    for name, v in losses.items():
        assert abs(v - losses["FWI"]) <= 2e-4 * abs(losses["FWI"]), (name, v, losses["FWI"])
    # rock physics: porosity / clay / saturation fields whose velocities respect the CFL limit of this grid
    rng = np.random.default_rng(4)
    phi = 0.25 + 0.05 * problems.smooth(rng.uniform(-1, 1, (nzo, nxo)), 4)
    cc = 0.3 + 0.1 * problems.smooth(rng.uniform(-1, 1, (nzo, nxo)), 4)
    sw = 0.6 + 0.2 * problems.smooth(rng.uniform(-1, 1, (nzo, nxo)), 4)
    for cls in (F.FWI_Rock_Physics_VRH, F.FWI_Rock_Physics_gassmann):
        th = [torch.tensor(a, dtype=torch.float32, device=dev, requires_grad=True) for a in (phi, cc, sw)]
        m = cls(*th, stf, opt)
        loss = m(ids, ngpu=1)
        loss.backward()
        assert loss.is_cuda and np.isfinite(loss.item()) and loss.item() > 0
        for n in m.NAMES:
            g = getattr(m, n).grad
            assert g is not None and g.is_cuda and torch.isfinite(g).all() and float(g.abs().max()) > 0, (cls.__name__, n)
    fwi_ops.clear_cache()
