"""CPU tests of the host layer (no compute calls: there is no GPU in the build container).

 * libsepfwi.so loads and exports every symbol include/sepfwi.h declares;
 * without a CUDA device every compute entry fails loudly (no CPU fallback);
 * para/survey JSON writers, padding rule, Ricker, survey parser (Src_Rec index shift);
 * shot sharding + the packed all-reduce on two gloo ranks;
 * the product package never imports the oracle.
"""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from sepfwi import _lib
    hdr = open(os.path.join(ROOT, "include", "sepfwi.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sepfwi_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (sepfwi_[a-z0-9_]+)", out))
    assert declared <= exported
    assert L.sepfwi_version() >= 100


def test_struct_layouts_match_header():
    """ctypes mirrors of sepfwi_params / sepfwi_shot have the C sizes (checked against a compiled probe)."""
    import ctypes as C
    from sepfwi import _lib
    src = '#include <stdio.h>\n#include "sepfwi.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(sepfwi_params), sizeof(sepfwi_shot), sizeof(sepfwi_data_options));return 0;}\n'
    exe = os.path.join(ROOT, "tests", "_probe_sizes")
    p = subprocess.run(["/usr/bin/gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src, text=True,
                       capture_output=True)
    assert p.returncode == 0, p.stderr
    a, b, c = map(int, subprocess.run([exe], capture_output=True, text=True).stdout.split())
    os.unlink(exe)
    assert C.sizeof(_lib.Params) == a and C.sizeof(_lib.Shot) == b and C.sizeof(_lib.DataOptions) == c


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sepfwi._lib import SepfwiError
    from sepfwi.engine import Propagator
    with pytest.raises(SepfwiError) as e:
        Propagator(64, 64, 8, 0, 10, 10.0, 10.0, 1e-3, 10.0)
    assert "no CPU fallback" in str(e.value) or e.value.code == -2
    # the cufd drop-in reports through its return code as well
    import ctypes as C
    from sepfwi import _lib
    z = np.zeros((4, 4), np.float32)
    ids = np.zeros(1, np.int32)
    rc = _lib.lib().sepfwi_cufd(None, None, None, None, None, z.ctypes.data, z.ctypes.data, z.ctypes.data, z.ctypes.data, 2, 0, 1,
                                ids.ctypes.data, b"/nonexistent/para.json")
    assert rc == -4 and b"parameter file" in _lib.lib().sepfwi_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sep-2023_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "README.md", os.path.join(dirpath, f)


def test_para_and_survey_files(tmp_path):
    from sepfwi import fwi_utils as ft
    para, survey, data = str(tmp_path / "para.json"), str(tmp_path / "survey.json"), str(tmp_path / "data")
    nz, nx, nPad = ft.padded_shape(101, 201, 32)
    assert (nz, nx, nPad) == (192, 265, 27)
    assert ft.padded_shape(64, 96, 16)[2] == 32                # already aligned -> a full 32
    ft.paraGen(nz, nx, 20.0, 20.0, 1501, 0.002, 10.0, 32, nPad, para, survey, data)
    ft.surveyGen(np.full(3, 1), np.array([10, 20, 30]), np.full(5, 95), np.arange(10, 15), survey, Src_rxz=[1.0, 0.5, 2.0])
    for fn in (para, survey):
        txt = open(fn).read()
        assert "\n" not in txt                                  # the C++ readers only take the first line
        json.loads(txt)
    p = ft.read_json_first_line(para)
    assert p["nPoints_pml"] == 32 and p["nSteps"] == 1501 and p["survey_fname"] == survey and os.path.isdir(data)
    sv = ft.load_survey(survey, [2, 0], 32)
    assert sv[0]["xs"] == 30 + 32 and sv[0]["zs"] == 1 + 32 and sv[0]["src_rxz"] == 2.0
    assert np.array_equal(sv[1]["xrec"], np.arange(10, 15) + 32) and np.all(sv[1]["zrec"] == 95 + 32)


def test_ricker_and_padding():
    import torch
    from sepfwi import fwi_utils as ft
    s = ft.sourceGene(10.0, 1501, 0.002)
    k = int(round(1.2 / 10.0 / 0.002))
    assert s.argmax() == k and s[k] == pytest.approx(1.0e7)
    a = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    p = ft.padding(a, a, a, 3, 4, 3, 4, 2, 1)[0]
    assert tuple(p.shape) == (3 + 4 + 1, 4 + 4)
    assert torch.equal(p[2:5, 2:6], a) and torch.all(p[0, 2:6] == a[0]) and torch.all(p[-1, 2:6] == a[-1])
    assert np.array_equal(ft.padding_numpy_array(a.numpy(), 2, 1), p.numpy())


def test_shard_bounds_and_errors():
    from sepfwi import dist
    assert dist.shard_bounds(19, 4) == [0, 4, 9, 14, 19]       # Torch_Fwi.cpp:59-60
    assert dist.shard(list(range(19)), 4, 1) == [4, 5, 6, 7, 8]
    with pytest.raises(RuntimeError):
        dist.shard([0, 1], 3, 0)                               # more GPUs than shots (Torch_Fwi.cpp:49-52)


_WORKER = r'''
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "sep-2023_b200"))
import torch, torch.distributed as td
from sepfwi import dist
rank = int(sys.argv[1])
td.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=2)
ids = dist.shard(list(range(5)), 2, rank)                       # 2 + 3 shots
assert ids == ([0, 1] if rank == 0 else [2, 3, 4]), ids
g = [torch.full((3, 4), float(rank + 1) * k) for k in (1, 2, 3)]
gstf = torch.zeros(5, 7)
for i in ids:
    gstf[i] = i + 1
J, gl, gm, gd, gs = dist.allreduce_gradients(10.0 * (rank + 1), g[0], g[1], g[2], gstf)
assert abs(J - 30.0) < 1e-6 and torch.all(gl == 3) and torch.all(gm == 6) and torch.all(gd == 9)
assert torch.equal(gs[:, 0], torch.arange(1, 6, dtype=torch.float32))   # every shot's stf row survives the reduction
td.destroy_process_group()
print("ok", rank)
'''


def test_packed_allreduce_on_two_gloo_ranks(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % dict(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    for r, p in enumerate(procs):
        out, err = p.communicate(timeout=120)
        assert p.returncode == 0 and "ok %d" % r in out, err[-2000:]


def test_fwi_module_parameterisations_map_to_lame():
    """The parameterisation maps (FWI_ops.py:124-125, 263-266, 324-328, 390-392) without running the op."""
    import torch
    from sepfwi import FWI_ops as F
    vp, vs, den = torch.tensor([3000.0]), torch.tensor([1700.0]), torch.tensor([2200.0])
    lam, mu, d = F.FWI.to_lame(vp, vs, den)
    assert mu.item() == pytest.approx(1700.0 ** 2 * 2200.0 / 1e6) and lam.item() == pytest.approx((3000.0 ** 2 - 2 * 1700.0 ** 2) * 2200.0 / 1e6)
    ip, is_ = vp * den / 1e3, vs * den / 1e3                    # impedances scaled so that IP^2/Den is in MPa
    l2, m2, _ = F.FWI_IP_IS_Den.to_lame(ip, is_, den)
    assert l2.item() == pytest.approx(lam.item(), rel=1e-6) and m2.item() == pytest.approx(mu.item(), rel=1e-6)
    l3, m3, d3 = F.FWI_Vp_Vs_IP.to_lame(vp, vs, vp * den)
    assert d3.item() == pytest.approx(2200.0) and m3.item() == pytest.approx(1700.0 ** 2 * 2200.0)
    l4, m4, d4 = F.FWI_Vp_Vs_IS.to_lame(vp, vs, vs * den)
    assert d4.item() == pytest.approx(2200.0) and l4.item() == pytest.approx(l3.item(), rel=1e-6) and m4.item() == pytest.approx(m3.item(), rel=1e-6)
    assert F.FWI_Lame_Den.to_lame(lam, mu, den) == (lam, mu, den)


def test_rock_physics_maps_match_reference_golden(golden_dir):
    """(PHI, CC, SW) -> (Lambda, Mu, Den) of FWI_Rock_Physics_VRH / _gassmann against vectors produced by executing the
    reference's own statements (tests/golden/make_rockphys_golden.py; FWI_ops.py:451-507, 567-619).  fp32 on both sides."""
    import torch
    from sepfwi import FWI_ops as F
    g = np.load(os.path.join(golden_dir, "rockphys.npz"))
    phi, cc, sw = (torch.from_numpy(g[k]) for k in ("PHI", "CC", "SW"))
    for key, cls in (("vrh", F.FWI_Rock_Physics_VRH), ("gassmann", F.FWI_Rock_Physics_gassmann)):
        lam, mu, den = cls.to_lame(phi, cc, sw)
        for name, mine in (("Lambda", lam), ("Mu", mu), ("Den", den)):
            ref = g["%s_%s" % (key, name)]
            scale = np.abs(ref).max()
            assert np.abs(mine.numpy() - ref).max() <= 2e-6 * scale, (key, name)


class _ToyFWI(object):
    pass


def test_pytorch_objective_drives_scipy_lbfgs():
    """PyTorchObjective (Ops/FWI/obj_wrapper.py:10-97 interface): x0 / bounds packing, one evaluation per new x,
    gradients handed to scipy's L-BFGS-B exactly as the reference drivers do (Main-001-...py:157-168)."""
    import torch
    from scipy import optimize
    from sepfwi.obj_wrapper import PyTorchObjective

    class Quad(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.A = torch.nn.Parameter(torch.zeros(3, 4))
            self.B = torch.nn.Parameter(torch.ones(5))
            self.register_buffer("A_ref", torch.zeros(3, 4))
            self.Bounds = {"B": (np.full(5, 0.5), np.full(5, 4.0))}

        def forward(self):
            ta = torch.arange(12, dtype=torch.float32).reshape(3, 4)
            return ((self.A - ta) ** 2).sum() + ((self.B - 0.25) ** 2).sum()

    m = Quad()
    obj = PyTorchObjective(m, lambda: m())
    assert obj.x0.dtype == np.float64 and obj.x0.size == 17 and list(obj.param_shapes) == ["A", "B"]
    assert obj.bounds.lb[:12].max() == -np.inf and np.all(obj.bounds.lb[12:] == 0.5) and np.all(obj.bounds.ub[12:] == 4.0)
    f0 = obj.fun(obj.x0)
    g0 = obj.jac(obj.x0)
    assert obj.nfev == 1                                     # fun + jac of the same x: one evaluation
    assert f0 == pytest.approx(sum(k * k for k in range(12)) + 5 * 0.75 ** 2)
    assert np.allclose(g0[:12], -2.0 * np.arange(12)) and np.allclose(g0[12:], 1.5)
    r = optimize.minimize(obj.fun, obj.x0, method="L-BFGS-B", jac=obj.jac, bounds=obj.bounds,
                          options={"maxiter": 50, "maxcor": 5, "ftol": 1e-14, "gtol": 1e-10})
    assert np.allclose(r.x[:12], np.arange(12), atol=1e-4) and np.allclose(r.x[12:], 0.5)    # B sits on its lower bound
    assert np.allclose(m.B.detach().numpy(), obj.unpack_parameters(obj.cached_x)["B"].numpy())


def test_resident_planner_invariants_over_many_grids():
    """sepfwi_plan_resident (host arithmetic only): for every grid it accepts, the tiles cover the live grid, fit the SMs, the
    per-tile shared memory fits the device limit, no tile spans both sides of a CPML strip pair (it holds the memory variables of
    one z side and one x side only), and the batch is split into whole launches.  Known plans: C2 -> 19 x 7 tiles of 67 rows, rows
    per thread 10; C3 and the 8000 x 2000 grid do not fit 148 SMs; the reference-size grid takes several shots per launch."""
    import ctypes as C
    from sepfwi import _lib
    L = _lib.lib()

    def plan(nz, nx, nPml, nPad, nshots, nsm=148, smem=232448, kernels=0):
        p = _lib.Params(nz, nx, nPml, nPad, 100, 10.0, 10.0, 1e-3, 10.0, 0, 0, nshots, 1, 0, kernels, 0)
        out = (C.c_int * 5)()
        assert L.sepfwi_plan_resident(C.byref(p), nshots, nsm, smem, out) == 0
        return tuple(out)

    assert plan(480, 1064, 32, 16, 1) == (10, 19, 7, 67, 1)                 # C2
    assert plan(416, 1764, 32, 2, 1)[0] == 0 and plan(2080, 8064, 32, 32, 1)[0] == 0
    rpt, ntx, ntz, orows, per = plan(192, 265, 32, 32, 19)                  # the reference's experiment, 19 shots
    assert rpt > 0 and per >= 4 and ntx == 5
    assert plan(480, 1064, 32, 16, 1, kernels=3)[0] == 0                    # streaming kernels forced
    assert plan(480, 1064, 32, 16, 1, nsm=64)[0] == 0                       # a smaller device cannot co-schedule the tiles
    assert plan(480, 1064, 32, 16, 1, smem=100 * 1024)[0] == 0              # ... nor one with little shared memory
    rng = np.random.default_rng(1)
    accepted = 0
    for _ in range(400):
        nPml = int(rng.choice([8, 16, 32]))
        nzo, nxo = int(rng.integers(12, 500)), int(rng.integers(12, 1500))
        nPad = int(32 - (nzo + 2 * nPml) % 32)
        nz, nx, nb = nzo + 2 * nPml + nPad, nxo + 2 * nPml, int(rng.integers(1, 20))
        rpt, ntx, ntz, orows, per = plan(nz, nx, nPml, nPad, nb)
        if rpt == 0:
            continue
        accepted += 1
        nzA, ER = nz - nPad, 8 * rpt
        assert 3 <= rpt <= 13 and 4 <= orows <= ER - 8
        assert ntx * 56 >= nx and (ntx - 1) * 56 < nx and ntz * orows >= nzA and (ntz - 1) * orows < nzA
        assert 1 <= per <= nb and ntx * ntz * per <= 148
        assert 4 * (5 * ER * 64 + 4 * 32 * 64 + 4 * ER * 32 + 6 * ER) <= 232448
        for t in range(ntz):
            lo, hi = max(t * orows - 4, 0), min(t * orows - 4 + ER - 1, nzA - 1)
            assert not (lo < nPml and hi > nzA - nPml - 1), (nz, nx, nPml, t)
        for t in range(ntx):
            lo, hi = max(t * 56 - 4, 0), min(t * 56 + 59, nx - 1)
            assert not (lo < nPml and hi > nx - nPml - 1), (nz, nx, nPml, t)
    assert accepted > 50


def test_streaming_work_lists_tile_the_grid_exactly():
    """sepfwi_plan_stream (host arithmetic only): for random grids, CPML widths and batch sizes the work list of each streaming
    kernel covers every (row, 120-column strip) of the live grid (the reverse sweep: of interior + ring) exactly once, interior (branch-free) items keep the nPml + 5 margin
    from every CPML strip, edge items come first, and no chunk is empty."""
    import ctypes as C
    from sepfwi import _lib
    L = _lib.lib()
    rng = np.random.default_rng(4)
    for trial in range(120):
        nPml = int(rng.choice([8, 16, 32]))
        nzo, nxo = int(rng.integers(12, 2200)), int(rng.integers(12, 8100))
        if trial % 3 == 0:
            nzo, nxo = int(rng.integers(12, 300)), int(rng.integers(12, 600))
        nPad = int(32 - (nzo + 2 * nPml) % 32)
        nz, nx, nb = nzo + 2 * nPml + nPad, nxo + 2 * nPml, int(rng.integers(1, 20))
        nzA, nstrips = nz - nPad, (nx + 119) // 120
        p = _lib.Params(nz, nx, nPml, nPad, 100, 10.0, 10.0, 1e-3, 10.0, 0, 0, nb, 1, 0, 3, 0)
        for which in range(3):
            n = C.c_int(0)
            assert L.sepfwi_plan_stream(C.byref(p), nb, 148, which, None, 0, C.byref(n)) == 0
            buf = (C.c_int * (4 * n.value))()
            assert L.sepfwi_plan_stream(C.byref(p), nb, 148, which, buf, n.value, C.byref(n)) == 0
            it = np.frombuffer(buf, np.int32).reshape(-1, 4)
            assert np.all(it[:, 0] % 120 == 0) and np.all(it[:, 2] > it[:, 1]) and np.all(it[:, 1] >= 0) and np.all(it[:, 2] <= nzA)
            cover = np.zeros((nstrips, nzA), np.int32)
            for x0, z0, z1, edge in it:
                cover[x0 // 120, z0:z1] += 1
                if not edge:
                    assert z0 >= nPml + 5 and z1 <= nzA - nPml - 5 and x0 - 4 >= nPml + 3 and x0 + 123 <= nx - nPml - 4
            if which == 1:      # the reverse sweep: interior + 2-cell ring exactly once, nothing launched that lies wholly outside it
                zl, zh, xl, xh = nPml - 2, nzA - nPml + 2, nPml - 2, nx - 1 - nPml + 2
                live = np.array([x0 + 120 > xl and x0 <= xh for x0 in range(0, nstrips * 120, 120)])
                assert np.all(cover[live, zl:zh] == 1) and np.all(cover <= 1), (nz, nx, nPml, nb, which)
                assert np.all(np.minimum(it[:, 2], zh) > np.maximum(it[:, 1], zl)) and np.all(live[it[:, 0] // 120])
            else:
                assert np.all(cover == 1), (nz, nx, nPml, nb, which)
            first_inner = np.argmax(it[:, 3] == 0) if np.any(it[:, 3] == 0) else len(it)
            assert np.all(it[first_inner:, 3] == 0)                  # edge items first


def test_reverse_time_plan_merges_the_two_sweeps_where_the_grid_is_a_few_waves():
    """sepfwi_plan_backward (host arithmetic only): one shot of the C3 grid and the 19-shot reference experiment are planned jointly
    (latency regime: chunk heights of both sweeps chosen together) and run as ONE launch per reverse-time step; 8 shots of the C3
    grid and one shot of the 8000 x 2000 grid merge without a joint plan (beyond L2); 64 shots of the C3 grid keep two launches.
    The plan is deterministic and every work list it implies is non-empty."""
    import ctypes as C
    from sepfwi import _lib
    L = _lib.lib()

    def plan(nz, nx, nPml, nPad, nb):
        p = _lib.Params(nz, nx, nPml, nPad, 100, 10.0, 10.0, 1e-3, 10.0, 0, 0, nb, 1, 1, 0, 0)
        out = (C.c_int * 7)()
        assert L.sepfwi_plan_backward(C.byref(p), nb, 148, out) == 0
        return list(out)

    c3 = plan(448, 1764, 32, 32, 1)      # padded C3 grid: 416 live rows + 32 alignment rows
    assert c3[0] == 1 and all(v > 0 for v in c3[1:5]), c3
    assert c3 == plan(448, 1764, 32, 32, 1)
    ref = plan(224, 265, 32, 32, 19)
    assert ref[0] == 1 and all(v > 0 for v in ref[1:5]), ref
    c3x8 = plan(448, 1764, 32, 32, 8)
    assert c3x8[0] == 1 and c3x8[1:5] == [0, 0, 0, 0], c3x8
    c3x64 = plan(448, 1764, 32, 32, 64)
    assert c3x64[0] == 0, c3x64
    c5 = plan(2112, 8064, 32, 32, 1)
    assert c5[0] == 1 and c5[1:5] == [0, 0, 0, 0], c5
    for r in (c3, ref, c3x8, c3x64, c5):
        assert r[5] > 0 and r[6] > 0
