"""Debug probe: C4 geometry (vertical fiber, 2 shots, nt small) -- gradient of each kernel family vs the CPU oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sep-2023_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import bench
from oracle import oracle as O
from sepfwi.engine import Propagator, ShotSpec
from util import rel_l2

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 301
w = bench.workload("c3"); w["stf"] = w["stf"][:nt]; w["nSteps"] = nt
zrec, xrec = np.arange(10, 340), np.full(330, 850)
src = [(2, 20 + 26 * 30), (2, 20 + 26 * 34)]
par = O.make_par(w["nz"], w["nx"], w["nPml"], w["nPad"], nt, w["dz"], w["dx"], w["dt"], w["f0"], fiber=1)
survey = {i: (zs, xs, zrec, xrec) for i, (zs, xs) in enumerate(src)}
stf = np.tile(w["stf"][None, :], (2, 1)).astype(np.float32)
obs = {i: O.forward(par, *w["true"], stf[i], zs, xs, zrec, xrec, comps=("ett",))["ett"] for i, (zs, xs) in enumerate(src)}
J, gl, gm, gd, gs = O.fwi_backward(par, *w["start"], stf, 1, np.arange(2), survey, obs)
P0 = w["nPml"]
res = {}
for tag, kern, merge in (("base", 1, "0"), ("stream2", 3, "0"), ("onepass", 3, "1")):
    os.environ["SEPFWI_MERGE_BWD"] = merge
    for nshot in (2, 1):
        with Propagator(w["nz"], w["nx"], w["nPml"], w["nPad"], nt, w["dz"], w["dx"], w["dt"], w["f0"], fiber=1, max_batch=2,
                        max_nrec=len(zrec), with_adjoint=True, device=0, kernels=kern) as P:
            shots = [ShotSpec(zs + P0, xs + P0, zrec + P0, xrec + P0, stf[i]) for i, (zs, xs) in enumerate(src)][:nshot]
            P.set_model(*w["start"])
            r = P.gradient(shots, [obs[i] for i in range(nshot)])
            res[(tag, nshot)] = r
            if nshot == 2:
                print(tag, "misfit", r["misfit"], "oracle", J, "glam", rel_l2(r["glam"], gl), "gmu", rel_l2(r["gmu"], gm), "grho", rel_l2(r["grho"], gd),
                      "gstf", rel_l2(np.stack(r["gstf"]), gs), "max|glam|", np.abs(r["glam"]).max(), np.abs(gl).max(), flush=True)
for nshot in (2, 1):
    b = res[("base", nshot)]
    for tag in ("stream2", "onepass"):
        r = res[(tag, nshot)]
        print(nshot, "shots:", tag, "vs base: glam", rel_l2(r["glam"], b["glam"]), "gmu", rel_l2(r["gmu"], b["gmu"]), "grho", rel_l2(r["grho"], b["grho"]))
# where do oracle and base differ?
d = np.abs(res[("base", 2)]["glam"] - gl)
z, x = np.unravel_index(np.argmax(d), d.shape)
print("largest |base - oracle| at", z, x, d[z, x], "values", res[("base", 2)]["glam"][z, x], gl[z, x])
rows = d.max(axis=1); cols = d.max(axis=0)
print("rows with diff > 1e-3 max:", np.nonzero(rows > 1e-3 * d.max())[0][[0, -1]], "cols:", np.nonzero(cols > 1e-3 * d.max())[0][[0, -1]])
# oracle per-shot consistency: sum of single-shot gradients
J0, gl0 = O.fwi_backward(par, *w["start"], stf, 1, np.arange(1), {0: survey[0]}, {0: obs[0]})[:2]
print("oracle shot0 only vs base shot0 only: glam", rel_l2(res[("base", 1)]["glam"], gl0), "misfit", res[("base", 1)]["misfit"], J0)
