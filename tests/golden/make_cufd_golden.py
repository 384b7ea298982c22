"""Generate golden vectors from the reference's OWN CUDA path (oracle/_ref/libcufd_ref.so,
built from /root/reference by oracle/Makefile).  Run on the GPU box:

    gpurun -- python tests/golden/make_cufd_golden.py        # writes gpurun_out/golden/cufd_*.npz

then copy the .npz files into tests/golden/ and commit them.  For each problem of
tests/problems.py: observed data of the true model (cufd calc_id=2 -> Shot_*.bin), then misfit
and gradients at the start model (calc_id=1), and the misfit-only path (calc_id=0).
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sep-2023_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import problems  # noqa: E402
from oracle import ref_cufd  # noqa: E402
from sepfwi import fwi_utils as ft  # noqa: E402


def run(prob, outdir):
    work = tempfile.mkdtemp(prefix="cufd_" + prob.name)
    para, survey, data = os.path.join(work, "para.json"), os.path.join(work, "survey.json"), os.path.join(work, "data")
    ft.paraGen(prob.nz, prob.nx, prob.dz, prob.dx, prob.nSteps, prob.dt, prob.f0, prob.nPml, prob.nPad, para, survey, data)
    ft.surveyGen(prob.z_src, prob.x_src, prob.z_rec, prob.x_rec, survey)
    ids = np.arange(prob.nshots, dtype=np.int32)
    ref_cufd.cufd(2, *prob.true, prob.stf, ids, para, fiber=prob.fiber)
    store = {}
    nrec = len(prob.x_rec)
    for i in ids:
        for c in ("pr", "vx", "vz", "ett"):
            store["obs_%s%d" % (c, i)] = np.fromfile(os.path.join(data, "Shot_%s%d.bin" % (c, i)), np.float32).reshape(nrec, prob.nSteps)
    J, gl, gm, gd, gs = ref_cufd.cufd(1, *prob.start, prob.stf, ids, para, fiber=prob.fiber)
    store.update(misfit=np.float32(J), glam=gl, gmu=gm, gden=gd, gstf=gs)
    J0 = ref_cufd.cufd(0, *prob.start, prob.stf, ids, para, fiber=prob.fiber)[0]
    store["misfit_calc0"] = np.float32(J0)
    # a second gradient run quantifies the reference's own atomic-order nondeterminism
    J2, gl2, gm2, gd2, _ = ref_cufd.cufd(1, *prob.start, prob.stf, ids, para, fiber=prob.fiber)
    store["rerun_relerr"] = np.array([np.linalg.norm(a - b) / np.linalg.norm(a) for a, b in ((gl, gl2), (gm, gm2), (gd, gd2))])
    path = os.path.join(outdir, "cufd_%s.npz" % prob.name)
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes; misfit", J, J0, "rerun", store["rerun_relerr"])


if __name__ == "__main__":
    out = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    # tiny_ezz: vertical fiber through oracle/_ref/libcufd_ref_ezz.so (the reference's ezz kernels, see oracle/Makefile)
    for prob in (problems.tiny(), problems.small(), problems.small_adj(), problems.tiny(fiber=1)):
        run(prob, out)
