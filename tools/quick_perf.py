"""Quick performance probe used while tuning kernels (not the bench contract).
    python tools/quick_perf.py [kernels] [nt]
Prints device loop time per time step for a few workloads, and the per-kernel profile on the
HBM-bound 8000x2000 grid."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sep-2023_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import bench
from sepfwi.engine import Propagator, ShotSpec

kernels = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 401
dev = torch.device("cuda", 0)


def run(name, batch, grad, nsteps, profile=False):
    w = bench.workload(name)
    w["stf"] = w["stf"][:nsteps]
    with Propagator(w["nz"], w["nx"], w["nPml"], w["nPad"], nsteps, w["dz"], w["dx"], w["dt"], w["f0"], max_batch=batch,
                    max_nrec=len(w["xrec"]), with_adjoint=grad, device=0, kernels=kernels) as P:
        shots = bench.make_shots(w, ShotSpec, batch)
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["true"]])
        obs = [o["ett"] for o in P.forward(shots, comps=("ett",), device_out=True)]
        for rep in range(2):
            if profile and rep == 1:
                P.set_profile(nsteps)
            if grad:
                P.set_model(*[torch.from_numpy(a).to(dev) for a in w["start"]])
                P.gradient(shots, obs, device=True)
            else:
                P.forward(shots, comps=("ett",), device_out=True)
        f, b = P.last_timing()
        cells = w["live"] * batch
        msg = "%-4s batch %d %s: fwd %.2f us/step (%.3e cell-upd/s)" % (name, batch, "grad" if grad else "fwd ", 1e3 * f / (nsteps - 1), cells * (nsteps - 1) / (f * 1e-3))
        if grad:
            msg += "  bwd %.2f us/step (%.3e cell-steps/s)  -> %.2f shot-grad/s at nt=4001" % (1e3 * b / (nsteps - 1), cells * (nsteps - 1) / (b * 1e-3), batch / ((f + b) * 1e-3 * 4000 / (nsteps - 1)))
        print(msg)
        if profile:
            pr = P.profile()
            print("     per-kernel us:", {k: round(1e3 * ms / n, 1) for k, (ms, n) in pr.items()})
            fk = [k for k in ("stream_fwd", "stress_fwd", "velocity_fwd") if k in pr]
            bk = [k for k in ("stream_bwd", "stream_recon", "stream_adj", "velocity_bwd", "stress_bwd", "velocity_adj", "stress_adj") if k in pr]
            tf = sum(pr[k][0] / pr[k][1] for k in fk) * 1e-3
            if tf > 0: print("     forward step: %.1f us -> %.0f GB/s algorithmic (52 B/cell) = %.2f of 6451" % (tf * 1e6, 52 * w["live"] / tf / 1e9, 52 * w["live"] / tf / 1e9 / 6451.2))
            if bk:
                tb = sum(pr[k][0] / pr[k][1] for k in bk) * 1e-3
                ab = (52 * w["live"] + 64 * w["interior"]) / tb / 1e9
                print("     backward step: %.1f us -> %.0f GB/s algorithmic = %.2f of 6451" % (tb * 1e6, ab, ab / 6451.2))


print("lib:", os.environ.get("SEPFWI_LIB", "default"), "kernels", kernels, "TMA", os.environ.get("SEPFWI_TMA", "0"), flush=True)
which = sys.argv[3].split(",") if len(sys.argv) > 3 else ["c2", "c2x8", "c3", "c3x8", "ref", "c5s"]
if "c2" in which: run("c2", 1, False, nt)
if "c2x8" in which: run("c2", 8, False, nt)
if "c3" in which: run("c3", 1, True, nt, profile=True)
if "c3x8" in which: run("c3", 8, True, nt, profile=True)
if "c3x16" in which: run("c3", 16, True, nt, profile=True)
if "c3x32" in which: run("c3", 32, True, nt, profile=True)
if "c3x64" in which: run("c3", 64, True, nt, profile=True)
if "ref" in which: run("ref", 19, True, nt, profile=True)
if "c5s" in which: run("c5s", 1, True, 60, profile=True)
