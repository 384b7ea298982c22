"""Short forward + gradient run for ncu (launch list / --set full captures).
    python tools/profile_step.py [workload] [nsteps] [batch] [kernels]
Runs `nsteps` time steps of forward modelling and one gradient on the workload's grid."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sep-2023_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import bench
from sepfwi.engine import Propagator, ShotSpec

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
kernels = int(sys.argv[4]) if len(sys.argv) > 4 else 0
w = bench.workload(name)
dev = torch.device("cuda", 0)
with Propagator(w["nz"], w["nx"], w["nPml"], w["nPad"], nsteps, w["dz"], w["dx"], w["dt"], w["f0"], max_batch=batch,
                max_nrec=len(w["xrec"]), with_adjoint=True, device=0, kernels=kernels) as P:
    w["stf"] = w["stf"][:nsteps]
    shots = bench.make_shots(w, ShotSpec, batch)
    P.set_model(*[torch.from_numpy(a).to(dev) for a in w["true"]])
    obs = [o["ett"] for o in P.forward(shots, comps=("ett",), device_out=True)]
    P.set_model(*[torch.from_numpy(a).to(dev) for a in w["start"]])
    r = P.gradient(shots, obs, device=True)
    torch.cuda.synchronize()
    print("misfit", r["misfit"], "launches", P.launches, "loops ms", P.last_timing())
