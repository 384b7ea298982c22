import os, sys
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "sep-2023_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch, bench
from sepfwi.engine import Propagator, ShotSpec
name = sys.argv[1]; nsteps = int(sys.argv[2]); batch = int(sys.argv[3])
w = bench.workload(name); w["stf"] = w["stf"][:nsteps]
dev = torch.device("cuda", 0)
with Propagator(w["nz"], w["nx"], w["nPml"], w["nPad"], nsteps, w["dz"], w["dx"], w["dt"], w["f0"], max_batch=batch, max_nrec=len(w["xrec"]), with_adjoint=False, device=0, kernels=0) as P:
    shots = bench.make_shots(w, ShotSpec, batch)
    P.set_model(*[torch.from_numpy(a).to(dev) for a in w["true"]])
    for rep in range(2): P.forward(shots, comps=("ett",), device_out=True)
    f, b = P.last_timing()
    print(name, "batch", batch, "LZ", os.environ.get("SEPFWI_LZ"), "FORCE", os.environ.get("SEPFWI_FORCE"), "fwd %.2f us/step" % (1e3 * f / (nsteps - 1)))
