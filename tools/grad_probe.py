"""Gradient-loop timing probe: python tools/grad_probe.py workload nsteps batch  (env SEPFWI_LZ / SEPFWI_LZE / SEPFWI_FORCE)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sep-2023_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch, bench
from sepfwi.engine import Propagator, ShotSpec
name = sys.argv[1]; nsteps = int(sys.argv[2]); batch = int(sys.argv[3])
w = bench.workload(name); w["stf"] = w["stf"][:nsteps]
dev = torch.device("cuda", 0)
with Propagator(w["nz"], w["nx"], w["nPml"], w["nPad"], nsteps, w["dz"], w["dx"], w["dt"], w["f0"], max_batch=batch, max_nrec=max(len(w["xrec"]), 2000), with_adjoint=True, device=0, kernels=0, fiber=1 if os.environ.get("FIBER") == "v" else 0) as P:
    if os.environ.get("FIBER") == "v":      # vertical fiber as in bench.py's C4-like section: every row of one strip has injection targets
        w["zrec"], w["xrec"] = np.arange(10, w["nz"] - w["nPad"] - 2 * w["nPml"] - 10), np.full(w["nz"] - w["nPad"] - 2 * w["nPml"] - 20, w["src"][0][1])
    shots = bench.make_shots(w, ShotSpec, batch)
    P.set_model(*[torch.from_numpy(a).to(dev) for a in w["true"]])
    obs = [o["ett"] for o in P.forward(shots, comps=("ett",), device_out=True)]
    P.set_model(*[torch.from_numpy(a).to(dev) for a in w["start"]])
    for rep in range(2):
        if rep == 1: P.set_profile(nsteps)
        P.gradient(shots, obs, device=True)
    pr = P.profile()
    print(name, "batch", batch, {k: os.environ.get(k) for k in ("SEPFWI_LZ", "SEPFWI_LZE", "SEPFWI_FORCE") if os.environ.get(k)},
          {k: round(1e3 * ms / n, 1) for k, (ms, n) in pr.items()})
