"""Timing experiments on the resident forward kernel: python tools/res_probe.py [workload] [nt]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sep-2023_b200"), os.path.join(ROOT, "tests")]
import torch
import bench
from sepfwi.engine import Propagator, ShotSpec

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 801
comps = tuple(sys.argv[3].split(",")) if len(sys.argv) > 3 else ("ett",)
w = bench.workload(name)
w["stf"] = w["stf"][:nt]
dev = torch.device("cuda", 0)
with Propagator(w["nz"], w["nx"], w["nPml"], w["nPad"], nt, w["dz"], w["dx"], w["dt"], w["f0"], max_batch=1,
                max_nrec=len(w["xrec"]), device=0) as P:
    shots = bench.make_shots(w, ShotSpec, 1)
    P.set_model(*[torch.from_numpy(a).to(dev) for a in w["true"]])
    for _ in range(3):
        P.forward(shots, comps=comps, device_out=True)
    f, _b = P.last_timing()
    print("%s nt %d dbg %s rpt %s comps %s: %.2f us/step  (resident launches %d)" % (
        name, nt, os.environ.get("SEPFWI_RES_DEBUG", "0"), os.environ.get("SEPFWI_RESIDENT_RPT", "auto"), ",".join(comps),
        1e3 * f / (nt - 1), P.resident_launches))
