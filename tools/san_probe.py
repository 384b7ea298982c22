"""Small forward + gradient through the default path (resident forward loop + streaming reverse-time kernels) and through the
streaming-only path, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/san_probe.py [nSteps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sep-2023_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import problems
from sepfwi.engine import Propagator, ShotSpec
from util import cuda_shots, make_prop

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
prob = problems.tiny()
prob.nSteps = nsteps
prob.stf = prob.stf[:, :nsteps]
for kern in (0, 3):
    with make_prop(Propagator, prob, max_batch=2, with_adjoint=True, kernels=kern) as P:
        P.set_model(*prob.true)
        shots = cuda_shots(prob, ShotSpec)
        fwd = P.forward(shots)
        P.set_model(*prob.start)
        g = P.gradient(shots, [f["ett"] for f in fwd])
        print("kernels", kern, "misfit %.6e" % g["misfit"], "resident launches", P.resident_launches, "launches", P.launches)
