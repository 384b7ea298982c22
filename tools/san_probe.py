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

# interior (branch-free) warps with their lead-in loops, several strips and chunks: the medium grid, streaming kernels
pm = problems.medium()
pm.nSteps = min(nsteps, 16)
pm.stf = pm.stf[:, :pm.nSteps]
with make_prop(Propagator, pm, max_batch=2, with_adjoint=True, kernels=3) as P:
    P.set_model(*pm.true)
    shots = cuda_shots(pm, ShotSpec)
    fwd = P.forward(shots)
    P.set_model(*pm.start)
    g = P.gradient(shots, [f["ett"] for f in fwd])
    print("medium grid, kernels 3: misfit %.6e" % g["misfit"], "launches", P.launches)

# round 2: the data-side operators, the sponge-flavour streaming kernel and the TMA operand path under the sanitizer too
from sepfwi import _lib
nrec = len(prob.x_rec)
ws = np.full(nrec, 0.005, np.float32); we = np.full(nrec, 0.03, np.float32); wt = np.linspace(0.5, 1.5, nrec).astype(np.float32)
with make_prop(Propagator, prob, max_batch=2, with_adjoint=True, kernels=3) as P:
    P.set_model(*prob.true)
    shots = cuda_shots(prob, ShotSpec)
    for sh in shots:
        sh.win_start, sh.win_end, sh.trace_weights, sh.src_weight = ws, we, wt, 1.1
    fwd = P.forward(shots)
    P.set_model(*prob.start)
    for opts in (dict(if_win=True, filter=[20.0, 40.0, 120.0, 200.0]), dict(if_cross_misfit=True), dict(if_src_update=True)):
        P.set_data_options(**opts)
        g = P.gradient(shots, [f["ett"] for f in fwd])
        print("data options", opts, "misfit %.6e" % g["misfit"])
nd = prob.nPml
with Propagator(prob.nz - prob.nPad, prob.nx, nd, 0, nsteps, prob.dz, prob.dx, prob.dt, prob.f0, flavour=_lib.FLAVOUR_SPONGE, max_batch=2,
                max_nrec=nrec, device=0) as P:
    lam, mu, rho = (np.ascontiguousarray(a[:prob.nz - prob.nPad]) for a in prob.true)
    P.set_model(lam * 1e6, mu * 1e6, rho)
    out = P.forward(cuda_shots(prob, ShotSpec), comps=("pr", "vx", "vz", "ett", "exx", "ezz", "exz"))
    print("sponge flavour: max |ett| %.3e" % np.abs(out[0]["ett"]).max(), "launches", P.launches)
os.environ["SEPFWI_TMA"] = "1"
with make_prop(Propagator, problems.medium(), max_batch=1, kernels=3) as P:
    m = problems.medium()
    m.stf = m.stf[:, :nsteps]
P2 = Propagator(m.nz, m.nx, m.nPml, m.nPad, nsteps, m.dz, m.dx, m.dt, m.f0, max_batch=1, max_nrec=len(m.x_rec), device=0, kernels=3)
P2.set_model(*m.true)
o = P2.forward(cuda_shots(m, ShotSpec, [0]))
print("TMA operand path: max |vx| %.3e" % np.abs(o[0]["vx"]).max(), "launches", P2.launches)
P2.close()
