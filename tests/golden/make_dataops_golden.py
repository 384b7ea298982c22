"""Golden vectors of the data-side operators from the reference's OWN kernels (Src/utilities.cu:733-1356, run through
oracle/ref_dataops_shim.cu inside oracle/_ref/libcufd_ref.so, in the order of the commented call sites Src/libCUFD.cu:353-457).
Run on the GPU box:

    gpurun -- python tests/golden/make_dataops_golden.py       # writes gpurun_out/golden/dataops_ref.npz

then copy the file into tests/golden/ and commit it.  Inputs are seeded (tests/problems.py::dataops_case) and stored with the outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import problems  # noqa: E402
from oracle import ref_dataops  # noqa: E402


def main():
    out = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    c = problems.dataops_case()
    store = dict(obs=c["obs"], cal=c["cal"], src=c["src"], dt=np.float32(c["dt"]), win_start=c["win_start"], win_end=c["win_end"],
                 weights=c["weights"], src_weight=np.float32(c["src_weight"]))
    # single operators
    store["op_bp"] = ref_dataops.bp_filter(c["obs"], c["dt"], c["filt"])
    store["op_win"] = ref_dataops.window_traces(c["obs"], c["dt"], c["win_start"], c["win_end"], c["weights"], c["src_weight"], 0.005)
    store["op_win_simple"] = ref_dataops.window_simple(c["obs"], c["dt"], 0.005)
    store["op_normfact"] = ref_dataops.normfact(c["obs"], c["cal"])
    for name, opts in problems.DATAOPS_CASES.items():
        kw = dict(opts)
        if kw.pop("filter", False):
            kw["filt"] = c["filt"]
        r = ref_dataops.condition(c["obs"], c["cal"], c["src"], c["dt"], win_start=c["win_start"], win_end=c["win_end"], weights=c["weights"],
                                  src_weight=c["src_weight"], **kw)
        for k in ("res", "cal", "obs", "src"):
            store["%s_%s" % (name, k)] = r[k]
        store["%s_misfit" % name] = np.float32(r["misfit"])
        store["%s_amp_ratio" % name] = np.float32(r["amp_ratio"])
        print(name, "misfit %.6e" % r["misfit"], "amp_ratio %.5f" % r["amp_ratio"], "|res| %.4e" % np.abs(r["res"]).max())
    np.savez_compressed(os.path.join(out, "dataops_ref.npz"), **store)
    print("wrote", os.path.join(out, "dataops_ref.npz"))


if __name__ == "__main__":
    main()
