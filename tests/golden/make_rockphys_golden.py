"""Golden vectors of the reference's rock-physics parameterisations (porosity, clay content, water saturation ->
Lambda [MPa], Mu [MPa], Den), DAS_Waveform_Inversion/Ops/FWI/FWI_ops.py:451-507 (Voigt-Reuss-Hill) and :567-619
(Gassmann).

Runs only in the build container (needs /root/reference).  The reference module cannot be imported here -- it
JIT-compiles its CUDA extension into its own (read-only) tree at import time -- so the statements of the two
`forward` methods between the masking and the `FWIFunction.apply` call are taken from the file's AST and
executed unmodified on seeded inputs.

    python tests/golden/make_rockphys_golden.py
"""
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/DAS_Waveform_Inversion/Ops/FWI/FWI_ops.py"


def forward_math(tree, cls_name):
    """Statements of cls_name.forward after the three *_mask_pad assignments and before the return."""
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name)
    fwd = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "forward")
    body = [s for s in fwd.body if not isinstance(s, ast.Return)]
    start = max(i for i, s in enumerate(body)
                if isinstance(s, ast.Assign) and getattr(s.targets[0], "id", "").endswith("_mask_pad")) + 1
    return ast.Module(body=body[start:], type_ignores=[])


def main():
    tree = ast.parse(open(SRC).read())
    rng = np.random.default_rng(2023)
    shape = (24, 40)
    phi = rng.uniform(0.05, 0.35, shape).astype(np.float32)
    cc = rng.uniform(0.0, 0.6, shape).astype(np.float32)
    sw = rng.uniform(0.0, 1.0, shape).astype(np.float32)
    out = dict(PHI=phi, CC=cc, SW=sw)
    for name in ("FWI_Rock_Physics_VRH", "FWI_Rock_Physics_gassmann"):
        ns = dict(torch=torch, PHI_mask_pad=torch.from_numpy(phi), CC_mask_pad=torch.from_numpy(cc),
                  SW_mask_pad=torch.from_numpy(sw))
        exec(compile(forward_math(tree, name), SRC, "exec"), ns)
        key = "vrh" if name.endswith("VRH") else "gassmann"
        for k in ("Lambda", "Mu", "Den"):
            out["%s_%s" % (key, k)] = ns[k].numpy()
        print(name, {k: (float(ns[k].min()), float(ns[k].max())) for k in ("Lambda", "Mu", "Den")})
    np.savez_compressed(os.path.join(HERE, "rockphys.npz"), **out)


if __name__ == "__main__":
    main()
