"""Instruction-mix digest of the hot kernels of libsepfwi.so (cuobjdump -sass): per kernel the code size, the opcode classes and
the innermost loop (the backward branch with the largest body) -- what the streaming rows compile to.
    python tools/sass_digest.py [lib] > profiles/r02_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sep-2023_b200", "libsepfwi.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
CLASSES = [("fp32 fma/add/mul", r"^(FFMA|FADD|FMUL)"), ("fp32 other (mufu, sel, min/max, cvt)", r"^(MUFU|FSEL|FMNMX|F2F|I2F|F2I|FSETP|FCHK)"),
           ("warp shuffle", r"^SHFL"), ("shared load (LDS)", r"^LDS"), ("shared store (STS)", r"^STS"),
           ("async copy global->shared (LDGSTS)", r"^LDGSTS"), ("cp.async group / wait (LDGDEPBAR, DEPBAR)", r"^(LDGDEPBAR|DEPBAR)"),
           ("global load (LDG)", r"^LDG"), ("global store (STG)", r"^STG"), ("local (spill) LDL/STL", r"^(LDL|STL)"),
           ("integer / address (IMAD, IADD3, LEA, LOP3, SHF, ISETP, ...)", r"^(IMAD|IADD|LEA|LOP3|SHF|ISETP|IABS|IMNMX|VIADD|VIMNMX|PRMT|SEL|PLOP3|P2R|R2P)"),
           ("register move (MOV)", r"^(MOV|UMOV)"), ("branch / control", r"^(BRA|BSSY|BSYNC|EXIT|CALL|RET|WARPSYNC|BAR|NANOSLEEP|YIELD|BREAK|BMOV)"),
           ("uniform datapath (U*)", r"^U[A-Z]"), ("memory fence / cache control", r"^(MEMBAR|CCTL|ERRBAR|CGAERRBAR|FENCE)"),
           ("tcgen05 / TMA (UTMA*, UTCMMA, SYNCS)", r"^(UTMA|UTC|SYNCS|UBLKCP)")]
kern = None
code = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        code[kern] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if m and kern:
        ins = re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())
        code[kern].append((int(m.group(1), 16), ins))
want = [k for k in code if re.search(r"k_stream_|k_resident_fwdILi10|k_dft_", k)]
print("SASS digest of %s (sm_100a; cuobjdump -sass; 16 bytes per instruction)\n" % os.path.basename(lib))
for k in want:
    ins = code[k]
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
    print("=" * 110)
    print("%s : %d instructions = %.1f KB" % (name, len(ins), len(ins) * 16 / 1024.0))
    cnt = collections.Counter()
    for _, i in ins:
        op = i.split()[0] if i else "?"
        for cname, rx in CLASSES:
            if re.match(rx, op):
                cnt[cname] += 1
                break
        else:
            cnt["other (" + op.split(".")[0] + ")"] += 1
    for cname, n in cnt.most_common():
        print("    %-62s %6d  %5.1f %%" % (cname, n, 100.0 * n / max(1, len(ins))))
    # loops: backward branches
    addr = {a: j for j, (a, _) in enumerate(ins)}
    loops = []
    for j, (a, i) in enumerate(ins):
        m = re.match(r"BRA(?:\.\S+)?\s+(?:\S+,\s*)?`\(\.L_x_\d+\)|BRA(?:\.\S+)?\s+.*0x([0-9a-f]+)", i)
        m2 = re.search(r"0x([0-9a-f]+)\s*$", i) if i.startswith("BRA") else None
        if m2:
            t = int(m2.group(1), 16)
            if t in addr and addr[t] < j:
                loops.append((j - addr[t] + 1, addr[t], j))
    loops.sort(reverse=True)
    for n, a0, a1 in loops[:3]:
        body = collections.Counter()
        for _, i in ins[a0:a1 + 1]:
            op = i.split()[0]
            for cname, rx in CLASSES:
                if re.match(rx, op):
                    body[cname] += 1
                    break
            else:
                body["other"] += 1
        top = ", ".join("%s %d" % (c.split(" (")[0], v) for c, v in body.most_common(7))
        print("    loop of %5d instructions (%.1f KB) at 0x%x: %s" % (n, n * 16 / 1024.0, ins[a0][0], top))
