"""DRAM traffic per launch of the time-stepping kernels, measured with ncu in a SEPARATE process (never the timed one).

    python tools/measure_traffic.py <workload> [nsteps] [batch] [--out file.json]

Runs `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none`
over tools/profile_step.py (a short forward + gradient on the workload's grid) for every k_stream_* / k_resident_* launch and
prints one JSON object  {kernel: {"launches": n, "dram_read": bytes per launch, "dram_write": ..., "dram_bytes": ...,
"time_us": ...}}  (means over the launches after the first five of each kernel).  Caches are NOT flushed between launches
(--cache-control none): on the grids that fit the 126 MB L2 the steady-state traffic of a time loop is what is wanted, not the
cold-cache figure.  bench.py calls this for `roofline.traffic`; the numbers printed under ncu are never used as timings.
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def measure(workload, nsteps=30, batch=1, timeout=300, extra_env=None):
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        raise RuntimeError("ncu not found")
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--cache-control", "none",
           "--clock-control", "none", "--print-units", "base", "-k", "regex:k_(stream|resident)_", "--csv",
           sys.executable, os.path.join(ROOT, "tools", "profile_step.py"), workload, str(nsteps), str(batch)]
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT).stdout
    lines = out.splitlines()
    start = next((i for i, l in enumerate(lines) if l.startswith('"ID"')), None)
    if start is None:
        raise RuntimeError("no ncu CSV in the output: " + out[-400:])
    per = {}
    for row in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
        name = row["Kernel Name"].split("(")[0]
        try:
            val = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        per.setdefault(name, {}).setdefault(row["Metric Name"], []).append(val)
    res = {}
    for name, m in per.items():
        rd, wr, tm = (m.get(k, []) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
        skip = 5 if len(rd) > 10 else 0
        mean = lambda v: float(sum(v[skip:]) / max(1, len(v[skip:])))
        res[name] = {"launches": len(rd), "dram_read": mean(rd), "dram_write": mean(wr), "dram_bytes": mean(rd) + mean(wr),
                     "time_us": mean(tm) / 1e3}
    return res


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    wl = args[0] if args else "c3"
    ns = int(args[1]) if len(args) > 1 else 30
    nb = int(args[2]) if len(args) > 2 else 1
    r = measure(wl, ns, nb)
    txt = json.dumps({"workload": wl, "nsteps": ns, "batch": nb, "kernels": r}, indent=1)
    if "--out" in sys.argv:
        open(sys.argv[sys.argv.index("--out") + 1], "w").write(txt + "\n")
    print(txt)
