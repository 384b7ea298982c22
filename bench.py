#!/usr/bin/env python
"""bench.py -- throughput of the elastic FWI hot path on B200 (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c5s] [--impl ours|reference]

Metric (BASELINE.json): elastic cell-updates/s (forward + adjoint) and FWI shot-gradients/s.  A "step" is one pass of the hot
path over one batch of synthetic input:
  c3 (default, BASELINE.json configs[2] -- the largest single-GPU configuration): ONE single-shot FWI gradient (forward with
      boundary saving + reverse-time reconstruction with imaging + adjoint sweep) on the Marmousi-like 1700 x 350 grid
      (padded 416 x 1764, nPml 32), nt = 4001, 1680-channel horizontal DAS fiber at z = 2.  Cell-updates are counted per
      field sweep: live cells x (nt - 1) x 3 sweeps (forward, reconstruction, adjoint) -- the same count for every arm.
  c2 (configs[1]): forward modelling of one shot on the layered 1000 x 400 model, nt = 4001 (kept as the `c2_forward` extra).
  c5s: forward + gradient sample on the 8000 x 2000 grid (nt = 120) -- the HBM-bound size.
The JSON line carries `value` (inputs resident in HBM), `e2e` (pinned host buffers through the public API: model H2D, observed
data H2D, gradients + stf gradient + misfit D2H inside the timed region), `roofline` for the dominant time-stepping kernel
(per-launch CUDA events inside the library; `traffic` measured by an ncu subprocess on the same kernel and workload),
`cpu_baseline` (the CPU oracle port on a bounded sample), and the extras:
  reference_gpu   the reference's own CUDA path (oracle/_ref/libcufd_ref.so, recompiled for sm_100a) on the same gradient
  numba_cpu       the reference's Numba CPU propagator (elasticSolver.forward(), Pool over shots) on C1 / C2, cores stated
  c2_forward      configs[1] forward modelling (the round-1 headline), both scheme flavours
  fwi             configs[3]-like multi-shot gradient, 8 shots per GPU (weak) + packed NCCL all-reduce
  c4_strong       configs[3] proper: 64 FIXED shots sharded over the ranks (strong scaling), all-reduce inside the timed region
  large           per-kernel roofline on the 8000 x 2000 grid (HBM-bound)
  c5_full         configs[4] reduced: 8000 x 2000, nt = 10 000, ONE shot per GPU, 201 MB all-reduce timed warm
  reference_experiment   one L-BFGS evaluation of the reference's own experiment (19 shots, 101 x 201)
`--impl reference` times the CPU restatement of the reference's gradient path (oracle port, all host threads) on a bounded
sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "sep-2023_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import problems  # noqa: E402

B_FWD = 52.0        # algorithmic bytes per forward cell-update: 5 fields R+W + lambda, mu, rho
B_ADJ = 52.0        # adjoint part of a backward step, per live cell
B_REC = 64.0        # reconstruction + imaging part of a backward step, per interior cell
ROUND = "r02"


# ------------------------------------------------------------------------------------------ workloads
def workload(name):
    if name == "c2":
        nz, nx, nt, f0 = 400, 1000, 4001, 15.0
        vp = problems.layered_vp(nz, nx, 2000.0, 4500.0, 8)
        src = [(2, 500)]
        zrec, xrec = np.full(980, 200), np.arange(10, 990)
        desc = "C2 2-D layered elastic 1000x400 (padded 480x1064), nt=4001, 1 shot, horizontal DAS fiber 980 ch, forward modelling"
    elif name == "c3":
        nz, nx, nt, f0 = 350, 1700, 4001, 15.0
        rng = np.random.default_rng(2023)
        vp = problems.layered_vp(nz, nx, 1500.0, 4300.0, 12, rng, nlens=30, lens_amp=0.1, sigma=(5, 40))
        src = [(2, 850)]
        zrec, xrec = np.full(1680, 2), np.arange(10, 1690)
        desc = "C3 Marmousi-like 1700x350 (padded 416x1764), nt=4001, single-shot FWI gradient (fwd + adjoint + boundary saving), fiber z=2 1680 ch"
    elif name in ("c5s", "c5"):
        nz, nx, f0 = 2000, 8000, 10.0
        nt = 120 if name == "c5s" else 10000
        vp = problems.layered_vp(nz, nx, 1500.0, 4500.0, 16)
        src = [(2, 4000)]
        zrec, xrec = np.full(7980, 2), np.arange(10, 7990)
        desc = "C5-size 8000x2000 (padded 2080x8064), nt=%d" % nt
    elif name == "ref":
        # the reference's own experiment grid (notebooks/Main-001-...py:28-72): 101 x 201, 181 adjacent receivers at z = 95
        nz, nx, nt, f0 = 101, 201, 1501, 10.0
        vp = np.full((nz, nx), 4000.0)
        vp[42:58, 42:58] += 80.0
        src = [(1, 100)]
        zrec, xrec = np.full(181, 95), np.arange(10, 191)
        desc = "reference experiment 101x201 (padded 192x265), nt=1501, fiber z=95"
    else:
        raise SystemExit("unknown workload %s" % name)
    nPml = 32
    NZ, NX, nPad = problems.pad_rule(nz, nx, nPml)
    vp_pad = problems.pad_model(vp, nPml, nPad)
    true = problems.lame_from_vp(vp_pad)
    start = problems.lame_from_vp(problems.pad_model(problems.smooth(vp, 10) if name not in ("c5s", "c5") else vp * 0.98, nPml, nPad))
    h, dt = (20.0, 2.0e-3) if name == "ref" else (10.0, 1.0e-3)
    return dict(name=name, desc=desc, nz=NZ, nx=NX, nPml=nPml, nPad=nPad, nSteps=nt, dz=h, dx=h, dt=dt, f0=f0,
                true=true, start=start, src=src, zrec=zrec, xrec=xrec, stf=problems.ricker(f0, nt, dt),
                live=(NZ - nPad) * NX, interior=nz * nx, vp=vp)


def make_shots(w, ShotSpec, nshots=1):
    P = w["nPml"]
    out = []
    for k in range(nshots):
        zs, xs = w["src"][0]
        xs = xs + 7 * k                      # shots of a batch differ by the source position only
        out.append(ShotSpec(zs + P, xs + P, w["zrec"] + P, w["xrec"] + P, w["stf"]))
    return out


def c4_shots(w, ShotSpec, ids):
    """configs[3]: 64 shots x = 20 + 26 k at z = 2, vertical fiber at x = 850, z = 10..339."""
    P0 = w["nPml"]
    zrec, xrec = np.arange(10, 340), np.full(330, 850)
    return [ShotSpec(2 + P0, 20 + 26 * int(k) + P0, zrec + P0, xrec + P0, w["stf"]) for k in ids], len(zrec)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.f = open(self.path, "w")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(dev), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                c = [t.strip() for t in line.split(",")]
                if len(c) < 9:
                    continue
                sm.append(float(c[1])); mx.append(float(c[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def config_of(w, is_grad, kernels):
    """The `config` object of the JSON line -- identical for both arms (the driver compares them key by key)."""
    sweeps = 3 if is_grad else 1
    ws = (28 if is_grad else 13) * w["live"] * 4
    return {"workload": w["desc"], "shots_per_gpu_per_step": 1,
            "cell_update_count": "live padded cells x (nt-1) x %d field sweeps%s" % (sweeps, " (forward, reconstruction, adjoint)" if is_grad else ""),
            "l2": "time loop of %d steps per step; single-shot working set (%.0f MB) %s the 126 MB L2, no flush possible between time steps"
                  % (w["nSteps"] - 1, ws / 1e6, "fits in" if ws < 120e6 else "exceeds"),
            "kernels": "baseline" if kernels == 1 else "default"}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(w, nt_sample, threads, grad):
    """The reference's path restated on the CPU (oracle port: C, OpenMP over rows) on the workload's grid: one forward
    modelling (grad = False) or one full misfit + gradient (grad = True: forward with boundary saving, reverse-time
    reconstruction + imaging, adjoint sweep) of `nt_sample` time steps.  Returns (cell-updates/s, seconds)."""
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from oracle import oracle as O
    par = O.make_par(w["nz"], w["nx"], w["nPml"], w["nPad"], nt_sample, w["dz"], w["dx"], w["dt"], w["f0"])
    zs, xs = w["src"][0]
    stf = np.ascontiguousarray(w["stf"][:nt_sample])
    if not grad:
        t0 = time.perf_counter()
        O.forward(par, *w["true"], stf, zs, xs, w["zrec"], w["xrec"], comps=("ett",))
        dt = time.perf_counter() - t0
        return w["live"] * (nt_sample - 1) / dt, dt
    obs = np.zeros((len(w["xrec"]), nt_sample), np.float32)      # residual = -synthetic: a full-strength adjoint source
    t0 = time.perf_counter()
    O.gradient_shot(par, *w["start"], stf, zs, xs, w["zrec"], w["xrec"], obs, with_adj=True)
    dt = time.perf_counter() - t0
    return 3.0 * w["live"] * (nt_sample - 1) / dt, dt


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args.workload)
    grad = args.workload == "c3"
    cores = os.cpu_count() or 1
    # bounded sample: a twentieth of the shot's time loop per step (c3 gradient: ~5 s on 16 cores; the CPU restatement of the
    # reverse-time half emulates the reference's atomics and is ~5x slower per sweep than its forward half), the whole run within minutes
    nt_sample = 201 if w["live"] < 2e6 else 21
    cpu_sample(w, 21, cores, grad)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_sample(w, nt_sample, cores, grad)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_sample(w, nt_sample, cores, grad)[1]
    sweeps = 3.0 if grad else 1.0
    value = sweeps * w["live"] * (nt_sample - 1) * args.steps / t
    sample = ("%d of %d time steps per step, %s, TorchFWI-flavour CPU restatement (oracle port, C + OpenMP), all %d host threads"
              % (nt_sample - 1, w["nSteps"] - 1, "misfit + gradient (3 sweeps)" if grad else "forward modelling", cores))
    line = {"impl": "reference", "metric": "elastic cell-updates/s (fwd+adj)", "value": value, "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(w, grad, args.kernels), "sample": sample,
            "shot_gradients_per_s": (value / (sweeps * w["live"] * (w["nSteps"] - 1))) if grad else None,
            "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernels", type=int, default=0, help="0 default path, 1 baseline kernels")
    ap.add_argument("--batch", type=int, default=8, help="shots per launch in the batched sections")
    ap.add_argument("--skip-extras", action="store_true")
    ap.add_argument("--no-ncu", action="store_true", help="do not spawn the ncu subprocess that measures DRAM traffic (use the committed capture)")
    ap.add_argument("--extras", default="", help="comma-separated subset of the extras to run (default: all)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from sepfwi import dist as sdist
    from sepfwi.engine import Propagator, ShotSpec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # the contract is ONE JSON line on stdout: whatever libraries print there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, reps):
        """reps calls of fn between barriers; device time by CUDA events on the launching stream, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(e0.elapsed_time(e1))
        barrier()
        return ms / 1e3

    ctx = dict(args=args, torch=torch, dist=dist, sdist=sdist, Propagator=Propagator, ShotSpec=ShotSpec, world=world, rank=rank,
               local=local, dev=dev, timed=timed, barrier=barrier, max_over_ranks=max_over_ranks)

    w = workload(args.workload)
    is_grad = args.workload == "c3"
    nrec = len(w["xrec"])
    P = Propagator(w["nz"], w["nx"], w["nPml"], w["nPad"], w["nSteps"], w["dz"], w["dx"], w["dt"], w["f0"],
                   max_batch=1, max_nrec=nrec, with_adjoint=is_grad, device=local, kernels=args.kernels)
    shots = make_shots(w, ShotSpec, 1)
    model_h = [torch.from_numpy(a).pin_memory() for a in (w["start"] if is_grad else w["true"])]
    model_d = [a.to(dev) for a in model_h]
    sweeps = 3.0 if is_grad else 1.0         # gradient: forward + reconstruction + adjoint sweeps over the live cells
    units = float(w["live"]) * (w["nSteps"] - 1) * sweeps

    if is_grad:
        P.set_model(*[a.to(dev) for a in map(torch.from_numpy, w["true"])])
        obs_d = [o["ett"] for o in P.forward(shots, comps=("ett",), device_out=True)]
        obs_h = [o.cpu().pin_memory() for o in obs_d]
        P.set_model(*model_d)
        step_dev = lambda: P.gradient(shots, obs_d, device=True)

        def step_e2e():
            P.set_model(*[m.numpy() for m in model_h])
            return P.gradient(shots, [o.numpy() for o in obs_h], device=False)
        h2d = 3 * w["nz"] * w["nx"] * 4 + nrec * w["nSteps"] * 4
        d2h = 3 * w["nz"] * w["nx"] * 4 + w["nSteps"] * 4 + 8
    else:
        P.set_model(*model_d)
        comps = ("pr", "vx", "vz", "ett")
        out_h = [{c: torch.empty((nrec, w["nSteps"]), dtype=torch.float32).pin_memory().numpy() for c in comps}]
        step_dev = lambda: P.forward(shots, comps=comps, device_out=True)

        def step_e2e():
            P.set_model(*[m.numpy() for m in model_h])
            return P.forward(shots, comps=comps, out=out_h)
        h2d = 3 * w["nz"] * w["nx"] * 4 + w["nSteps"] * 4
        d2h = 4 * nrec * w["nSteps"] * 4

    # ---- value: inputs resident in HBM
    for _ in range(args.warmup):
        step_dev()
    l0 = P.launches
    clocks = ClockSampler(local)
    t_dev = timed(step_dev, args.steps)
    launches = P.launches - l0
    clk = clocks.stop()
    fwd_ms, bwd_ms = P.last_timing()
    value = world * units * args.steps / t_dev

    # ---- e2e: host (pinned) buffers through the public API
    step_e2e()
    t_e2e = timed(step_e2e, args.steps)
    e2e = world * units * args.steps / t_e2e

    # ---- roofline of the time-stepping kernels (per-launch CUDA events inside the library, on the launching stream)
    P.set_model(*model_d)
    P.set_profile(400)
    step_dev()
    prof = P.profile()
    P.set_profile(0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    P.close()
    roof = headline_roofline(w, prof, peak, peak_src, rank, world, args)

    line = {"metric": "elastic cell-updates/s (fwd+adj)", "value": value, "unit": "cell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(w, is_grad, args.kernels),
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": e2e, "unit": "cell-updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * t_e2e / args.steps},
            "roofline": roof, "loop_ms": {"forward": fwd_ms, "backward": bwd_ms}}
    if is_grad:
        line["shot_gradients_per_s"] = world * args.steps / t_dev
        line["gradient_cell_steps_per_s"] = world * float(w["live"]) * (w["nSteps"] - 1) * args.steps / t_dev
        line["e2e"]["shot_gradients_per_s"] = world * args.steps / t_e2e

    # ---- extras (bounded; an extra must never hide the headline)
    extras = [("reference_gpu", extra_reference_gpu), ("c2_forward", extra_c2_forward), ("fwi", extra_fwi), ("c4_strong", extra_c4_strong),
              ("large", extra_large), ("c5_full", extra_c5_full), ("reference_experiment", extra_reference_experiment),
              ("numba_cpu", extra_numba_cpu)]
    only = [e for e in args.extras.split(",") if e]
    if not args.skip_extras:
        for name, fn in extras:
            if only and name not in only:
                continue
            try:
                line[name] = fn(ctx, peak)
            except Exception as e:
                line[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            torch.cuda.empty_cache()
    if isinstance(line.get("reference_gpu"), dict) and line["reference_gpu"].get("c3_gradient_s"):
        line["vs_reference_gpu"] = {"c3_gradient_speedup_e2e": line["reference_gpu"]["c3_gradient_s"] / (t_e2e / args.steps),
                                    "note": "reference CUDA path (host buffers in, host buffers out, like its pybind entry) vs our e2e step, same box, same gradient"}

    # ---- CPU baseline on the box's host cores (rank 0, N = 1 only): bounded sample of the SAME workload
    if rank == 0 and world == 1:
        cores = os.cpu_count() or 1
        nt_sample = 401 if w["live"] < 2e6 else 21
        cpu_sample(w, 21, cores, is_grad)
        tt, n = 0.0, 0
        while tt < 10.0 and n < 6:
            tt += cpu_sample(w, nt_sample, cores, is_grad)[1]
            n += 1
        v = sweeps * w["live"] * (nt_sample - 1) * n / tt
        line["cpu_baseline"] = {"value": v, "unit": "cell-updates/s", "cores": cores, "kind": "port",
                                "sample": "%d x %d of %d time steps of the same workload (%s), CPU oracle port (C, OpenMP over rows), %.1f s"
                                          % (n, nt_sample - 1, w["nSteps"] - 1, "misfit + gradient, 3 sweeps" if is_grad else "forward", tt)}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


def git_head():
    """Commit of the tree being measured: git where there is one, else what build() recorded next to the library."""
    try:
        h = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True, timeout=10).stdout.strip()
        if h:
            return h
    except Exception:
        pass
    try:
        return json.load(open(os.path.join(ROOT, "sep-2023_b200", "build_info.json"))).get("commit")
    except Exception:
        return None


def measured_traffic(name, nsteps, batch, rank, world, skip):
    """DRAM bytes per launch of the time-stepping kernels on workload `name`: measured now by an ncu subprocess (rank 0 at N = 1),
    else the committed capture of this round (profiles/r02_traffic.json, keyed by workload, with the commit it was taken at)."""
    if rank == 0 and world == 1 and not skip:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import measure_traffic
            r = measure_traffic.measure(name, nsteps, batch, timeout=240)
            if r:
                return {k.replace("k_", ""): v for k, v in r.items()}, "measured in this run: ncu --cache-control none subprocess, tools/measure_traffic.py %s %d %d" % (name, nsteps, batch)
        except Exception as e:
            err = "%s: %s" % (type(e).__name__, str(e)[:120])
        else:
            err = "no kernels captured"
    else:
        err = "not rank 0 / N > 1 / skipped"
    try:
        js = json.load(open(os.path.join(ROOT, "profiles", ROUND + "_traffic.json")))
        ent = js[name]
        return ent["kernels"], "profiles/%s_traffic.json (ncu capture at commit %s; live measurement unavailable: %s)" % (ROUND, ent.get("commit"), err)
    except Exception:
        return {}, "unavailable (%s)" % err


def headline_roofline(w, prof, peak, peak_src, rank, world, args):
    def avg(name):
        ms, n = prof.get(name, (0.0, 0))
        return ms / n if n else 0.0
    is_grad = w["name"] == "c3"
    resident = "resident_fwd" in prof
    steps_per = {k: ((w["nSteps"] - 1) if k == "resident_fwd" else 1) for k in prof}
    alg = {"resident_fwd": B_FWD * w["live"] * (w["nSteps"] - 1), "stream_fwd": B_FWD * w["live"],
           "stream_adj": B_ADJ * w["live"], "stream_recon": B_REC * w["interior"],
           "stream_bwd": B_ADJ * w["live"] + B_REC * w["interior"]}      # reconstruction + adjoint sweep in one launch
    kern = {k: {"avg_launch_us": 1e3 * avg(k), "launches_timed": prof[k][1], "algorithmic_bytes_per_launch": alg[k],
                "achieved_GBs": alg[k] / (avg(k) * 1e-3) / 1e9, "frac": alg[k] / (avg(k) * 1e-3) / 1e9 / peak}
            for k in prof if k in alg and avg(k) > 0}
    if not kern:
        return None
    # share of one time step: the dominant kernel is the one the step spends most of its time in
    step_us = {k: v["avg_launch_us"] / steps_per[k] for k, v in kern.items()}
    dom = max(step_us, key=step_us.get)
    traffic, tsrc = measured_traffic(w["name"], 40, 1, rank, world, args.no_ncu)
    for k in kern:
        if k in traffic:
            kern[k]["traffic"] = traffic[k]["dram_bytes"]
            kern[k]["traffic_over_algorithmic"] = traffic[k]["dram_bytes"] / alg[k]
    roof = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_GBs"], "peak": peak, "unit": "GB/s",
            "frac": kern[dom]["frac"], "traffic": kern[dom].get("traffic"), "traffic_source": tsrc, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg[dom], "avg_launch_us": kern[dom]["avg_launch_us"],
            "share_of_time_step": step_us[dom] / sum(step_us.values()), "per_kernel": kern, "commit": git_head(),
            "note": ("dominant kernel of the step by time; single-shot working set fits the 126 MB L2, so DRAM traffic may be far below the "
                     "algorithmic bytes -- the kernel is then latency / L2-bound and frac is an equivalent bandwidth (see traffic); "
                     "the HBM-bound figures are in `large` and `fwi`")}
    if is_grad:
        t_step = sum(step_us.values()) * 1e-6
        a_all = (B_FWD * w["live"] + B_ADJ * w["live"] + B_REC * w["interior"]) / t_step / 1e9      # each byte count once: stream_bwd = recon + adj
        roof["whole_gradient_step"] = {"us": t_step * 1e6, "algorithmic_GBs": a_all, "frac": a_all / peak}
    if resident:
        roof["time_steps_per_launch"] = w["nSteps"] - 1
    return roof


def _prop(ctx, w, **kw):
    return ctx["Propagator"](w["nz"], w["nx"], w["nPml"], w["nPad"], w["nSteps"], w["dz"], w["dx"], w["dt"], w["f0"],
                             device=ctx["local"], kernels=ctx["args"].kernels, **kw)


def extra_reference_gpu(ctx, peak):
    """The reference's OWN CUDA shot driver (libCUFD.cu `cufd`, recompiled for sm_100a by oracle/Makefile, unmodified) on this
    box: C3 single-shot gradient (calc_id 1) and one evaluation of the reference experiment (19 shots).  Wall time of the call --
    host pointers in and out, like its pybind entry (Torch_Fwi.cpp:38-104); it allocates, parses and reads its files per call."""
    if ctx["rank"] != 0 or ctx["world"] != 1:
        return None
    from oracle import ref_cufd
    from sepfwi import fwi_utils as ft
    if not ref_cufd.available():
        return {"unavailable": "oracle/_ref/libcufd_ref.so not present"}
    out = {"what": "reference CUDA path (Ops/FWI/Src, nvcc sm_100a), wall time per cufd call, same GPU"}
    tmp = tempfile.mkdtemp(prefix="refgpu_")

    def run(w, z_src, x_src, tag, reps):
        para, survey, data = os.path.join(tmp, tag + "_para.json"), os.path.join(tmp, tag + "_survey.json"), os.path.join(tmp, tag + "_d")
        ft.paraGen(w["nz"], w["nx"], w["dz"], w["dx"], w["nSteps"], w["dt"], w["f0"], w["nPml"], w["nPad"], para, survey, data)
        ft.surveyGen(z_src, x_src, w["zrec"], w["xrec"], survey)
        stf = np.tile(w["stf"][None, :], (len(x_src), 1)).astype(np.float32)
        ids = np.arange(len(x_src), dtype=np.int32)
        ref_cufd.cufd(2, *w["true"], stf, ids, para)
        ref_cufd.cufd(1, *w["start"], stf, ids, para)          # warm-up (context, module load)
        t0 = time.perf_counter()
        for _ in range(reps):
            r = ref_cufd.cufd(1, *w["start"], stf, ids, para)
        return (time.perf_counter() - t0) / reps, r[0]

    w = workload("c3")
    t, J = run(w, [w["src"][0][0]], [w["src"][0][1]], "c3", 2)
    out["c3_gradient_s"] = t
    out["c3_shot_gradients_per_s"] = 1.0 / t
    out["c3_cell_updates_per_s"] = 3.0 * w["live"] * (w["nSteps"] - 1) / t
    out["c3_misfit"] = J
    w = workload("ref")
    xs = np.arange(10, 191, 10)
    t, J = run(w, np.full(len(xs), 1), xs, "ref", 2)
    out["reference_experiment_s_per_evaluation"] = t
    return out


def extra_numba_cpu(ctx, peak):
    """The reference's Numba CPU propagator (DAS_Waveform_Modeling/src/elasticSolver.py:156-182, `forward()` = multiprocessing
    Pool over shots, fp64, sponge flavour) on this box's host cores: C1 (201 x 201, nt 1001) and a bounded C2 sample."""
    if ctx["rank"] != 0 or ctx["world"] != 1:
        return None
    from oracle import numba_ref
    if not numba_ref.available():
        return {"unavailable": "oracle/_ref/numba_ref/elasticSolver.py or numba not present"}
    cores = os.cpu_count() or 1
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    out = {"what": "reference Numba CPU propagator, elasticSolver.forward() with one shot per host core (its own Pool over shots)",
           "cores": cores, "cpu": cpu}
    v, wall = numba_ref.time_forward(201, 201, 40, 10.0, 10.0, 1e-3, 1001, 10.0, np.full((201, 201), 4000.0), cores)
    out["c1"] = {"cell_updates_per_s": v, "wall_s": wall, "shots": cores, "grid": "201x201 + 2x40 sponge, nt=1001 (configs[0])"}
    w = workload("c2")
    nt = 201
    v, wall = numba_ref.time_forward(1000, 400, 32, 10.0, 10.0, 1e-3, nt, 15.0, np.ascontiguousarray(w["vp"].T), cores)
    out["c2"] = {"cell_updates_per_s": v, "wall_s": wall, "shots": cores, "grid": "1000x400 + 2x32 sponge, %d of 4001 steps (configs[1] sample)" % nt}
    return out


def extra_c2_forward(ctx, peak):
    """configs[1] forward modelling, one shot, both scheme flavours: the CPML flavour through the
    shared-memory-resident forward loop (round-1 headline), the sponge flavour (elasticSolver's scheme) through its kernels."""
    torch, ShotSpec, timed, world = ctx["torch"], ctx["ShotSpec"], ctx["timed"], ctx["world"]
    from sepfwi import _lib
    w = workload("c2")
    nrec = len(w["xrec"])
    out = {"workload": w["desc"]}
    with _prop(ctx, w, max_batch=1, max_nrec=nrec) as P:
        P.set_model(*[torch.from_numpy(a).to(ctx["dev"]) for a in w["true"]])
        shots = make_shots(w, ShotSpec, 1)
        fn = lambda: P.forward(shots, comps=("pr", "vx", "vz", "ett"), device_out=True)
        fn(); fn()
        t = timed(fn, 5) / 5
        out["cpml"] = {"cell_updates_per_s": world * w["live"] * (w["nSteps"] - 1) / t, "ms_per_shot": 1e3 * t,
                       "us_per_time_step": 1e6 * t / (w["nSteps"] - 1), "resident_launches": P.resident_launches,
                       "equivalent_GBs": B_FWD * w["live"] * (w["nSteps"] - 1) / t / 1e9, "frac": B_FWD * w["live"] * (w["nSteps"] - 1) / t / 1e9 / peak}
    # sponge flavour: lambda, mu in Pa on the (nz + 2 ndamp, nx + 2 ndamp) grid, no alignment padding
    nd = 32
    vp = np.pad(w["vp"], nd, mode="edge")
    lam, mu, rho = problems.lame_from_vp(vp)
    with ctx["Propagator"](vp.shape[0], vp.shape[1], nd, 0, w["nSteps"], w["dz"], w["dx"], w["dt"], w["f0"], flavour=_lib.FLAVOUR_SPONGE,
                           max_batch=1, max_nrec=nrec, device=ctx["local"]) as P:
        P.set_model(*[torch.from_numpy(np.ascontiguousarray(a * s, np.float32)).to(ctx["dev"]) for a, s in ((lam, 1e6), (mu, 1e6), (rho, 1.0))])
        shots = [ShotSpec(2 + nd, 500 + nd, w["zrec"] + nd, w["xrec"] + nd, problems.ricker(w["f0"], w["nSteps"], w["dt"], amp=1.0))]
        fn = lambda: P.forward(shots, comps=("pr", "vx", "vz", "ett"), device_out=True)
        fn()
        t = timed(fn, 2) / 2
        cells = float(vp.shape[0] * vp.shape[1])
        out["sponge"] = {"cell_updates_per_s": world * cells * w["nSteps"] / t, "ms_per_shot": 1e3 * t, "us_per_time_step": 1e6 * t / w["nSteps"]}
    return out


def extra_fwi(ctx, peak):
    """Multi-shot FWI gradient (configs[3]-like: Marmousi-like 1700x350, vertical DAS fiber, nt=4001), `batch` shots per GPU (weak
    scaling), per-GPU gradients written straight into the persistent packed buffer and summed by ONE NCCL all-reduce."""
    args, torch, ShotSpec, timed, world, rank, sdist, dev = (ctx[k] for k in ("args", "torch", "ShotSpec", "timed", "world", "rank", "sdist", "dev"))
    w = workload("c3")
    B = args.batch
    ids = [(rank * B + k) % 64 for k in range(B)]
    shots, nrec = c4_shots(w, ShotSpec, ids)
    with _prop(ctx, w, fiber=1, max_batch=B, max_nrec=nrec, with_adjoint=True) as P:
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["true"]])
        obs = [o["ett"] for o in P.forward(shots, comps=("ett",), device_out=True)]
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["start"]])
        pk = sdist.PackedGradients((w["nz"], w["nx"]), (64, w["nSteps"]), dev)
        ar = {}

        def fn():
            r = P.gradient(shots, obs, device=True, grad_out=pk.views())
            pk.set_local(r["misfit64"], dict(zip(ids, r["gstf"])))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = pk.allreduce()
            e1.record()
            ar["ev"] = (e0, e1)
            return out

        fn()
        t = timed(fn, 1)
        torch.cuda.synchronize()
        f_ms, b_ms = P.last_timing()
        nst = w["nSteps"] - 1
        a_f = B * B_FWD * w["live"] * nst / (f_ms * 1e-3) / 1e9
        a_b = B * (B_ADJ * w["live"] + B_REC * w["interior"]) * nst / (b_ms * 1e-3) / 1e9
        return {"workload": "C4-like multi-shot FWI gradient, Marmousi-like 1700x350 (padded 416x1764), nt=4001, vertical DAS fiber 330 ch",
                "shots_per_gpu": B, "n_gpus": world, "scaling": "weak", "shot_gradients_per_s": world * B / t,
                "cell_updates_per_s": world * B * 3.0 * w["live"] * nst / t,
                "s_per_evaluation": t, "forward_loop_ms": f_ms, "backward_loop_ms": b_ms,
                "forward_loop_frac": a_f / peak, "backward_loop_frac": a_b / peak,
                "gradient_frac": B * (B_FWD * w["live"] + B_ADJ * w["live"] + B_REC * w["interior"]) * nst / ((f_ms + b_ms) * 1e-3) / 1e9 / peak,
                "working_set_MB": B * 28 * w["live"] * 4 / 1e6,
                "allreduce_ms": ar["ev"][0].elapsed_time(ar["ev"][1]) if world > 1 else 0.0, "allreduce_bytes": pk.nbytes}


def extra_c4_strong(ctx, peak):
    """configs[3] proper: the 64-shot survey (x = 20 + 26 k, vertical fiber at x = 850) as ONE misfit + gradient evaluation,
    shots contiguous-sharded over the ranks exactly like the reference (Torch_Fwi.cpp:59-80) -- 64 / 32 / 16 / 8 per GPU -- with
    the packed all-reduce inside the timed region: STRONG scaling."""
    args, torch, ShotSpec, timed, world, rank, sdist, dev = (ctx[k] for k in ("args", "torch", "ShotSpec", "timed", "world", "rank", "sdist", "dev"))
    w = workload("c3")
    ids = sdist.shard(list(range(64)), world, rank)
    shots, nrec = c4_shots(w, ShotSpec, ids)
    # shots per launch: as many as fit (the rank's whole share when memory allows -- 64 slots of this grid are ~125 GB): measured 5.7 /
    # 6.6 / 7.2 / 7.8 shot-gradients/s at 8 / 16 / 32 / 64 shots per launch
    from sepfwi import fwi_ops
    B = fwi_ops.auto_batch(w["nz"], w["nx"], w["nPad"], w["nSteps"], nrec, w["nPml"], True, len(ids), ctx["local"])
    with _prop(ctx, w, fiber=1, max_batch=B, max_nrec=nrec, with_adjoint=True) as P:
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["true"]])
        obs = [o["ett"] for o in P.forward(shots, comps=("ett",), device_out=True)]
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["start"]])
        pk = sdist.PackedGradients((w["nz"], w["nx"]), (64, w["nSteps"]), dev)
        res = {}

        def fn():
            r = P.gradient(shots, obs, device=True, grad_out=pk.views())
            pk.set_local(r["misfit64"], dict(zip(ids, r["gstf"])))
            res["J"] = pk.allreduce()[0]

        fn()                          # warm-up evaluation
        t = timed(fn, 1)
        return {"workload": "configs[3]: 64-shot elastic FWI gradient, vertical DAS fiber, 1700x350, nt=4001, one evaluation incl. all-reduce",
                "scaling": "strong", "shots_total": 64, "shots_this_rank": len(ids), "shots_per_launch": B, "n_gpus": world,
                "s_per_evaluation": t, "shot_gradients_per_s": 64.0 / t, "cell_updates_per_s": 64 * 3.0 * w["live"] * (w["nSteps"] - 1) / t,
                "misfit": res["J"], "allreduce_bytes": pk.nbytes}


def extra_reference_experiment(ctx, peak):
    """One misfit + gradient evaluation of the reference's own experiment (notebooks/Main-001-...py: 101 x 201 grid padded to
    192 x 265, nt = 1501, 19 shots, 181 adjacent receivers) -- what one L-BFGS function evaluation costs per GPU."""
    torch, ShotSpec, timed, dev = ctx["torch"], ctx["ShotSpec"], ctx["timed"], ctx["dev"]
    w = workload("ref")
    P0 = w["nPml"]
    xs = np.arange(10, 191, 10)
    shots = [ShotSpec(1 + P0, int(x) + P0, w["zrec"] + P0, w["xrec"] + P0, w["stf"]) for x in xs]
    with _prop(ctx, w, max_batch=len(shots), max_nrec=len(w["xrec"]), with_adjoint=True) as P:
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["true"]])
        obs = [o["ett"] for o in P.forward(shots, comps=("ett",), device_out=True)]
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["start"]])
        fn = lambda: P.gradient(shots, obs, device=True)
        fn()
        t0 = time.perf_counter()
        t = timed(fn, 3) / 3
        wall = (time.perf_counter() - t0) / 3
        f_ms, b_ms = P.last_timing()
        return {"workload": w["desc"] + ", 19 shots in one batch", "s_per_evaluation": t, "wall_s_per_evaluation": wall,
                "shot_gradients_per_s": len(shots) / t,
                "forward_loop_ms": f_ms, "backward_loop_ms": b_ms, "resident_forward_launches": P.resident_launches}


def _kernel_roofline(w, prof, peak, traffic):
    us = {k: 1e3 * ms / n for k, (ms, n) in prof.items()}
    alg = {"stream_fwd": B_FWD * w["live"], "stream_adj": B_ADJ * w["live"], "stream_recon": B_REC * w["interior"],
           "stream_bwd": B_ADJ * w["live"] + B_REC * w["interior"]}
    per = {}
    for k in us:
        if k in alg:
            per[k] = {"avg_launch_us": us[k], "algorithmic_bytes_per_launch": alg[k], "achieved_GBs": alg[k] / (us[k] * 1e-6) / 1e9,
                      "frac": alg[k] / (us[k] * 1e-6) / 1e9 / peak}
            if k in traffic:
                per[k]["traffic"] = traffic[k]["dram_bytes"]
                per[k]["traffic_over_algorithmic"] = traffic[k]["dram_bytes"] / alg[k]
                per[k]["dram_GBs"] = traffic[k]["dram_bytes"] / (us[k] * 1e-6) / 1e9
    return us, per


def extra_large(ctx, peak):
    """HBM-bound size (8000 x 2000 grid, 67 MB per field, working set >> L2): per-kernel roofline with measured DRAM traffic."""
    torch, ShotSpec, dev = ctx["torch"], ctx["ShotSpec"], ctx["dev"]
    w = workload("c5s")
    nrec = len(w["xrec"])
    with _prop(ctx, w, max_batch=1, max_nrec=nrec, with_adjoint=True) as P:
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["true"]])
        shots = make_shots(w, ShotSpec, 1)
        obs = [o["ett"] for o in P.forward(shots, comps=("ett",), device_out=True)]
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["start"]])
        P.gradient(shots, obs, device=True)
        P.set_profile(w["nSteps"])
        P.gradient(shots, obs, device=True)
        prof = P.profile()
    traffic, tsrc = measured_traffic("c5s", 24, 1, ctx["rank"], ctx["world"], ctx["args"].no_ncu)
    us, per = _kernel_roofline(w, prof, peak, traffic)
    fk = [k for k in ("stream_fwd",) if k in us]
    bk = [k for k in ("stream_recon", "stream_adj", "stream_bwd") if k in us]
    tf, tb = sum(us[k] for k in fk) * 1e-6, sum(us[k] for k in bk) * 1e-6
    af = B_FWD * w["live"] / tf / 1e9
    ab = (B_ADJ * w["live"] + B_REC * w["interior"]) / tb / 1e9
    ag = (B_FWD * w["live"] + B_ADJ * w["live"] + B_REC * w["interior"]) / (tf + tb) / 1e9
    return {"workload": w["desc"], "per_kernel_us": us, "per_kernel_roofline": per, "traffic_source": tsrc,
            "forward_step": {"kernels": "+".join(fk), "achieved_GBs": af, "frac": af / peak, "cell_updates_per_s": w["live"] / tf},
            "backward_step": {"kernels": "+".join(bk), "us": tb * 1e6, "achieved_GBs": ab, "frac": ab / peak, "cell_steps_per_s": w["live"] / tb},
            "gradient_step": {"us": (tf + tb) * 1e6, "achieved_GBs": ag, "frac": ag / peak, "gradient_cell_steps_per_s": w["live"] / (tf + tb)},
            "peak_GBs": peak}


def extra_c5_full(ctx, peak):
    """configs[4] reduced but honest: the 8000 x 2000 model (padded 2080 x 8064), nt = 10 000, ONE shot per GPU (of the 256; 32 per
    GPU would take 32x as long), 7980-channel fiber, 20 GB boundary ring per shot; misfit + gradient, then the packed all-reduce
    (201 MB) timed WARM (second of two)."""
    torch, ShotSpec, timed, world, rank, sdist, dev = (ctx[k] for k in ("torch", "ShotSpec", "timed", "world", "rank", "sdist", "dev"))
    free = torch.cuda.mem_get_info(dev)[0]
    if free < 40e9:
        return {"skipped": "needs ~30 GB of free device memory, %.0f GB free" % (free / 1e9)}
    w = workload("c5")
    nrec = len(w["xrec"])
    P0 = w["nPml"]
    xs = 200 + (rank * 997) % 7600
    shots = [ShotSpec(2 + P0, xs + P0, w["zrec"] + P0, w["xrec"] + P0, w["stf"])]
    with _prop(ctx, w, max_batch=1, max_nrec=nrec, with_adjoint=True) as P:
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["true"]])
        obs = [o["ett"] for o in P.forward(shots, comps=("ett",), device_out=True)]
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["start"]])
        pk = sdist.PackedGradients((w["nz"], w["nx"]), (256, w["nSteps"]), dev)
        ar = {}

        def fn():
            r = P.gradient(shots, obs, device=True, grad_out=pk.views())
            pk.set_local(r["misfit64"], {rank: r["gstf"][0]})
            pk.allreduce()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ar["J"] = pk.allreduce()[0]          # warm: the second all-reduce of the same buffer (values are not used further)
            e1.record()
            ar["ev"] = (e0, e1)

        t = timed(fn, 1)
        torch.cuda.synchronize()
        f_ms, b_ms = P.last_timing()
        nst = w["nSteps"] - 1
        ag = (B_FWD * w["live"] + B_ADJ * w["live"] + B_REC * w["interior"]) * nst / ((f_ms + b_ms) * 1e-3) / 1e9
        return {"workload": w["desc"] + ", 1 shot per GPU (256-shot survey: x32 per GPU), misfit + gradient + 201 MB all-reduce",
                "n_gpus": world, "s_per_shot_gradient": (f_ms + b_ms) * 1e-3, "s_per_evaluation_1_shot_per_gpu": t,
                "forward_loop_ms": f_ms, "backward_loop_ms": b_ms, "gradient_frac": ag / peak,
                "gradient_cell_steps_per_s_per_gpu": w["live"] * nst / ((f_ms + b_ms) * 1e-3),
                "projected_s_per_256_shot_evaluation": 32.0 * (f_ms + b_ms) * 1e-3 * 8.0 / max(world, 1),
                "allreduce_warm_ms": ar["ev"][0].elapsed_time(ar["ev"][1]) if world > 1 else 0.0, "allreduce_bytes": pk.nbytes,
                "ring_GB": 5 * 10 * ((w["nz"] - w["nPad"] - 2 * P0 + 4) + (w["nx"] - 2 * P0 + 4)) * w["nSteps"] * 4 / 1e9}


if __name__ == "__main__":
    main()
