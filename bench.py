#!/usr/bin/env python
"""bench.py -- throughput of the elastic propagator hot path on B200 (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c5s] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input:
  c2 (default, BASELINE.json configs[1]): forward modelling of one shot on the 2-D layered elastic
      model 1000 x 400 (padded 480 x 1064), nt = 4001, horizontal DAS fiber of 980 channels,
      CPML flavour; unit = cell-updates (padded live cells x time steps).
  c3: one single-shot FWI gradient (forward + adjoint + boundary-saving reconstruction) on the
      Marmousi-like 1700 x 350 grid, nt = 4001.
  c5s: forward + gradient sample on the 8000 x 2000 grid (nt = 120) -- the HBM-bound size.
The JSON line carries `value` (inputs resident in HBM), `e2e` (host buffers through the public API,
H2D/D2H inside the timed region), `roofline` for the time-stepping kernels, `cpu_baseline` (the CPU
oracle port on the host cores), and -- for the record -- an FWI-gradient section with the NCCL
all-reduce (`fwi`) and an HBM-bound large-grid section (`large`).
`--impl reference` times the reference's CPU path (oracle port, all host threads) on the same workload.
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
        desc = "C3 Marmousi-like 1700x350 (padded 416x1764), nt=4001, single-shot FWI gradient, fiber z=2"
    elif name == "c5s":
        nz, nx, nt, f0 = 2000, 8000, 120, 10.0
        vp = problems.layered_vp(nz, nx, 1500.0, 4500.0, 16)
        src = [(2, 4000)]
        zrec, xrec = np.full(7980, 2), np.arange(10, 7990)
        desc = "C5-size 8000x2000 (padded 2080x8064) sample, nt=120"
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
    start = problems.lame_from_vp(problems.pad_model(problems.smooth(vp, 10) if name != "c5s" else vp * 0.98, nPml, nPad))
    h, dt = (20.0, 2.0e-3) if name == "ref" else (10.0, 1.0e-3)
    return dict(name=name, desc=desc, nz=NZ, nx=NX, nPml=nPml, nPad=nPad, nSteps=nt, dz=h, dx=h, dt=dt, f0=f0,
                true=true, start=start, src=src, zrec=zrec, xrec=xrec, stf=problems.ricker(f0, nt, dt),
                live=(NZ - nPad) * NX, interior=nz * nx)


def make_shots(w, ShotSpec, nshots=1):
    P = w["nPml"]
    out = []
    for k in range(nshots):
        zs, xs = w["src"][0]
        xs = xs + 7 * k                      # shots of a batch differ by the source position only
        out.append(ShotSpec(zs + P, xs + P, w["zrec"] + P, w["xrec"] + P, w["stf"]))
    return out


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


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_forward_sample(w, nt_sample, threads):
    """The reference's CPU path restated (oracle port, OpenMP over rows) on the workload's grid."""
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from oracle import oracle as O
    par = O.make_par(w["nz"], w["nx"], w["nPml"], w["nPad"], nt_sample, w["dz"], w["dx"], w["dt"], w["f0"])
    zs, xs = w["src"][0]
    t0 = time.perf_counter()
    O.forward(par, *w["true"], w["stf"][:nt_sample], zs, xs, w["zrec"], w["xrec"], comps=("ett",))
    dt = time.perf_counter() - t0
    return w["live"] * (nt_sample - 1) / dt, dt


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args.workload)
    cores = os.cpu_count() or 1
    nt_sample = w["nSteps"] if w["live"] < 2e6 else 41        # c2 / c3: the whole time loop of the shot (4 - 8 s on 16 cores)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_forward_sample(w, 21, cores)
    vals, t = [], 0.0
    for _ in range(args.steps):
        v, dt = cpu_forward_sample(w, nt_sample, cores)
        vals.append(v); t += dt
    value = w["live"] * (nt_sample - 1) * args.steps / t
    sample = "%d of %d time steps per step, forward modelling, TorchFWI-flavour CPU restatement" % (nt_sample - 1, w["nSteps"] - 1)
    line = {"impl": "reference", "metric": "elastic cell-updates/s", "value": value, "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "sample": sample},
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
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernels", type=int, default=0, help="0 default path, 1 baseline kernels")
    ap.add_argument("--batch", type=int, default=8, help="shots per launch in the batched sections")
    ap.add_argument("--skip-extras", action="store_true")
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

    w = workload(args.workload)
    is_grad = args.workload == "c3"
    nrec = len(w["xrec"])
    P = Propagator(w["nz"], w["nx"], w["nPml"], w["nPad"], w["nSteps"], w["dz"], w["dx"], w["dt"], w["f0"],
                   max_batch=1, max_nrec=nrec, with_adjoint=is_grad, device=local, kernels=args.kernels)
    shots = make_shots(w, ShotSpec, 1)
    model_h = [torch.from_numpy(a).pin_memory() for a in (w["start"] if is_grad else w["true"])]
    model_d = [a.to(dev) for a in model_h]
    units = float(w["live"]) * (w["nSteps"] - 1) * (3.0 if is_grad else 1.0)   # gradient: forward + reconstruction + adjoint sweeps

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
        d2h = 3 * w["nz"] * w["nx"] * 4 + w["nSteps"] * 4 + 4
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

    # ---- roofline of the time-stepping kernels (per-launch CUDA events inside the library)
    P.set_model(*model_d)
    P.set_profile(200)
    step_dev()
    prof = P.profile()
    P.set_profile(0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"

    def avg(name):
        ms, n = prof.get(name, (0.0, 0))
        return ms / n if n else 0.0

    resident = "resident_fwd" in prof      # the whole time loop is ONE cooperative launch (kernels_resident.cuh)
    fwd_kernels = ["resident_fwd"] if resident else [k for k in ("stream_fwd", "fused_fwd", "stress_fwd", "velocity_fwd") if k in prof]
    steps_per_launch = (w["nSteps"] - 1) if resident else 1
    t_fwd_step = sum(avg(k) for k in fwd_kernels) * 1e-3
    traffic = None      # DRAM bytes per launch of the same kernel from the committed ncu --set full capture (profiles/)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(w["name"], {})
        if fwd_kernels and all(k in tr for k in fwd_kernels):
            traffic = float(sum(tr[k] for k in fwd_kernels))
    except Exception:
        pass
    roof = None
    if t_fwd_step > 0:
        ach = B_FWD * w["live"] * steps_per_launch / t_fwd_step / 1e9
        note = ("one launch = the whole time loop (%d steps) of the shot, tiles resident in shared memory: DRAM traffic is far below the "
                "algorithmic bytes, so frac may exceed 1 (see traffic)" % steps_per_launch) if resident else \
               ("one forward time step = %s; working set %.0f MB (%s the 126 MB L2)"
                % ("+".join(fwd_kernels), 13 * w["live"] * 4 / 1e6, "fits in" if 13 * w["live"] * 4 < 100e6 else "exceeds"))
        roof = {"bound": "hbm", "kernel": "+".join(fwd_kernels), "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": B_FWD * w["live"] * steps_per_launch, "avg_launch_us": t_fwd_step * 1e6,
                "time_steps_per_launch": steps_per_launch, "us_per_time_step": t_fwd_step * 1e6 / steps_per_launch,
                "note": note,
                "per_kernel_us": {k: 1e3 * avg(k) for k in prof}}
        if is_grad:
            bk = [k for k in ("stream_recon", "stream_adj", "fused_recon", "fused_adj", "velocity_bwd", "stress_bwd", "velocity_adj", "stress_adj", "inject") if k in prof]
            t_b = sum(avg(k) for k in bk) * 1e-3
            if t_b > 0:
                achb = (B_ADJ * w["live"] + B_REC * w["interior"]) / t_b / 1e9
                roof["backward"] = {"kernel": "+".join(bk), "achieved": achb, "frac": achb / peak, "avg_step_us": t_b * 1e6}

    line = {"metric": "elastic cell-updates/s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "shots_per_gpu_per_step": 1, "l2": "time loop of %d steps per step; working set %s L2, no flush possible between time steps"
                       % (w["nSteps"] - 1, "fits in" if 13 * w["live"] * 4 < 100e6 else "exceeds"),
                       "kernels": "baseline" if args.kernels == 1 else "default"},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": e2e, "unit": "cell-updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * t_e2e / args.steps},
            "roofline": roof, "loop_ms": {"forward": fwd_ms, "backward": bwd_ms}}

    # ---- extras (bounded): batched shots, FWI gradient with all-reduce, HBM-bound large grid
    if not args.skip_extras:
        try:
            line["batched"] = extra_batched(args, w, P.__class__, ShotSpec, torch, dev, local, world, timed)
        except Exception as e:   # an extra must never hide the headline
            line["batched"] = {"error": str(e)[:200]}
        P.close()
        try:
            line["fwi"] = extra_fwi(args, Propagator, ShotSpec, torch, dev, local, world, rank, timed, sdist)
        except Exception as e:
            line["fwi"] = {"error": str(e)[:200]}
        try:
            line["large"] = extra_large(args, Propagator, ShotSpec, torch, dev, local, peak)
        except Exception as e:
            line["large"] = {"error": str(e)[:200]}
        try:
            line["reference_experiment"] = extra_reference_experiment(Propagator, ShotSpec, torch, dev, local, timed)
        except Exception as e:
            line["reference_experiment"] = {"error": str(e)[:200]}
    else:
        P.close()

    # ---- CPU baseline on the box's host cores (rank 0, N = 1 only)
    if rank == 0 and world == 1:
        cores = os.cpu_count() or 1
        nt_sample = w["nSteps"] if w["live"] < 2e6 else 41    # c2 / c3: the whole time loop of the shot, twice
        cpu_forward_sample(w, 21, cores)
        v0, dt0 = cpu_forward_sample(w, nt_sample, cores)
        v1, dt1 = cpu_forward_sample(w, nt_sample, cores)
        v, dt = w["live"] * (nt_sample - 1) * 2 / (dt0 + dt1), dt0 + dt1
        line["cpu_baseline"] = {"value": v, "unit": "cell-updates/s", "cores": cores, "kind": "port",
                                "sample": "2 x %d of %d forward time steps of the same workload, CPU oracle port (C, OpenMP over rows), %.1f s"
                                          % (nt_sample - 1, w["nSteps"] - 1, dt)}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


def extra_batched(args, w, Propagator, ShotSpec, torch, dev, local, world, timed):
    """Same workload with `batch` shots per launch (blockIdx.z = shot): what a multi-shot survey sees."""
    B = args.batch
    nrec = len(w["xrec"])
    is_grad = w["name"] == "c3"
    if is_grad:
        return None
    with Propagator(w["nz"], w["nx"], w["nPml"], w["nPad"], w["nSteps"], w["dz"], w["dx"], w["dt"], w["f0"],
                    max_batch=B, max_nrec=nrec, device=local, kernels=args.kernels) as P:
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["true"]])
        shots = make_shots(w, ShotSpec, B)
        fn = lambda: P.forward(shots, comps=("ett",), device_out=True)
        fn()
        t = timed(fn, 2)
        return {"shots_per_launch": B, "value": world * B * w["live"] * (w["nSteps"] - 1) * 2 / t, "unit": "cell-updates/s",
                "ms_per_step": 1e3 * t / 2}


def extra_fwi(args, Propagator, ShotSpec, torch, dev, local, world, rank, timed, sdist):
    """Multi-shot FWI gradient (C4-like: Marmousi-like 1700x350, vertical DAS fiber, nt=4001), `batch` shots per GPU,
    per-GPU gradients + misfit summed by ONE NCCL all-reduce."""
    w = workload("c3")
    B = args.batch
    zrec, xrec = np.arange(10, 340), np.full(330, 850)
    P0 = w["nPml"]
    shots = [ShotSpec(2 + P0, 20 + 26 * (rank * B + k) % 1600 + P0, zrec + P0, xrec + P0, w["stf"]) for k in range(B)]
    with Propagator(w["nz"], w["nx"], w["nPml"], w["nPad"], w["nSteps"], w["dz"], w["dx"], w["dt"], w["f0"], fiber=1,
                    max_batch=B, max_nrec=len(zrec), with_adjoint=True, device=local, kernels=args.kernels) as P:
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["true"]])
        obs = [o["ett"] for o in P.forward(shots, comps=("ett",), device_out=True)]
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["start"]])
        gstf = torch.zeros((B * world, w["nSteps"]), dtype=torch.float32, device=dev)
        ar = {}

        def fn():
            r = P.gradient(shots, obs, device=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = sdist.allreduce_gradients(r["misfit"], r["glam"], r["gmu"], r["grho"], gstf)
            e1.record()
            ar["ev"] = (e0, e1)
            return out

        fn()
        t = timed(fn, 1)
        torch.cuda.synchronize()
        f_ms, b_ms = P.last_timing()
        return {"workload": "C4-like multi-shot FWI gradient, Marmousi-like 1700x350 (padded 416x1764), nt=4001, vertical DAS fiber 330 ch",
                "shots_per_gpu": B, "n_gpus": world, "shot_gradients_per_s": world * B / t,
                "cell_updates_per_s": world * B * 3.0 * w["live"] * (w["nSteps"] - 1) / t,
                "s_per_evaluation": t, "forward_loop_ms": f_ms, "backward_loop_ms": b_ms,
                "allreduce_ms": ar["ev"][0].elapsed_time(ar["ev"][1]) if world > 1 else 0.0,
                "allreduce_bytes": int(4 * (3 * w["nz"] * w["nx"] + B * world * w["nSteps"] + 1))}


def extra_reference_experiment(Propagator, ShotSpec, torch, dev, local, timed):
    """One misfit + gradient evaluation of the reference's own experiment (notebooks/Main-001-...py: 101 x 201 grid padded to
    192 x 265, nt = 1501, 19 shots, 181 adjacent receivers) -- what one L-BFGS function evaluation costs per GPU."""
    w = workload("ref")
    P0 = w["nPml"]
    xs = np.arange(10, 191, 10)
    shots = [ShotSpec(1 + P0, int(x) + P0, w["zrec"] + P0, w["xrec"] + P0, w["stf"]) for x in xs]
    with Propagator(w["nz"], w["nx"], w["nPml"], w["nPad"], w["nSteps"], w["dz"], w["dx"], w["dt"], w["f0"],
                    max_batch=len(shots), max_nrec=len(w["xrec"]), with_adjoint=True, device=local) as P:
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["true"]])
        obs = [o["ett"] for o in P.forward(shots, comps=("ett",), device_out=True)]
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["start"]])
        fn = lambda: P.gradient(shots, obs, device=True)
        fn()
        t = timed(fn, 2) / 2
        f_ms, b_ms = P.last_timing()
        return {"workload": w["desc"] + ", 19 shots in one batch", "s_per_evaluation": t, "shot_gradients_per_s": len(shots) / t,
                "forward_loop_ms": f_ms, "backward_loop_ms": b_ms, "resident_forward_launches": P.resident_launches}


def extra_large(args, Propagator, ShotSpec, torch, dev, local, peak):
    """HBM-bound size (8000 x 2000 grid, 67 MB per field, working set >> L2): per-kernel roofline."""
    w = workload("c5s")
    nrec = len(w["xrec"])
    with Propagator(w["nz"], w["nx"], w["nPml"], w["nPad"], w["nSteps"], w["dz"], w["dx"], w["dt"], w["f0"],
                    max_batch=1, max_nrec=nrec, with_adjoint=True, device=local, kernels=args.kernels) as P:
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["true"]])
        shots = make_shots(w, ShotSpec, 1)
        obs = [o["ett"] for o in P.forward(shots, comps=("ett",), device_out=True)]
        P.set_model(*[torch.from_numpy(a).to(dev) for a in w["start"]])
        P.gradient(shots, obs, device=True)
        P.set_profile(w["nSteps"])
        P.gradient(shots, obs, device=True)
        prof = P.profile()
        f_ms, b_ms = P.last_timing()
        us = {k: 1e3 * ms / n for k, (ms, n) in prof.items()}
        fk = [k for k in ("stream_fwd", "fused_fwd", "stress_fwd", "velocity_fwd") if k in us]
        bk = [k for k in ("stream_recon", "stream_adj", "fused_recon", "fused_adj", "velocity_bwd", "stress_bwd", "velocity_adj", "stress_adj") if k in us]
        tf, tb = sum(us[k] for k in fk) * 1e-6, sum(us[k] for k in bk) * 1e-6
        af = B_FWD * w["live"] / tf / 1e9
        ab = (B_ADJ * w["live"] + B_REC * w["interior"]) / tb / 1e9
        alg = {"stream_fwd": B_FWD * w["live"], "fused_fwd": B_FWD * w["live"], "stream_adj": B_ADJ * w["live"], "fused_adj": B_ADJ * w["live"],
               "stream_recon": B_REC * w["interior"], "fused_recon": B_REC * w["interior"]}
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get("c5s", {})
        except Exception:
            tr = {}
        per = {k: {"avg_launch_us": us[k], "algorithmic_bytes_per_launch": alg[k], "achieved_GBs": alg[k] / (us[k] * 1e-6) / 1e9,
                   "frac": alg[k] / (us[k] * 1e-6) / 1e9 / peak, "traffic": tr.get(k)} for k in us if k in alg}
        return {"workload": w["desc"], "per_kernel_us": us, "per_kernel_roofline": per,
                "forward_step": {"kernels": "+".join(fk), "achieved_GBs": af, "frac": af / peak, "cell_updates_per_s": w["live"] / tf},
                "backward_step": {"kernels": "+".join(bk), "achieved_GBs": ab, "frac": ab / peak, "cell_steps_per_s": w["live"] / tb},
                "peak_GBs": peak}


if __name__ == "__main__":
    main()
