#!/usr/bin/env python
"""The reference's anomaly inversions end to end on this stack (Main-00{1,2,3}-FWI-Anomaly-*.py):

    python tools/main_fwi_anomaly.py --problem 001 --exp_name /tmp/anomaly --generate_data
    python tools/main_fwi_anomaly.py --problem 001 --exp_name /tmp/anomaly --nIter 100 --ngpu 1

Prints the L-BFGS trace (`At iterate k  f= ...  |proj g|= ...`) next to the one logged in the reference's notebook
and the wall time per function + gradient evaluation."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "sep-2023_b200")]
from sepfwi import drivers  # noqa: E402

# first iterates logged in notebooks/00{1,2,3}-*.ipynb (cell 7 output)
NOTEBOOK_F = {"001": [1.51116e4, 1.13748e4, 3.05521e3, 2.11215e3, 1.41392e3, 1.03573e3],
              "002": [1.08179e4], "003": [3.52002e4]}

ap = argparse.ArgumentParser()
ap.add_argument('--problem', default='001', choices=['001', '002', '003'])
ap.add_argument('--generate_data', action='store_true')
ap.add_argument('--exp_name', type=str, default='/tmp/sepfwi-anomaly')
ap.add_argument('--nIter', type=int, default=5)
ap.add_argument('--ngpu', type=int, default=1)
ap.add_argument('--model_device', default='cuda', choices=['cpu', 'cuda'],
                help="where the model tensors of the FWI module live: 'cpu' is what the reference does (its op only takes CPU tensors); "
                     "'cuda' keeps the parameterisation maps and the gradients on the device")
ap.add_argument('--ref_race_compat', action='store_true', help="reproduce the reference's lost seam update in the residual injection")
args = ap.parse_args()

# one process per GPU under torchrun: shots are sharded over the ranks inside fwi_ops (sepfwi.dist), gradients and misfit are
# summed by one NCCL all-reduce, every rank then takes the identical L-BFGS step
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
device = None
if world == 1 and args.model_device == 'cuda':
    import torch
    if torch.cuda.is_available():
        device = torch.device('cuda', 0)
if world > 1:
    import torch
    import torch.distributed as td
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    td.init_process_group("nccl", device_id=device)

prob = drivers.anomaly_problem(args.problem)
if rank == 0:
    files = drivers.write_files(prob, args.exp_name, ref_race_compat=args.ref_race_compat or None)
if world > 1:
    td.barrier()
    files = dict(para=os.path.join(args.exp_name, "para_file.json"), survey=os.path.join(args.exp_name, "survey_file.json"),
                 data=os.path.join(args.exp_name, "Data"))
if args.generate_data:
    drivers.generate_data(prob, files, ngpu=args.ngpu, device=device)
    if world > 1:
        td.barrier()
        td.destroy_process_group()
    if rank == 0:
        print('End of Data Generation')
    sys.exit(0)
fwi, obj, log = drivers.invert(prob, files, nIter=args.nIter, ngpu=args.ngpu, device=device)
if world > 1:
    td.destroy_process_group()
if rank != 0:
    sys.exit(0)
ref = NOTEBOOK_F[args.problem]
for k, f, g, t in log:
    print("At iterate %4d    f= %.5E    |proj g|= %.5E    (%.2f s)%s" % (k, f, g, t, "    notebook f= %.5E" % ref[k] if k < len(ref) else ""))
print("%d function + gradient evaluations, %.3f s each" % (obj.nfev, log[-1][3] / max(1, obj.nfev)))
