"""Shot-parallel plumbing: the reference's contiguous shot sharding and ONE collective.

Reference: DAS_Waveform_Inversion/Ops/FWI/Src/Torch_Fwi.cpp:59-103 -- OpenMP thread i owns
GPU i and the shots [bars[i], bars[i+1]) with bars = linspace(0, nshots, ngpu+1) truncated to
int; gradients and misfit are summed on the host, and only GPU 0's grad_stf survives.
Here: one process per GPU (torch.distributed, NCCL over NVLink; gloo for the CPU tests), the
same shard bounds, and a single all-reduce of one packed buffer
    [ glam | gmu | grho | gstf (nSrc x nSteps, rows of foreign shots zero) | misfit ]
which also repairs the reference's dropped grad_stf rows.
"""
import numpy as np


def shard_bounds(group_size, ngpu):
    """Torch_Fwi.cpp:59-60,78-80: float32 linspace, then int truncation."""
    bars = np.linspace(0, group_size, ngpu + 1, dtype=np.float32)
    return [int(b) for b in bars]


def shard(shot_ids, world_size, rank):
    ids = list(np.asarray(shot_ids).reshape(-1).tolist())
    if world_size > len(ids):
        raise RuntimeError("The number of GPUs should be smaller than the number of shots!")   # Torch_Fwi.cpp:49-52
    b = shard_bounds(len(ids), world_size)
    return ids[b[rank]:b[rank + 1]]


def is_distributed():
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except ImportError:
        return False


def world():
    import torch.distributed as dist
    if is_distributed():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def pack(misfit, glam, gmu, grho, gstf):
    """One flat float32 tensor on the gradients' device; misfit travels as the last element."""
    import torch
    dev = glam.device
    parts = [glam.reshape(-1), gmu.reshape(-1), grho.reshape(-1), gstf.reshape(-1).to(dev),
             torch.tensor([misfit], dtype=torch.float32, device=dev)]
    return torch.cat([p.to(torch.float32) for p in parts])


def unpack(buf, shape, stf_shape):
    n = shape[0] * shape[1]
    m = stf_shape[0] * stf_shape[1]
    glam, gmu, grho = (buf[k * n:(k + 1) * n].view(*shape) for k in range(3))
    gstf = buf[3 * n:3 * n + m].view(*stf_shape)
    return float(buf[3 * n + m].item()), glam, gmu, grho, gstf


def allreduce_gradients(misfit, glam, gmu, grho, gstf, group=None):
    """Sum over ranks with a single all-reduce (no-op when not distributed)."""
    if not is_distributed():
        return misfit, glam, gmu, grho, gstf
    import torch.distributed as dist
    buf = pack(misfit, glam, gmu, grho, gstf)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return unpack(buf, tuple(glam.shape), tuple(gstf.shape))
