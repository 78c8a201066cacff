"""Multi-GPU plumbing: independent sequences shard one (or more) per rank; the only collective is
one broadcast of the calibration / parameter block.  `torch.distributed` is
plumbing here: NCCL on GPUs, gloo in the CPU tests."""
import numpy as np

from .capi import Params


def shard_sequences(n_sequences, rank, world_size):
    """sequence i -> rank i mod world_size (config 5: 16 sequences on 8 GPUs = two per rank)"""
    return [s for s in range(n_sequences) if s % world_size == rank]


def broadcast_params(params, src=0, device="cpu"):
    """Broadcast lvt_parameters from `src` (every field is exactly representable in float64)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return params
    blob = params.to_array() if dist.get_rank() == src else np.zeros(len(Params._fields_), np.float64)
    t = torch.from_numpy(blob).to(device)
    dist.broadcast(t, src=src)
    return Params.from_array(t.cpu().numpy())


def gather_trajectories(local, device="cpu"):
    """all ranks -> {sequence id: poses (n x 12)} on every rank; `local` = {sequence id: array}."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, {k: np.asarray(v) for k, v in local.items()})
    merged = {}
    for d in out:
        merged.update(d)
    return merged
