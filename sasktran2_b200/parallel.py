"""Wavelength sharding across the GPUs of one box: contiguous wavelength blocks, one process
per GPU, no exchange inside the solve, one final gather of radiances / weighting functions.

`torch.distributed` is only plumbing here (rendezvous, barrier, gather); the solve itself is the CUDA library.
"""
from __future__ import annotations

import numpy as np


def wavelength_block(nwavel: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [start, start+count) of rank `rank`: GPU g gets ceil-balanced consecutive wavelengths."""
    base, rem = divmod(int(nwavel), int(world))
    start = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return start, count


def shard_scenario(sc, rank: int, world: int):
    """Slice a Scenario to this rank's wavelength block (arrays are wavelength-slowest, so blocks are
    contiguous ranges of the caller's buffers — rust/sasktran2-rs/src/bindings/atmosphere_storage.rs:21-33)."""
    import copy

    start, count = wavelength_block(sc.nwavel, rank, world)
    sl = slice(start, start + count)
    out = copy.copy(sc)
    out.ssa = np.asfortranarray(sc.ssa[:, sl])
    out.total_extinction = np.asfortranarray(sc.total_extinction[:, sl])
    out.leg_coeff = np.asfortranarray(sc.leg_coeff[:, :, sl])
    out.albedo = np.ascontiguousarray(sc.albedo[sl])
    out.solar_irradiance = np.ascontiguousarray(sc.solar_irradiance[sl])
    out.mappings = {}
    for name, mp in sc.mappings.items():
        out.mappings[name] = {k: (np.asfortranarray(v[..., sl]) if isinstance(v, np.ndarray) and v.ndim >= 2 and
                                  v.shape[-1] == sc.nwavel else v) for k, v in mp.items()}
    return out, start, count


def gather_wavelength_blocks(local: np.ndarray, nwavel: int, wavelength_axis: int = 0, dst: int = 0):
    """Gather per-rank arrays (blocks along `wavelength_axis`) onto rank `dst` with torch.distributed.
    Works with the nccl backend (device tensors) and gloo (CPU tensors).  Returns the full array on `dst`,
    None elsewhere."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    rank = dist.get_rank()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    moved = np.ascontiguousarray(np.moveaxis(local, wavelength_axis, 0))
    rest = moved.shape[1:]
    counts = [wavelength_block(nwavel, r, world)[1] for r in range(world)]
    maxc = max(counts)
    buf = torch.zeros((maxc,) + rest, dtype=torch.float64, device=dev)
    if moved.shape[0] > 0:
        buf[: moved.shape[0]] = torch.from_numpy(moved).to(dev)
    if rank == dst:
        outs = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, outs, dst=dst)
        full = np.concatenate([o[:c].cpu().numpy() for o, c in zip(outs, counts)], axis=0)
        return np.moveaxis(full, 0, wavelength_axis)
    dist.gather(buf, None, dst=dst)
    return None
