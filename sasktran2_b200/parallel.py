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


def init_nccl_comm(rank: int, world: int) -> None:
    """Create the library's NCCL communicator on every rank of an initialised torch.distributed job: rank 0 draws the
    128-byte NCCL id (sk_b200_comm_unique_id), torch.distributed only carries it to the other ranks."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from . import _lib

    buf = C.create_string_buffer(128)
    if rank == 0:
        _lib.check(_lib.lib().sk_b200_comm_unique_id(buf, 128), "comm_unique_id")
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    _lib.check(_lib.lib().sk_b200_comm_init(raw, int(rank), int(world)), "comm_init")


class SharedResult:
    """Result arrays of one spectrum in POSIX shared memory, visible to every rank of a one-box job: rank 0 is "the
    caller" that ends up holding radiances and weighting functions; every rank writes its wavelength block straight
    into them (N concurrent device -> host copies into disjoint slices of one buffer).  The
    segment is page-locked in every process (sk_b200_host_register) so that the copies run at PCIe speed."""

    def __init__(self, shapes: dict, rank: int, tag: str, barrier, pin: bool = True):
        from multiprocessing import shared_memory

        from . import _lib

        self.rank = rank
        self._pinned = False
        sizes = {k: int(np.prod(s)) * 8 for k, s in shapes.items()}
        total = sum(sizes.values())
        name = f"sk_b200_{tag}"
        if rank == 0:
            try:
                old = shared_memory.SharedMemory(name=name)
                old.close()
                old.unlink()
            except FileNotFoundError:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=max(total, 8))
        barrier()
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=name)
            # Python < 3.13 registers attached segments with the resource tracker as if this process owned them and
            # then warns about a "leak" (and tries a second unlink) at exit; rank 0 owns and unlinks the segment
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.arrays, off = {}, 0
        for k, s in shapes.items():
            self.arrays[k] = np.ndarray(s, dtype=np.float64, buffer=self.shm.buf, offset=off)
            off += sizes[k]
        self._base = np.ndarray((max(total, 8),), dtype=np.uint8, buffer=self.shm.buf)
        if pin:
            rc = _lib.lib().sk_b200_host_register(self._base.ctypes.data, self._base.nbytes)
            self._pinned = rc == 0
        barrier()

    @property
    def pinned(self) -> bool:
        return self._pinned

    def close(self, barrier) -> None:
        from . import _lib

        if self._pinned:
            _lib.lib().sk_b200_host_unregister(self._base.ctypes.data)
        self.arrays = {}
        self._base = None
        barrier()
        try:
            self.shm.close()
        except BufferError:
            pass
        if self.rank == 0:
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass
