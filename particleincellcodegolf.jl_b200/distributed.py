"""Multi-GPU plumbing: one process per GPU (torchrun), particles shard by contiguous global index
range, every rank holds a full grid, rho is summed with one NCCL all-reduce per sweep inside
libpicgolf (the CPU analogue is the per-thread grids + `phi .= sum(ns, dims=3)` of
src/Electrostatic2D3V.jl:114,126-141).  torch.distributed is used only to ship the 128-byte NCCL
unique id from rank 0 to the other ranks."""
from __future__ import annotations

import os

from . import PIC, comm_unique_id


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_bytes(payload, nbytes: int, src: int = 0) -> bytes:
    """Broadcast a small byte string over the default torch.distributed group (gloo or nccl)."""
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def connect(pic: PIC) -> None:
    """Create the handle's NCCL communicator (collective over all ranks of the default group)."""
    import torch.distributed as dist

    if pic.cfg.nranks == 1:
        return
    uid = comm_unique_id() if dist.get_rank() == 0 else None
    uid = broadcast_bytes(uid, 128, src=0)
    pic.comm_init(uid)
