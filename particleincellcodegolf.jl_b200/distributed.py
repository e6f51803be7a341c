"""Multi-GPU plumbing: one process per GPU (torchrun), particles shard by contiguous global index
range, every rank holds a full grid, rho is summed with one NCCL all-reduce per sweep inside
libpicgolf (the CPU analogue is the per-thread grids + `phi .= sum(ns, dims=3)` of
src/Electrostatic2D3V.jl:114,126-141).  torch.distributed is used only to ship the 128-byte NCCL
unique id from rank 0 to the other ranks and to all-gather the 64-byte cudaIpc handles of the peer-memory reduction
(pg_peer.cuh), which replaces the per-sweep all-reduce of the 1D schemes."""
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


def connect(pic: PIC, peer=None) -> None:
    """Create the handle's NCCL communicator (collective over all ranks of the default group) and the peer-memory reduction
    of the 1D charge grids (pg_peer.cuh).  peer=None: PICGOLF_PEER=1 / 0 forces it on / off; otherwise it is on for the
    Gaussian fixed point, whose device-driven sweep loop (a CUDA-graph WHILE node) cannot contain NCCL calls, and off
    (one ncclAllReduce per step) for the schemes with a single solve per step."""
    import torch.distributed as dist

    if pic.cfg.nranks == 1:
        return
    uid = comm_unique_id() if dist.get_rank() == 0 else None
    uid = broadcast_bytes(uid, 128, src=0)
    pic.comm_init(uid)
    if peer is None:
        env = os.environ.get("PICGOLF_PEER")
        peer = (env != "0") if env is not None else pic.cfg.scheme == 3  # GAUSS_FIXEDPOINT
    if peer and pic.cfg.scheme not in (4, 5, 6):  # 2D3V and the Simpson schemes sum their grids with NCCL
        connect_peers(pic)


def gather_bytes(payload: bytes) -> bytes:
    """All-gather equal-length byte strings over the default group, concatenated in rank order."""
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)


def connect_peers(pic: PIC) -> bool:
    """Set up the peer-memory reduction of the 1D charge grids (pg_peer.cuh): exchange the cudaIpc handles and open
    them.  All ranks of one node; if any rank cannot (no peer access, more than 16 ranks), every rank stays on NCCL."""
    import torch
    import torch.distributed as dist

    handle, ok = b"\0" * 64, 1
    try:
        handle = pic.peer_export()
    except Exception:
        ok = 0
    handles = gather_bytes(handle + bytes([ok]))
    n = dist.get_world_size()
    if not all(handles[65 * q + 64] for q in range(n)):
        return False
    try:
        pic.peer_connect(b"".join(handles[65 * q:65 * q + 64] for q in range(n)))
    except Exception:
        ok = 0
    # the switch must be collective: a rank that failed to open a handle keeps everyone on NCCL
    flags = gather_bytes(bytes([ok]))
    if not all(flags):
        raise RuntimeError("peer-memory reduction: some ranks could not open the cudaIpc handles (unset PICGOLF_PEER)")
    return True
