"""Multi-GPU plumbing: one process per GPU (torchrun), points sharded by contiguous ranges, the
only collective is the all-reduce of the normal equations inside libacm (NCCL over NVLink).
torch.distributed is used for the rendezvous only: it carries the 128-byte NCCL unique id."""
from __future__ import annotations

import os

from .runtime import Context


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of rank `rank`: [r*n/G, (r+1)*n/G)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    return (n * rank) // world, (n * (rank + 1)) // world


def env_rank_world() -> tuple[int, int, int]:
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def bind_to_gpu_numa(local_rank: int) -> list[int] | None:
    """Pin this process to the CPUs the driver reports as local to GPU `local_rank` (NVML affinity), so that
    pinned host buffers allocated afterwards land on that NUMA node and the H2D streams of the ranks do not
    share one socket's memory controllers / inter-socket link.  Returns the CPU list, or None when NVML or
    the affinity call is unavailable (nothing changed)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        # honour CUDA_VISIBLE_DEVICES: map the local ordinal to the physical index when it is a plain list
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = local_rank
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                index = int(ids[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Broadcast a fixed-size byte string over the default torch.distributed group (any backend)."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def attach_communicator(ctx: Context) -> int:
    """Give `ctx` an NCCL communicator spanning the torch.distributed world. Returns world size."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return 1
    uid = Context.comm_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 128, src=0)
    ctx.comm_init_rank(world, rank, uid)
    return world


def all_gather_bytes(payload: bytes) -> bytes:
    """Concatenation over ranks of equal-sized byte strings (default torch.distributed group)."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)


def attach_peers(ctx: Context) -> int:
    """NVLink peer exchange: every rank exports its exchange buffer (CUDA IPC handle), the handles
    are all-gathered over torch.distributed and opened.  After this the cross-rank sum of the normal
    equations and the LM step run inside the streaming kernel (no NCCL call per iteration)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return 1
    handles = all_gather_bytes(ctx.peer_export())
    ctx.peer_attach(world, rank, handles)
    dist.barrier()
    return world
