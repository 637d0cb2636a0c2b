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


class ContextGroup:
    """Multi-GPU from ONE host thread (acm_comm_init_all): one Context per device of this process, bound
    into a group whose *_multi entry points drive every GPU from the calling thread -- the form a
    single-threaded, launcher-less host like the reference's converter `main`
    (bin/camera_converter.rs:127-343) binds.  Shards are lists with one `Points` per context."""

    def __init__(self, devices):
        import ctypes as C
        from . import _native as N
        self._C, self._N, self._lib = C, N, N.lib
        self.ctxs = [Context(d) for d in devices]
        self._arr = (C.c_void_p * len(self.ctxs))(*[c.handle.value for c in self.ctxs])
        self.ctxs[0].check(self._lib.acm_comm_init_all(self._arr, len(self.ctxs)))
        self._open = True

    def __len__(self):
        return len(self.ctxs)

    def _handles(self, pts):
        C = self._C
        if len(pts) != len(self.ctxs):
            raise ValueError("one shard per context expected")
        return (C.c_void_p * len(pts))(*[p.handle.value for p in pts])

    def linearize(self, cam_block, residual_kind, xyz, uv):
        ne = self._N.NormalEquations()
        self.ctxs[0].check(self._lib.acm_linearize_multi(self._arr, len(self), self._C.byref(cam_block), residual_kind, self._handles(xyz),
                                                          self._handles(uv), self._C.byref(ne)))
        return ne

    def lm_solve(self, cam_block, residual_kind, xyz, uv, lower=None, upper=None, config=None):
        C, N = self._C, self._N
        P = cam_block.n_params
        lo = (C.c_double * P)(*lower) if lower is not None else None
        hi = (C.c_double * P)(*upper) if upper is not None else None
        out = (C.c_double * N.ACM_MAX_PARAMS)()
        res = N.LMResult()
        self.ctxs[0].check(self._lib.acm_lm_solve_multi(self._arr, len(self), C.byref(cam_block), residual_kind, self._handles(xyz), self._handles(uv),
                                                         lo, hi, C.byref(config) if config is not None else None, out, C.byref(res)))
        return [out[i] for i in range(P)], res

    def linear_estimation(self, cam_block, xyz, uv):
        self.ctxs[0].check(self._lib.acm_linear_estimation_multi(self._arr, len(self), self._C.byref(cam_block), self._handles(xyz), self._handles(uv)))
        return cam_block

    def reprojection_error(self, cam_block, xyz, uv):
        pe = self._N.ProjectionError()
        self.ctxs[0].check(self._lib.acm_reprojection_error_multi(self._arr, len(self), self._C.byref(cam_block), self._handles(xyz), self._handles(uv),
                                                                   self._C.byref(pe)))
        return pe

    def sample_points(self, cam_block, n_requested):
        """Shard i of the grid on context i -> (uv shards, xyz shards, kept counts)."""
        from .runtime import Points
        C = self._C
        n = len(self)
        uv = (C.c_void_p * n)(); xyz = (C.c_void_p * n)(); kept = (C.c_size_t * n)()
        self.ctxs[0].check(self._lib.acm_sample_points_multi(self._arr, n, C.byref(cam_block), n_requested, uv, xyz, kept))
        U = [Points(c, 2, int(kept[i]), _handle=C.c_void_p(uv[i])) for i, c in enumerate(self.ctxs)]
        X = [Points(c, 3, int(kept[i]), _handle=C.c_void_p(xyz[i])) for i, c in enumerate(self.ctxs)]
        return U, X, [int(k) for k in kept]

    def close(self):
        if getattr(self, "_open", False):
            self._lib.acm_comm_destroy_all(self._arr, len(self.ctxs))
            for c in self.ctxs:
                c.close()
            self._open = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
