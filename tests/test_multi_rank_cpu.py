"""World-size-2 `gloo` test (CPU) of the N > 1 host logic: contiguous shard ranges, the byte
broadcast that carries the NCCL unique id, and the algebra the data path relies on -- the normal
equations of the whole set equal the rank-ordered sum of the per-shard normal equations (the
per-shard values come from the oracle here; on the GPU box the same sum is the NCCL all-reduce
inside libacm)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from apex_camera_models_b200.distributed import broadcast_bytes, env_rank_world, shard_range
    from oracle import oracle as O
    assert env_rank_world() == (rank, rank, world)
    # 1. unique-id style broadcast
    payload = bytes(range(128)) if rank == 0 else None
    got = broadcast_bytes(payload, 128, src=0)
    assert got == bytes(range(128))
    # 2. shard -> per-rank normal equations -> all-reduce == whole
    ds = O.make_model(O.DS, [348.112754378549, 347.1109973814674, 365.8121721753254, 249.3555778487899, 0.5657413673629862, -0.24425190195168348], 752, 480)
    lo, hi = shard_range(n, rank, world)
    xyz = O.synth_points3(0xACE50004, lo, hi - lo, float(np.cos(np.deg2rad(85.0))), False)  # counter-based: shard == slice of the whole
    uv, st = O.project(ds, xyz)
    uv = uv + 0.2
    H, g, cost, nv = O.linearize(ds, O.RES_PIXEL, xyz, uv)
    vec = torch.from_numpy(np.concatenate([H.ravel(), g, [cost, float(nv)]]))
    dist.all_reduce(vec)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), vec.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_two_gloo(tmp_path, O):
    n, world = 20_001, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    red = np.load(tmp_path / "reduced.npy")
    ds = O.make_model(O.DS, [348.112754378549, 347.1109973814674, 365.8121721753254, 249.3555778487899, 0.5657413673629862, -0.24425190195168348], 752, 480)
    xyz = O.synth_points3(0xACE50004, 0, n, float(np.cos(np.deg2rad(85.0))), False)
    uv, _ = O.project(ds, xyz)
    H, g, cost, nv = O.linearize(ds, O.RES_PIXEL, xyz, uv + 0.2)
    whole = np.concatenate([H.ravel(), g, [cost, float(nv)]])
    assert red[-1] == nv == n
    assert np.allclose(red, whole, rtol=1e-12, atol=1e-9)
