"""World-size-2 (and 3) gloo test of the multi-GPU host logic on CPU: tile ownership, the single gather
and the layout the untile kernel expects.  The tiles are fabricated from a known image (no rendering)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _untile_host(allt, w, h, world):
    """numpy statement of untile_resolve_kernel's addressing (csrc/rtw_kernels.cuh)."""
    from ray_tracing_weekend_b200 import dist as D
    out = np.zeros((h, w, 3), dtype=allt.dtype)
    for k in range(D.tiles_total(w, h)):
        r, l = D.tile_owner(k, world)
        i0, j0, i1, j1 = D.tile_rect(k, w, h)
        out[j0:j1, i0:i1] = allt[r, l, : j1 - j0, : i1 - i0]
    return out


def _worker(rank, world, port, w, h, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from ray_tracing_weekend_b200 import dist as D
    r, wd, _ = D.init_from_env(backend="gloo")
    assert (r, wd) == (rank, world)
    img = np.arange(h * w * 3, dtype=np.float32).reshape(h, w, 3)          # the "rendered" image every rank agrees on
    tpr = D.tiles_per_rank(w, h, world)
    local = torch.zeros((tpr, 16, 16, 3), dtype=torch.float32)
    for n, k in enumerate(D.local_tile_ids(w, h, rank, world)):
        i0, j0, i1, j1 = D.tile_rect(k, w, h)
        local[n, : j1 - j0, : i1 - i0] = torch.from_numpy(img[j0:j1, i0:i1])
    allt = D.gather_tiles(local, 0)
    if rank == 0:
        assert allt.shape == (world, tpr, 16, 16, 3)
        q.put(bool(np.array_equal(_untile_host(allt.numpy(), w, h, world), img)))
    else:
        assert allt is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,w,h", [(2, 100, 70), (3, 64, 48), (2, 3, 2)])
def test_gather_of_partitioned_tiles_rebuilds_the_image(world, w, h):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, w, h, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
