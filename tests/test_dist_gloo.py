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


def _poison_word(nan_rgb, inf_rgb):
    """csrc/rtw_kernels.cuh: poison_nan(c) = 1 << 4c, poison_inf(c) = 1 << (12 + 4c)."""
    return sum(1 << (4 * c) for c in range(3) if nan_rgb[c]) | sum(1 << (12 + 4 * c) for c in range(3) if inf_rgb[c])


def _accum_worker(rank, world, port, w, h, spp, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from ray_tracing_weekend_b200 import dist as D
    D.init_from_env(backend="gloo")
    slots = D.tiles_total(w, h) * 256
    block = torch.zeros(D.accum_words(w, h), dtype=torch.int64)
    # every rank "renders" its samples: each sample adds (slot + 1) * 2^32 / 8 to the red accumulator of its slot
    b, c = D.sample_range(spp, rank, world)
    acc = block[: 3 * slots].view(slots, 3)
    acc[:, 0] = (torch.arange(slots, dtype=torch.int64) + 1) * (1 << 29) * c
    poison = block[3 * slots:].view(torch.int32)
    poison[rank] = _poison_word((1, 0, 0), (0, 0, 1))                      # slot `rank`: NaN in red, overflow in blue
    poison[7] = _poison_word((0, 1, 0), (0, 0, 0))                         # slot 7: every rank flags green
    D.reduce_accum(block, 0)
    if rank == 0:
        ok = bool((acc[:, 0] == (torch.arange(slots, dtype=torch.int64) + 1) * (1 << 29) * spp).all())
        nib = lambda word, k: (int(word) >> (4 * k)) & 15
        ok &= all(nib(poison[r], 0) == 1 and nib(poison[r], 5) == 1 and nib(poison[r], 1) == 0 for r in range(world) if r != 7)
        ok &= nib(poison[7], 1) == world and nib(poison[7], 2) == 0 and nib(poison[7], 3) == 0
        q.put((ok, b, c))
    else:
        q.put((True, b, c))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,spp", [(2, 500), (3, 10)])
def test_sample_partition_reduce(world, spp):
    """The sample partition's host logic on CPU: contiguous sample shares that cover [0, spp), one integer reduce of the accumulator
    block, poison flags surviving the sum in their 4-bit fields."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_accum_worker, args=(r, world, port, 100, 70, spp, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = [q.get(timeout=5) for _ in range(world)]
    assert all(g[0] for g in got)
    ranges = sorted((g[1], g[2]) for g in got)
    assert ranges[0][0] == 0 and all(ranges[i][0] + ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    assert ranges[-1][0] + ranges[-1][1] == spp and max(c for _, c in ranges) - min(c for _, c in ranges) <= 1
