"""Multi-GPU render: one process per GPU (torchrun), ONE collective per frame (NCCL over NVLink; gloo on CPU for the host-logic
tests).  Two partitions:

  samples (FP32 renderers, the default when samples_per_pixel >= world): every rank renders ALL pixels but only its share of the
      samples into 64-bit fixed-point accumulators; one reduce (integer sum) to rank 0, resolve there.  Integer sums commute, so
      the image is the single-GPU image bit for bit, and the ranks' loads differ only by the rounding of spp / world.
  tiles (f64, and the fallback): image tiles dealt to the ranks, one gather of the tile buffers to rank 0, untile + resolve there.

The reference has no distributed path; its only parallelism is rayon over pixels
(shared/src/camera.rs:353).  Pixels are independent and the RNG is keyed by the absolute pixel index, so
the image does not depend on the number of GPUs or on which GPU renders which tile.

Partition (same constants as include/rtw.h): tiles of 16x16 pixels; tile (tx, ty) has slot
k = ty * tiles_x + (tx + rot(ty)) % tiles_x (every tile row rotated pseudo-randomly, so that no rank ends up with
whole tile columns); slot k is owned by rank k % world and is that rank's local tile k // world.  Every rank holds
tiles_per_rank = ceil(tiles_total / world) local tiles (the tail is padding and stays zero), so the gather
moves equally sized buffers and needs no size exchange.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist

TILE_W = TILE_H = 16


def tiles_xy(width: int, height: int):
    return (width + TILE_W - 1) // TILE_W, (height + TILE_H - 1) // TILE_H


def tiles_total(width: int, height: int) -> int:
    tx, ty = tiles_xy(width, height)
    return tx * ty


def tiles_per_rank(width: int, height: int, world: int) -> int:
    return (tiles_total(width, height) + world - 1) // world


def tile_row_rotation(ty: int, tiles_x: int) -> int:
    """csrc/rtw_device.cuh: tile_row_rotation (32-bit arithmetic)."""
    return (((ty * 0x9E3779B1) & 0xFFFFFFFF) >> 15) % tiles_x


def tile_owner(slot: int, world: int):
    """(rank, local index) of tile slot `slot`."""
    return slot % world, slot // world


def local_tile_ids(width: int, height: int, rank: int, world: int):
    """Tile slots of this rank's local tiles, in local order (padding slots excluded)."""
    return list(range(rank, tiles_total(width, height), world))


def tile_rect(slot: int, width: int, height: int):
    """Pixel rectangle (i0, j0, i1, j1) of the tile in slot `slot`, clipped to the image; j = 0 is the bottom row."""
    tiles_x, _ = tiles_xy(width, height)
    ty, c = divmod(slot, tiles_x)
    tx = (c - tile_row_rotation(ty, tiles_x)) % tiles_x
    i0, j0 = tx * TILE_W, ty * TILE_H
    return i0, j0, min(i0 + TILE_W, width), min(j0 + TILE_H, height)


def init_from_env(backend: Optional[str] = None):
    """Join the process group torchrun described (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local_rank


def gather_tiles(local_tiles: torch.Tensor, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """The one collective of the render: gather every rank's [tiles_per_rank, 16, 16, 3] buffer on `dst`.
    Returns [world, tiles_per_rank, 16, 16, 3] on dst and None elsewhere."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_tiles.unsqueeze(0)
    rank = dist.get_rank(group)
    if rank == dst:
        out = torch.empty((world,) + tuple(local_tiles.shape), dtype=local_tiles.dtype, device=local_tiles.device)
        dist.gather(local_tiles, gather_list=list(out.unbind(0)), dst=dst, group=group)
        return out
    dist.gather(local_tiles, gather_list=None, dst=dst, group=group)
    return None


def sample_range(spp: int, rank: int, world: int):
    """This rank's samples [begin, begin + count) of every pixel: contiguous shares that differ by at most one sample."""
    begin, end = rank * spp // world, (rank + 1) * spp // world
    return begin, end - begin


def accum_words(width: int, height: int) -> int:
    """int64 words of one rank's accumulator block: [slots][3] u64 radiance sums followed by [slots] u32 poison words."""
    slots = tiles_total(width, height) * TILE_W * TILE_H
    return 3 * slots + (slots + 1) // 2


def reduce_accum(block: torch.Tensor, dst: int = 0, group=None):
    """The one collective of the sample partition: integer SUM of every rank's accumulator block on `dst` (in place).  The poison
    words travel in the same buffer: their six flags sit in separate 4-bit fields, so adding the words of up to 15 ranks keeps a
    flag set iff any rank set it."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(block, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return block


class DistributedRenderer:
    """Holds the per-rank device buffers so repeated renders (bench steps) allocate nothing."""

    def __init__(self, scene, camera, opts, rank: int, world: int, want_sum: bool = False, want_rgb8: bool = True, partition: Optional[str] = None):
        from . import api
        self.api = api
        self.scene, self.camera, self.opts, self.rank, self.world = scene, camera, opts, rank, world
        w, h = camera.image_width, camera.image_height
        spp = camera.pod.samples_per_pixel
        fixed_point = opts.precision == api.RTW_F32 and not (opts.flags & api._lib.RTW_FLAG_LANE_PER_PIXEL)
        if partition is None:
            partition = "samples" if (fixed_point and world > 1 and world <= 15 and spp >= world) else "tiles"
        if partition == "samples" and not fixed_point:
            raise ValueError("the sample partition needs a fixed-point FP32 renderer")
        if partition == "samples" and world > 15:
            # the poison words are SUMMED by the reduce: each flag owns a 4-bit field, 16 ranks would carry into the next one
            raise ValueError("the sample partition adds the ranks' poison words: at most 15 ranks (use partition='tiles')")
        self.partition = partition
        self.comm = None
        self.rgb_sum = torch.zeros((h, w, 3), dtype=torch.float64, device="cuda") if (want_sum and rank == 0) else None
        self.rgb8 = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda") if (want_rgb8 and rank == 0) else None
        if partition == "samples":
            self.slots = tiles_total(w, h) * TILE_W * TILE_H
            self.block = torch.zeros(accum_words(w, h), dtype=torch.int64, device="cuda")
            self.sample_begin, self.sample_count = sample_range(spp, rank, world)
            self.local = None
        else:
            self.tpr = tiles_per_rank(w, h, world)
            dt = torch.float32 if opts.precision == api.RTW_F32 else torch.float64
            self.local = torch.zeros((self.tpr, TILE_H, TILE_W, 3), dtype=dt, device="cuda")

    def use_library_collective(self):
        """Run the frame's collective INSIDE the C library (rtw_comm_* / rtw_render_rank_device: NCCL from C) instead of through
        torch.distributed: rank 0 creates the NCCL unique id, torch.distributed only carries its 128 bytes to the other ranks.
        The partition the library picks is the one chosen here (same rule), so the images are the same bit for bit."""
        if self.world == 1 or self.comm is not None:
            return self
        ids = [self.api.Comm.unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        self.comm = self.api.Comm(ids[0], self.rank, self.world)
        return self

    def render_local(self, opts=None, want_stats: bool = False):
        """This rank's share of the frame (kernels only, no collective) on torch's current stream."""
        stream = torch.cuda.current_stream().cuda_stream
        opts = opts or self.opts
        if self.partition == "samples":
            return self.scene.render_samples_device(self.camera, opts, self.sample_begin, self.sample_count, self.block.data_ptr(),
                                                    self.block.data_ptr() + 8 * 3 * self.slots, stream, want_stats=want_stats)
        return self.scene.render_tiles_device(self.camera, opts, self.rank, self.world, self.local.data_ptr(), stream, want_stats=want_stats)

    def combine(self):
        """The frame's one collective + the resolve on rank 0."""
        stream = torch.cuda.current_stream().cuda_stream
        cam = self.camera.pod
        sum_ptr = self.rgb_sum.data_ptr() if self.rgb_sum is not None else 0
        rgb8_ptr = self.rgb8.data_ptr() if self.rgb8 is not None else 0
        if self.partition == "samples":
            reduce_accum(self.block, 0)
            if self.rank == 0:
                self.api.resolve_accum_device(self.block.data_ptr(), self.block.data_ptr() + 8 * 3 * self.slots, cam.image_width,
                                              cam.image_height, cam.samples_per_pixel, sum_ptr, rgb8_ptr, stream)
            return
        allt = gather_tiles(self.local, 0)
        if self.rank == 0:
            self.api.untile_resolve_device(allt.data_ptr(), self.opts.precision, cam.image_width, cam.image_height, self.world,
                                           cam.samples_per_pixel, sum_ptr, rgb8_ptr, stream)
            self._keep = allt       # keep the gathered buffer alive until the stream has consumed it

    def render(self, want_stats: bool = False):
        """One frame: this rank's share, the collective, the resolve on rank 0.  Work is enqueued on torch's current stream; returns
        the kernel stats dict when asked (that syncs)."""
        if self.comm is not None:
            stream = torch.cuda.current_stream().cuda_stream
            sum_ptr = self.rgb_sum.data_ptr() if self.rgb_sum is not None else 0
            rgb8_ptr = self.rgb8.data_ptr() if self.rgb8 is not None else 0
            return self.scene.render_rank_device(self.camera, self.opts, self.comm, sum_ptr, rgb8_ptr, stream, want_stats=want_stats)
        st = self.render_local(want_stats=want_stats)
        self.combine()
        return st

    def check(self):
        """Joins the asynchronous renders of this rank and raises if a path did what makes the reference panic (a light sample from
        an empty lights list): the *_device entry points only report it when they are asked for stats."""
        return self.scene.sync()
