#!/usr/bin/env python
"""bench.py — Mrays/s / Mpaths/s of the CUDA render path on BASELINE config C2:
random-spheres (scenes::simple, seeded) 1920x1080, 500 spp, depth 50, N B200s; the wavefront renderer is timed
(the faster of the two FP32 renderers), the megakernel is measured beside it outside the timed region.

A step is one full render of that frame: every rank renders its share (N > 1: its samples of every pixel into fixed-point
accumulators, one NCCL reduce; N = 1 / f64: rtw_render_tiles_device, tiles interleaved across
ranks), one NCCL gather of the tile buffers on rank 0, untile + resolve there.  Contract: see the task
brief — prints ONE JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261018
WIDTH, HEIGHT, SPP, DEPTH = 1920, 1080, 500, 50
WORKLOAD = f"random-spheres (scenes::simple seed {SEED}, reference ground plane) {WIDTH}x{HEIGHT}, {SPP} spp, depth {DEPTH}"

# Algorithmic FP32 flop per event (SURVEY §8d / DESIGN.md §Roofline); FMA = 2.
FLOP = dict(ray_setup=8, node_visit=2 * 22, sphere_test=21, hit_record=20, lambertian=84, light_test=34, metal=46,
            dielectric=45, path_setup=30)


def algorithmic_flops(st):
    return (st["rays"] * (FLOP["ray_setup"]) + st["node_visits"] * FLOP["node_visit"] + st["sphere_tests"] * FLOP["sphere_test"]
            + (st["rays"] - st["missed"]) * FLOP["hit_record"] + st["lambertian"] * FLOP["lambertian"]
            + st["light_tests"] * FLOP["light_test"] + st["metal"] * FLOP["metal"] + st["dielectric"] * FLOP["dielectric"]
            + st["paths"] * FLOP["path_setup"])


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0), "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2]))
                except ValueError:
                    continue
                for n, v in zip(names, c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        os.unlink(self.path)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_reference_run(steps, warmup, sample_spp, faithful=False):
    """The reference's CPU renderer restated (oracle/, C++ f64 port — the Rust reference cannot be built here),
    all host threads, on a bounded sample of the workload: the full 1080p frame at `sample_spp` spp."""
    from oracle import pyoracle as O
    desc = O.scene_simple(SEED)
    sc = O.Scene(desc)
    cam = O.camera_for(desc, WIDTH, HEIGHT, sample_spp, DEPTH)
    opt = O.options(seed=SEED, faithful_bvh=faithful)
    for _ in range(warmup):
        sc.render(cam, opt)
    secs, rays, paths = 0.0, 0, 0
    for _ in range(steps):
        _, s, cnt, _ = sc.render(cam, opt)
        secs += s; rays += cnt["rays"]; paths += cnt["paths"]
    return dict(seconds=secs, rays=rays, paths=paths, cores=O.hardware_threads(),
                sample=f"{WIDTH}x{HEIGHT} at {sample_spp} spp (of {SPP}), depth {DEPTH}, same scene/seed; "
                       f"{'faithful (boxes recomputed per visit, bvh.rs:145-151)' if faithful else 'cached node boxes'}")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup, sample_spp=1)
    mrays = r["rays"] / r["seconds"] * 1e-6
    line = dict(impl="reference", metric="Mrays/s", value=mrays, unit="Mrays/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=r["seconds"] / args.steps * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                data="synthetic", config=dict(workload=WORKLOAD, sample="each step renders 1 of the frame's 500 samples per pixel (1/500 of the frame): Mrays/s is a rate, "
                                              "so it compares with the CUDA arm's full frames",
                                              note="C++ restatement of the reference's CPU renderer (oracle/), not the reference "
                                              "binary: no Rust toolchain in this image"),
                mpaths_per_s=r["paths"] / r["seconds"] * 1e-6,
                cpu_baseline=dict(value=mrays, unit="Mrays/s", cores=r["cores"], kind="port", sample=r["sample"] + "; each step is one such frame"),
                e2e=dict(value=mrays, unit="Mrays/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    _emit(line)


def run_cuda(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import ray_tracing_weekend_b200 as R
    from ray_tracing_weekend_b200 import dist as D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA path has no CPU fallback (use --impl reference for the CPU baseline)")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the one JSON line
    rank, world, local_rank = D.init_from_env()
    if world != args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torchrun --nproc-per-node {args.gpus}")
    R.load()
    world_h, lights_h, cb = R.scenes.simple(SEED)
    cam = (cb.with_vfov(40.).with_aspect_ratio(WIDTH / HEIGHT).with_max_depth(DEPTH).with_image_width(WIDTH).with_image_height(HEIGHT)
           .with_samples_per_pixel(SPP).build())
    base_flags = R.RTW_FLAG_LANE_PER_PIXEL if args.lane_per_pixel else 0
    mode = R.RTW_WAVEFRONT if args.mode == "wavefront" else R.RTW_MEGAKERNEL
    opts = R.RenderOptions(seed=SEED, precision=R.RTW_F32, mode=mode, flags=base_flags)
    scene = R.Scene(world_h, lights_h)
    renderer = D.DistributedRenderer(scene, cam, opts, rank, world, want_sum=False, want_rgb8=True)
    if world > 1 and not args.torch_collective:
        renderer.use_library_collective()       # the frame's one collective runs inside librtw_cuda.so (rtw_render_rank_device: NCCL from C)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")          # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # event counts of one step (deterministic: counter-based RNG) — an untimed pass with counters on
    copts = R.RenderOptions(seed=SEED, precision=R.RTW_F32, mode=mode, flags=R.RTW_FLAG_COUNT_EVENTS | base_flags)
    cnt = renderer.render_local(copts, want_stats=True)
    keys = ["paths", "rays", "node_visits", "sphere_tests", "light_tests", "lambertian", "metal", "dielectric", "absorbed", "missed", "depth_out"]
    tot = torch.tensor([cnt[k] for k in keys], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot)
    total = {k: float(v) for k, v in zip(keys, tot.tolist())}
    local_flops = algorithmic_flops(total) / world           # every rank's share of the frame's events (the library may split the frame by pixels)
    # the same frame with every camera ray walking the BVH (RTW_FLAG_NO_CANDIDATES): the event counts of the plain tree-walk algorithm
    # SURVEY 8(d) describes.  `roofline.achieved` uses the events the timed kernel EXECUTES (above); this count is reported beside it.
    tw = None
    if mode == R.RTW_WAVEFRONT and not args.lane_per_pixel:
        twc = renderer.render_local(R.RenderOptions(seed=SEED, precision=R.RTW_F32, mode=mode, flags=R.RTW_FLAG_COUNT_EVENTS | R.RTW_FLAG_NO_CANDIDATES | base_flags), want_stats=True)
        twt = torch.tensor([twc[k] for k in keys], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(twt)
        twtot = {k: float(v) for k, v in zip(keys, twt.tolist())}
        tw = dict(flop_per_launch=algorithmic_flops(twtot) / world, node_visits=twtot["node_visits"] / world, sphere_tests=twtot["sphere_tests"] / world)

    for _ in range(args.warmup):
        renderer.render()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    stream = torch.cuda.current_stream().cuda_stream
    for k in range(args.steps):
        flush.fill_(k & 0xff)                                # evict L2 between timed iterations (outside the event pairs)
        ev[k][0].record()
        if renderer.comm is not None:
            renderer.render()                                # one C call: this rank's share + the NCCL collective + resolve on rank 0
        else:
            ev[k][2].record()
            renderer.render_local()                          # this rank's share: its samples of every pixel (or its tiles)
            ev[k][3].record()
            renderer.combine()                               # the frame's one collective + resolve on rank 0
        ev[k][1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    last_kernel_ms = renderer.check()                        # a path that would make the reference panic surfaces here
    step_ms = sum(ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps))
    if renderer.comm is not None:
        # the library call fuses kernels + collective: the kernels' own time is the CUDA-event time the library took around this rank's
        # render kernels in the last timed step (rtw_scene_sync)
        kern_ms = float(last_kernel_ms) * args.steps
    else:
        kern_ms = sum(ev[k][2].elapsed_time(ev[k][3]) for k in range(args.steps))
    t = torch.tensor([step_ms, kern_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, kern_ms = t.tolist()

    # ---- end to end through the public host API: scene upload + render + read-back, every step --------------
    pinned = torch.empty((HEIGHT, WIDTH, 3), dtype=torch.uint8, pin_memory=True) if rank == 0 else None   # the result's host buffer
    arr = R.scenes.simple_arrays(SEED)          # the same scene as plain host arrays: what a step uploads

    def make_scene():
        return R.Scene.from_arrays(arr["spheres"], arr["sphere_materials"], arr["planes"], arr["plane_materials"], arr["lights"])

    def e2e_step():
        if world == 1:
            sc = make_scene()                                            # H2D: BVH build + upload
            _, rgb8, _ = sc.render(cam, opts, want_sum=False, out_rgb8=pinned.numpy())       # D2H: resolved image into pinned memory
            sc.close()
            return rgb8
        sc = make_scene()
        rr = D.DistributedRenderer(sc, cam, opts, rank, world, want_sum=False, want_rgb8=True)
        rr.comm = renderer.comm                                          # the communicator outlives scenes
        rr.render()
        if rank == 0:
            pinned.copy_(rr.rgb8, non_blocking=True)
        torch.cuda.synchronize()
        sc.close()
        return pinned
    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())

    # the other renderer on the same frame, outside the timed region (same image bit for bit)
    other = None
    if not args.lane_per_pixel:
        omode = R.RTW_MEGAKERNEL if mode == R.RTW_WAVEFRONT else R.RTW_WAVEFRONT
        oopts = R.RenderOptions(seed=SEED, precision=R.RTW_F32, mode=omode)
        renderer.render_local(oopts)
        barrier()
        oev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        oev[0].record()
        for _ in range(2):
            renderer.render_local(oopts)
        oev[1].record()
        barrier()
        ot = torch.tensor([oev[0].elapsed_time(oev[1]) / 2], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ot, op=dist.ReduceOp.MAX)
        other = dict(mode="megakernel (pooled path stream)" if omode == R.RTW_MEGAKERNEL else "wavefront", kernel_ms_per_step=float(ot.item()),
                     mrays_per_s=total["rays"] / float(ot.item()) * 1e-3)

    # ---- the exact (f64) path on the same frame at reduced spp: the only path with bit-exact parity has a number too ----
    f64 = None
    if world == 1 and not args.no_f64:
        f64_spp = 4
        fcam = (cb.with_vfov(40.).with_aspect_ratio(WIDTH / HEIGHT).with_max_depth(DEPTH).with_image_width(WIDTH).with_image_height(HEIGHT)
                .with_samples_per_pixel(f64_spp).build())
        fopts = R.RenderOptions(seed=SEED, precision=R.RTW_F64, flags=R.RTW_FLAG_COUNT_EVENTS)
        _, _, fst = scene.render(fcam, fopts, want_sum=False, want_rgb8=True)
        best = None
        for _ in range(2):
            _, _, st2 = scene.render(fcam, R.RenderOptions(seed=SEED, precision=R.RTW_F64), want_sum=False, want_rgb8=True)
            best = st2 if best is None or st2["kernel_ms"] < best["kernel_ms"] else best
        fp64_peak = 148 * 64 * 2 * float(measured_peaks()[0].get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
        f64_tflops = algorithmic_flops(fst) / (best["kernel_ms"] * 1e-3) / 1e12
        f64 = dict(kernel="render_mega_kernel<double, exact>", sample=f"{WIDTH}x{HEIGHT} at {f64_spp} spp (of {SPP}), depth {DEPTH}: bit-identical to the oracle",
                   ms_per_step=best["kernel_ms"], mrays_per_s=best["rays"] / best["kernel_ms"] * 1e-3, mpaths_per_s=best["paths"] / best["kernel_ms"] * 1e-3,
                   rays_per_path=best["rays"] / best["paths"], achieved_tflops=f64_tflops, fp64_peak_tflops=fp64_peak,
                   frac_of_fp64_peak=f64_tflops / fp64_peak,
                   peak_source="148 SM x 64 FP64 lanes x 2 x sm_max_mhz (derived; MEASURED_PEAKS.json has no FP64 figure)")
    # ---- BASELINE C3 (3840x2160, 1024 spp: the config BASELINE.json names for 1/2/4/8 GPUs), one frame outside the timed region ----
    c3 = None
    if not args.no_c3 and args.workload == "C2" and args.spp == 500:
        c3cam = (cb.with_vfov(40.).with_aspect_ratio(3840 / 2160).with_max_depth(DEPTH).with_image_width(3840).with_image_height(2160)
                 .with_samples_per_pixel(1024).build())
        c3r = D.DistributedRenderer(scene, c3cam, opts, rank, world, want_sum=False, want_rgb8=True)
        c3r.comm = renderer.comm
        c3r.render()
        barrier()
        c3ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        c3ev[0].record()
        c3st = c3r.render()
        c3ev[1].record()
        barrier()
        c3t = torch.tensor([c3ev[0].elapsed_time(c3ev[1])], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(c3t, op=dist.ReduceOp.MAX)
        c3_ms = float(c3t.item())
        c3_rays = total["rays"] / (WIDTH * HEIGHT * SPP) * (3840 * 2160 * 1024)        # same scene and estimator: rays per path carry over
        c3 = dict(workload="random-spheres 3840x2160, 1024 spp, depth 50 (BASELINE C3)", n_gpus=world, ms_per_step=c3_ms,
                  mrays_per_s=c3_rays / c3_ms * 1e-3, mpaths_per_s=3840 * 2160 * 1024 / c3_ms * 1e-3, steps=1,
                  note="one frame after one warm-up frame, device-timed, max over ranks; rays = paths x the C2 frame's measured rays per path")
        del c3r

    if rank == 0:
        peaks, peak_src = measured_peaks()
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12                      # TFLOP/s at max clock
        secs = step_ms * 1e-3
        mrays = total["rays"] * args.steps / secs * 1e-6
        achieved = local_flops / (kern_ms / args.steps * 1e-3) / 1e12        # this rank's kernel: flop / its duration
        kernel_name = "render_wavefront_kernel" if args.mode == "wavefront" else ("render_mega_kernel<float>" if args.lane_per_pixel else "render_pool_kernel")
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp) and SPP == 500 and world == 1:           # the capture is of the default single-GPU workload
            with open(tp) as f:
                traffic = json.load(f).get(kernel_name, {}).get("dram_bytes_per_launch")
        line = dict(
            metric="Mrays/s", value=mrays, unit="Mrays/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=step_ms / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic",
            config=dict(workload=WORKLOAD, mode=("wavefront (warp-private queues in shared memory)" if args.mode == "wavefront" else
                              "megakernel (lane per pixel, diagnostic)" if args.lane_per_pixel else "megakernel (pooled path stream)"), parallelism=((f"pixels (chunks of the per-frame ordered work queue) dealt to {world} GPU(s), all samples each, into whole-image fixed-point accumulators + 1 NCCL reduce" if (renderer.partition == "samples" and renderer.comm is not None and args.mode == "wavefront" and not args.lane_per_pixel and os.environ.get("RTW_MULTI_PARTITION") != "samples")
                                      else f"samples of every pixel split over {world} GPU(s) (fixed-point accumulators) + 1 NCCL reduce" if renderer.partition == "samples"
                                      else f"tiles16x16 interleaved over {world} GPU(s) + 1 NCCL gather") +
                                     ("; collective inside the C library (rtw_render_rank_device)" if renderer.comm is not None else ("; collective through torch.distributed" if world > 1 else ""))),
                        tmin="RTW_TMIN_REFERENCE: machine epsilon of the working precision (the reference uses f64::EPSILON in f64)", l2="256 MiB fill between timed steps (scene is 40 KB, shared-memory resident)"),
            mpaths_per_s=total["paths"] * args.steps / secs * 1e-6, rays_per_path=total["rays"] / total["paths"],
            kernel_ms_per_step=kern_ms / args.steps,
            roofline=dict(bound="fp32", achieved=achieved, peak=fp32_peak, unit="TFLOP/s", frac=achieved / fp32_peak, traffic=traffic,
                          peak_source=f"148 SM x 128 lanes x 2 x sm_max_mhz ({peak_src} MEASURED_PEAKS.json); no FP32 figure is in that file",
                          kernel=kernel_name, flop_per_launch=local_flops,
                          frac_at_observed_clock=(achieved / (fp32_peak * clocks["sm_mhz"] / sm_max)) if clocks and clocks.get("sm_mhz") else None,
                          # the HBM view of the same launch, to show why "hbm" is not the bound: measured DRAM bytes / kernel time against the
                          # measured copy bandwidth (the scene lives in shared memory; traffic is the accumulators and the image)
                          hbm=(dict(achieved=traffic / (kern_ms / args.steps * 1e-3) / 1e9, peak=float(peaks.get("hbm_gbs", 6650.0)), unit="GB/s",
                                    frac=traffic / (kern_ms / args.steps * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6650.0))) if traffic else None)),
            e2e=dict(value=total["rays"] * e2e_steps / e2e_s * 1e-6, unit="Mrays/s", h2d_bytes_per_step=scene.upload_bytes * world,
                     d2h_bytes_per_step=WIDTH * HEIGHT * 3 + 88, steps=e2e_steps, ms_per_step=e2e_s / e2e_steps * 1e3),
            # per step: tiles partition = render (+ fixed-point -> tiles conversion) per rank + untile on rank 0; samples partition =
            # render per rank + resolve on rank 0
            # per step and rank: candidate pre-pass (2 launches) + queue order + queue split + background-chunk kernel (wavefront) + render kernel (+ fixed-point -> tiles conversion on the tile partition); on
            # rank 0 one untile / resolve kernel
            gpu_launches=args.steps * (((1 if (renderer.partition == "samples" or (args.lane_per_pixel and args.mode == "megakernel")) else 2)
                                        + (5 if (args.mode == "wavefront" and not args.lane_per_pixel) else 0)) * world + 1),
            clocks=clocks,
            events_per_step={k: total[k] for k in keys},
            other_renderer=other,
            f64_path=f64,
            c3=c3,
        )
        if tw:
            # the algorithmic work of a launch is the event count of the plain algorithm (every ray walks the BVH) on the same
            # flattened tree and ray set.  The timed kernel answers camera rays from per-pixel candidate lists and rays that re-hit the
            # isolated sphere they leave without a walk — same hits, same image — so it EXECUTES fewer node visits and sphere tests;
            # `achieved` / `frac` are the algorithmic figure, `executed` is what the kernel's own counters say it did.
            rf = line["roofline"]
            rf["executed"] = dict(flop_per_launch=rf["flop_per_launch"], achieved=rf["achieved"], frac=rf["frac"],
                                  node_visits=cnt["node_visits"], sphere_tests=cnt["sphere_tests"],
                                  note="events the timed kernel executes (events_per_step), same flop table")
            alg = tw["flop_per_launch"] / (kern_ms / args.steps * 1e-3) / 1e12
            rf.update(flop_per_launch=tw["flop_per_launch"], achieved=alg, frac=alg / fp32_peak,
                      frac_at_observed_clock=(alg / (fp32_peak * clocks["sm_mhz"] / sm_max)) if clocks and clocks.get("sm_mhz") else None,
                      algorithmic=dict(node_visits=tw["node_visits"], sphere_tests=tw["sphere_tests"],
                                       note="event counts of an untimed RTW_FLAG_NO_CANDIDATES pass of the same frame: every ray walks the tree"))
        if world == 1 and not args.no_cpu_baseline:
            try:
                c = cpu_reference_run(1, 0, sample_spp=4)
                line["cpu_baseline"] = dict(value=c["rays"] / c["seconds"] * 1e-6, unit="Mrays/s", cores=c["cores"], kind="port",
                                            sample=c["sample"], mpaths_per_s=c["paths"] / c["seconds"] * 1e-6)
                f = cpu_reference_run(1, 0, sample_spp=1, faithful=True)
                line["cpu_baseline"]["faithful_bvh_value"] = f["rays"] / f["seconds"] * 1e-6
            except Exception as e:                                         # the oracle is only a reported baseline
                line["cpu_baseline"] = dict(value=None, unit="Mrays/s", cores=0, kind="port", sample=f"unavailable: {e}")
        _emit(line)
    scene.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_FD = None


def _protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout whatever
    NCCL_DEBUG_FILE says), so fd 1 is pointed at stderr for the whole process and the JSON line goes to a private copy of the
    original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _protect_stdout()
    global SPP, WORKLOAD, WIDTH, HEIGHT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    # profiling aids — any non-default value is recorded in config and is NOT the headline workload
    ap.add_argument("--spp", type=int, default=SPP)
    ap.add_argument("--workload", default="C2", choices=["C2", "C3"], help="C3 = BASELINE config 3 (3840x2160, 1024 spp), for the multi-GPU table")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-f64", action="store_true", help="skip the exact-path measurement (N = 1)")
    ap.add_argument("--no-c3", action="store_true", help="skip the extra BASELINE C3 frame")
    ap.add_argument("--lane-per-pixel", action="store_true", help="diagnostic: the pre-pooling kernel")
    ap.add_argument("--mode", default="wavefront", choices=["megakernel", "wavefront"])
    ap.add_argument("--torch-collective", action="store_true", help="N > 1: run the collective through torch.distributed instead of inside the C library")
    args = ap.parse_args()
    if args.workload == "C3":
        WORKLOAD = WORKLOAD.replace(f"{WIDTH}x{HEIGHT}, {SPP} spp", "3840x2160, 1024 spp [NON-DEFAULT workload: BASELINE C3]")
        WIDTH, HEIGHT, SPP = 3840, 2160, 1024
        args.spp = SPP
    if args.spp != SPP:
        SPP = args.spp
        WORKLOAD = WORKLOAD.replace("500 spp", f"{SPP} spp [NON-DEFAULT profiling workload]")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
