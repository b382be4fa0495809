"""N GPUs behind the C ABI (rtw_render_multi, rtw_comm_* / rtw_render_rank): the image must equal the single-GPU image bit for
bit whatever the number of GPUs, the partition and the collective.  The 2-GPU cases skip on a single-GPU box; the world = 1
cases exercise the same entry points there."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SEED = 20261018
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cam(cb, w, h, spp, depth=50):
    return cb.with_vfov(40.).with_aspect_ratio(w / h).with_max_depth(depth).with_image_width(w).with_image_height(h).with_samples_per_pixel(spp).build()


def test_render_multi_one_gpu_is_rtw_render(rtw, simple_scene):
    sc = rtw.Scene(simple_scene["world"], simple_scene["lights"])
    cam = _cam(simple_scene["cb"], 96, 54, 8)
    opts = rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32)
    ref_sum, ref8, st = sc.render(cam, opts)
    got_sum, got8, st1 = sc.render_multi(cam, opts, n_gpus=1)
    assert np.array_equal(ref_sum, got_sum, equal_nan=True) and np.array_equal(ref8, got8) and st["rays"] == st1["rays"]
    with pytest.raises(rtw.RtwError):
        sc.render_multi(cam, opts, n_gpus=rtw.device_count() + 1)
    sc.close()


def test_render_rank_world_one(rtw, simple_scene):
    """The rank API with a one-rank communicator (NCCL initialised, no peer): same image as rtw_render, both partitions."""
    comm = rtw.Comm(rtw.Comm.unique_id(), 0, 1)
    sc = rtw.Scene(simple_scene["world"], simple_scene["lights"])
    cam = _cam(simple_scene["cb"], 96, 54, 8)
    for prec in (rtw.RTW_F32, rtw.RTW_F64):
        opts = rtw.RenderOptions(seed=SEED, precision=prec)
        ref_sum, ref8, st = sc.render(cam, opts)
        got_sum, got8, st1 = sc.render_rank(cam, opts, comm, want_sum=True)
        assert np.array_equal(ref_sum, got_sum, equal_nan=True) and np.array_equal(ref8, got8) and st["rays"] == st1["rays"]
    assert sc.sync() > 0.
    sc.close()
    comm.close()


needs2 = pytest.mark.skipif("__import__('ray_tracing_weekend_b200').device_count() < 2", reason="needs two GPUs")


@needs2
@pytest.mark.parametrize("collective", ["peer", "nccl"])
def test_render_multi_two_gpus_bit_identical(rtw, simple_scene, collective):
    coll = rtw.RTW_COLLECTIVE_PEER if collective == "peer" else rtw.RTW_COLLECTIVE_NCCL
    sc = rtw.Scene(simple_scene["world"], simple_scene["lights"])
    cam = _cam(simple_scene["cb"], 200, 113, 33)             # odd sizes: padding tiles, uneven sample shares
    for mode in (rtw.RTW_WAVEFRONT, rtw.RTW_MEGAKERNEL):
        opts = rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32, mode=mode)
        ref_sum, ref8, st = sc.render(cam, opts)
        got_sum, got8, st2 = sc.render_multi(cam, opts, n_gpus=2, collective=coll)
        assert np.array_equal(ref_sum, got_sum, equal_nan=True) and np.array_equal(ref8, got8)
        assert st2["paths"] == st["paths"] and st2["rays"] == st["rays"]
    sc.close()


@needs2
def test_render_multi_two_gpus_f64_and_general(rtw, simple_scene):
    sc = rtw.Scene(simple_scene["world"], simple_scene["lights"])
    cam = _cam(simple_scene["cb"], 80, 45, 4)
    opts = rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64)
    ref_sum, ref8, st = sc.render(cam, opts)
    got_sum, got8, st2 = sc.render_multi(cam, opts, n_gpus=2)           # tile partition: ordered f64 sums per pixel
    assert np.array_equal(ref_sum, got_sum, equal_nan=True) and np.array_equal(ref8, got8) and st2["rays"] == st["rays"]
    sc.close()
    gw, gl, gcb = rtw.scenes.cornell_box()
    gsc = rtw.Scene(gw, gl)
    gcam = gcb.with_vfov(40.).with_aspect_ratio(1.0).with_max_depth(20).with_image_width(64).with_image_height(64).with_samples_per_pixel(16).build()
    gopts = rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F32)
    ref_sum, ref8, st = gsc.render(gcam, gopts)
    got_sum, got8, st2 = gsc.render_multi(gcam, gopts, n_gpus=2)
    assert np.array_equal(ref_sum, got_sum, equal_nan=True) and np.array_equal(ref8, got8) and st2["rays"] == st["rays"]
    gsc.close()


@needs2
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_render_rank_two_processes(rtw, simple_scene, tmp_path, precision):
    """One process per GPU, no torch: the NCCL unique id travels through a file, rank 0 receives the image."""
    id_file, out = str(tmp_path / "nccl_id"), str(tmp_path / "img.npy")
    helper = os.path.join(ROOT, "tests", "helpers", "rank_render.py")
    procs = [subprocess.Popen([sys.executable, helper, str(r), "2", id_file, out, precision], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    logs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    sc = rtw.Scene(simple_scene["world"], simple_scene["lights"])
    cam = _cam(simple_scene["cb"], 160, 90, 33)
    ref_sum, ref8, _ = sc.render(cam, rtw.RenderOptions(seed=SEED, precision=rtw.RTW_F64 if precision == "f64" else rtw.RTW_F32))
    sc.close()
    assert np.array_equal(ref_sum, np.load(out), equal_nan=True) and np.array_equal(ref8, np.load(out + ".rgb8.npy"))


_OWN_CHILD = r'''
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
import ray_tracing_weekend_b200 as R
SEED = 20261018
world, lights, cb = R.scenes.simple(SEED)
sc = R.Scene(world, lights)
cam = cb.with_vfov(40.).with_aspect_ratio(16 / 9).with_max_depth(50).with_image_width(192).with_image_height(108).with_samples_per_pixel(int(sys.argv[3])).build()
s, rgb8, st = sc.render(cam, R.RenderOptions(seed=SEED, precision=R.RTW_F32))
np.savez(sys.argv[2], rgb_sum=s, rays=st["rays"], paths=st["paths"])
sc.close()
'''


@pytest.mark.parametrize("spp", [8, 160, 320])     # 320: a chunk is one pixel and the last chunks of the queue go out in eighths (chunk_split_kernel)
def test_pixel_ownership_covers_the_frame_once(rtw, tmp_path, spp):
    """The multi-GPU split of one frame by OWNED chunks (chunk_order_kernel), checked on ONE GPU: with RTW_DEBUG_OWN=r,n a plain render
    produces what GPU r of n would — the owned pixels at all their samples, zeros elsewhere.  The n partial frames must be disjoint and
    add up to the full frame bit for bit (rgb sums, ray and path counts): every chunk has exactly one owner.  (On a multi-GPU box
    test_render_multi_two_gpus_bit_identical checks the same through rtw_render_multi.)"""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "own_child.py"
    script.write_text(_OWN_CHILD)

    def run(tag, own):
        env = dict(os.environ)
        env.pop("RTW_DEBUG_OWN", None)
        if own:
            env["RTW_DEBUG_OWN"] = own
        out = tmp_path / f"{tag}.npz"
        subprocess.run([sys.executable, str(script), root, str(out), str(spp)], env=env, check=True, timeout=300)
        return np.load(out)
    full = run("full", None)
    n = 3
    parts = [run(f"own{r}", f"{r},{n}") for r in range(n)]
    covered = np.zeros(full["rgb_sum"].shape[:2], dtype=np.int32)
    total = np.zeros_like(full["rgb_sum"])
    for p in parts:
        s = np.nan_to_num(p["rgb_sum"], nan=-1.0)            # a NaN-poisoned pixel is a rendered pixel
        covered += (s != 0).any(axis=2)
        total += np.where(np.isnan(p["rgb_sum"]), 0.0, p["rgb_sum"])
    ref = full["rgb_sum"]
    assert covered.max() <= 1                                # no pixel rendered by two owners
    assert np.array_equal(np.isnan(ref), np.any([np.isnan(p["rgb_sum"]) for p in parts], axis=0))
    assert np.array_equal(np.where(np.isnan(ref), 0.0, ref), total)
    assert sum(int(p["rays"]) for p in parts) == int(full["rays"]) and sum(int(p["paths"]) for p in parts) == int(full["paths"])
