"""One rank of a one-process-per-GPU render through the C ABI (rtw_comm_* / rtw_render_rank), no torch involved.
usage: rank_render.py RANK WORLD ID_FILE OUT_NPY PRECISION   (the NCCL unique id travels through ID_FILE)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
rank, world, id_file, out, precision = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
os.environ["CUDA_VISIBLE_DEVICES"] = str(rank)          # one GPU per process, as a launcher would arrange
import ray_tracing_weekend_b200 as R  # noqa: E402

SEED = 20261018
if rank == 0:
    uid = R.Comm.unique_id()
    with open(id_file + ".tmp", "wb") as f:
        f.write(uid)
    os.replace(id_file + ".tmp", id_file)
else:
    for _ in range(600):
        if os.path.exists(id_file):
            break
        time.sleep(0.1)
    with open(id_file, "rb") as f:
        uid = f.read()
comm = R.Comm(uid, rank, world)
world_h, lights_h, cb = R.scenes.simple(SEED)
cam = cb.with_vfov(40.).with_aspect_ratio(160 / 90).with_max_depth(50).with_image_width(160).with_image_height(90).with_samples_per_pixel(33).build()
sc = R.Scene(world_h, lights_h)
prec = R.RTW_F64 if precision == "f64" else R.RTW_F32
rgb_sum, rgb8, st = sc.render_rank(cam, R.RenderOptions(seed=SEED, precision=prec), comm, want_sum=True, want_rgb8=True)
if rank == 0:
    np.save(out, rgb_sum)
    np.save(out + ".rgb8.npy", rgb8)
print(f"rank {rank}: paths {st['paths']} rays {st['rays']}", flush=True)
sc.close()
comm.close()
