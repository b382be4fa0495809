set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python - <<'PY'
import time, numpy as np
import ray_tracing_weekend_b200 as R
w,l,cb=R.scenes.simple(20261018)
sc=R.Scene(w,l)
print(sc.info())
for (W,H,spp,prec) in [(1920,1080,16,R.RTW_F32),(1920,1080,100,R.RTW_F32),(1920,1080,4,R.RTW_F64)]:
    cam=cb.with_vfov(40.).with_aspect_ratio(W/H).with_max_depth(50).with_image_width(W).with_image_height(H).with_samples_per_pixel(spp).build()
    for it in range(2):
        t0=time.time(); s,q,st=sc.render(cam,R.RenderOptions(precision=prec),want_sum=False); t1=time.time()
        print(W,H,spp,prec,"kernel_ms",round(st['kernel_ms'],2),"total_ms",round(st['total_ms'],2),"wall",round(t1-t0,3),"Mpaths/s",round(st['paths']/st['kernel_ms']*1e-3,1),"Mrays/s",round(st['rays']/st['kernel_ms']*1e-3,1),"rays/path",round(st['rays']/st['paths'],3))
R.write_ppm("gpurun_out/first.ppm", q)
PY
