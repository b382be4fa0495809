#!/bin/bash
# round 2, session 3, call 8: plane loop not unrolled (wavefront kernel 3 144 -> 2 552 instructions).  A/B first; the evidence of the round's last commit is
# refreshed in the same visit only if the new build wins on C2 and does not lose on C5
mkdir -p gpurun_out
O=gpurun_out/r2c_g8.jsonl; : > $O
timeout 200 python scripts/variant_bench.py --config C2 --spp 500 --reps 3 c4:c4 default 2>&1 | tee -a $O | cut -c1-90
timeout 200 python scripts/variant_bench.py --config C5 --spp 256 --reps 2 c4:c4 default 2>&1 | tee -a $O | cut -c1-90
python - <<'PY'
import json,sys
r=[json.loads(l) for l in open('gpurun_out/r2c_g8.jsonl') if l.startswith('{')]
c2=[x for x in r if x['config']=='C2']; c5=[x for x in r if x['config']=='C5']
ok = len(c2)==2 and len(c5)==2 and all(x.get('same_image_as_first') for x in r) and c2[1]['speedup_vs_first']>1.01 and c5[1]['speedup_vs_first']>0.99
print('ADOPT' if ok else 'KEEP', c2[1]['speedup_vs_first'] if len(c2)==2 else None, c5[1]['speedup_vs_first'] if len(c5)==2 else None)
sys.exit(0 if ok else 1)
PY
[ $? -eq 0 ] || exit 0
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_gpu.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/r2_bench_n1.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_wavefront -s 2 -c 1 -o gpurun_out/r2_prof_wavefront_bench -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-f64 --no-c3 > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches_bench_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-f64 --no-c3 > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"
timeout 200 python scripts/stress_configs.py C1 C5 C4 > gpurun_out/r2_stress_configs_b.jsonl 2>/dev/null; echo "stress rc=$?"
