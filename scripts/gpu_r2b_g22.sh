#!/bin/bash
# why does the 4K frame take 1368 ms through rtw_render and 1252 ms through bench.py?
mkdir -p gpurun_out
echo default; python scripts/c3_once.py 3
echo no-bg; RTW_CHEAP_TAIL_PATHS=-1 python scripts/c3_once.py 2
echo no-order; RTW_NO_CHUNK_ORDER=1 python scripts/c3_once.py 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2b_launches_c3.csv python scripts/c3_once.py 2 > gpurun_out/r2b_launches_c3.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2b_launches_c3.csv')))
for r in rows[-10:]:
    if len(r)>10: print('  ',r[4][:60], r[-1])
PY
