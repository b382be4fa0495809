#!/bin/bash
# turn the round-2 capture of the headline kernel into the text files profiles/ keeps (run here, not on the GPU box)
REP=${1:-gpurun_out/r2_prof_wavefront_bench.ncu-rep}
K=render_wavefront_kernelILb0ELi640ELi112ELb1ENS_9SceneViewIfEELb0ELi0E
ncu -i $REP --page details > profiles/r2_ncu_details_render_wavefront_kernel.txt
(echo "# ncu --set full of render_wavefront_kernel inside \`python bench.py\` (BASELINE C2, 500 spp) at the round's last commit; per-line roll-up by scripts/ncu_lines.py"; python scripts/ncu_lines.py $REP $K --top 60 | cut -c1-220) > profiles/r2_ncu_lines_render_wavefront_kernel.txt
ncu -i $REP --page raw --csv | python3 -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
g=lambda k: float(r[h.index(k)])
rd,wr=g('dram__bytes_read.sum'),g('dram__bytes_write.sum')
unit_r=rows[1][h.index('dram__bytes_read.sum')]; unit_w=rows[1][h.index('dram__bytes_write.sum')]
mul={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}
print(json.dumps(dict(dram_bytes_read=rd*mul[unit_r], dram_bytes_write=wr*mul[unit_w], dram_bytes_per_launch=rd*mul[unit_r]+wr*mul[unit_w],
  duration_ms=g('gpu__time_duration.sum')*({'ms':1,'us':1e-3,'ns':1e-6,'s':1e3}[rows[1][h.index('gpu__time_duration.sum')]]),
  issue_active_pct=g('smsp__issue_active.avg.pct_of_peak_sustained_active'), threads_per_inst=g('smsp__thread_inst_executed_per_inst_executed.ratio'),
  icc_hit_pct=g('sm__icc_request_hit_rate.pct'), fma_pipe_pct=g('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'), alu_pipe_pct=g('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
  no_instruction_per_issue=g('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio'))))
"
