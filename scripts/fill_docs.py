"""Fill the @@PLACEHOLDER@@ numbers of scripts/templates/*.tmpl from a bench.py JSON line (+ an optional JSON dict of extra values) and write
DESIGN.md, profiles/README.md and README.md.  usage: fill_docs.py bench_line.json [extra.json]"""
import json,sys,re,os
d=json.loads(open(sys.argv[1]).read())
extra=json.loads(open(sys.argv[2]).read()) if len(sys.argv)>2 else {}
rf=d['roofline']
vals={
 'C2_MS': f"{d['ms_per_step']:.1f}", 'C2_GRAYS': f"{d['value']/1e3:.2f}", 'C2_GPATHS': f"{d['mpaths_per_s']/1e3:.2f}",
 'C2_MRAYS': f"{d['value']:,.0f}".replace(',',' '), 'C2_MPATHS': f"{d['mpaths_per_s']:,.0f}".replace(',',' '),
 'MEGA_GRAYS': f"{d['other_renderer']['mrays_per_s']/1e3:.1f}", 'MEGA_MRAYS': f"{d['other_renderer']['mrays_per_s']:,.0f}".replace(',',' '), 'MEGA_MS': f"{d['other_renderer']['kernel_ms_per_step']:.1f}",
 'E2E_GRAYS': f"{d['e2e']['value']/1e3:.2f}", 'E2E_MRAYS': f"{d['e2e']['value']:,.0f}".replace(',',' '), 'E2E_MS': f"{d['e2e']['ms_per_step']:.1f}",
 'F64_MRAYS': f"{d['f64_path']['mrays_per_s']:.0f}", 'CPU_MRAYS': f"{d['cpu_baseline']['value']:.1f}", 'CPU_CORES': str(d['cpu_baseline']['cores']),
 'C2_TFLOPS': f"{rf['achieved']:.1f}", 'C2_FRAC': f"{rf['frac']:.3f}", 'C2_EXEC_FRAC': f"{rf['executed']['frac']:.3f}",
 'C3_MS': f"{d['c3']['ms_per_step']:.0f}", 'C3_GRAYS': f"{d['c3']['mrays_per_s']/1e3:.1f}",
}
vals.update(extra)
# multi-GPU lines, when they exist
for n in (2, 4, 8):
    f=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', f'r2_bench_n{n}.json')
    if os.path.exists(f):
        m=json.loads(open(f).read())
        vals.update({f'N{n}_MRAYS': f"{m['value']:,.0f}".replace(',',' '), f'N{n}_MS': f"{m['ms_per_step']:.2f}", f'N{n}_KMS': f"{m['kernel_ms_per_step']:.2f}",
                     f'N{n}_X': f"{float(extra.get('SCALING_N1_MS', d['ms_per_step']))/m['ms_per_step']:.2f}", f'N{n}_E2E': f"{m['e2e']['value']:,.0f}".replace(',',' '),
                     f'N{n}_C3': f"{m['c3']['ms_per_step']:.0f}" if m.get('c3') else '—'})
    else:
        vals.update({f'N{n}_{k}': 'not measured' for k in ('MRAYS','MS','KMS','X','E2E','C3')})
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for t, p in (('DESIGN.md.tmpl', 'DESIGN.md'), ('profiles_README.md.tmpl', 'profiles/README.md'), ('README.md.tmpl', 'README.md')):
    s=open(os.path.join(ROOT, 'scripts', 'templates', t)).read()
    p=os.path.join(ROOT, p)
    for k,v in vals.items(): s=s.replace('@@'+k+'@@', v)
    open(p,'w').write(s)
    left=set(re.findall(r'@@(\w+)@@', s))
    print(p,'left:',left)
