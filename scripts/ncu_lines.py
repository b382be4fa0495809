#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction samples by CUDA source line.

  python scripts/ncu_lines.py gpurun_out/prof.ncu-rep <kernel-mangled-substring> [lib.so] [--top N]

ncu's CSV source page is SASS-only; the line table comes from `nvdisasm -g` on the cubin inside the
shared library (compiled with -lineinfo).  Instruction i of the report == instruction i of the function.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    for i, r in enumerate(rows):
        if "# Samples" in r:
            hdr = r
            body = rows[i + 1:]
            break
    else:
        raise SystemExit("no source page in report")
    return hdr, [r for r in body if len(r) == len(hdr)]


def line_table(lib, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    for f in sorted(os.listdir(tmp)):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        m = re.search(r"^\.text\.(\S*%s\S*):\n" % re.escape(kernel_sub), txt, re.M)
        if not m:
            continue
        body = txt[m.end():]
        end = re.search(r"^//-+ \.", body, re.M)
        if end:
            body = body[:end.start()]
        table, cur = [], ("?", 0)
        for ln in body.splitlines():
            mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if mm:
                cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                table.append((cur, ln.split("*/", 1)[1].strip()))
        return table
    raise SystemExit(f"kernel {kernel_sub} not found in {lib}")


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    top = 40
    if "--top" in sys.argv:
        top = int(sys.argv[sys.argv.index("--top") + 1]); args = [a for a in args if a != str(top)]
    rep, ksub = args[0], args[1]
    lib = args[2] if len(args) > 2 else "ray_tracing_weekend_b200/lib/librtw_cuda.so"
    hdr, rows = sass_rows(rep)
    table = line_table(lib, ksub)
    if len(table) != len(rows):
        print(f"warning: {len(rows)} SASS rows in the report vs {len(table)} in the cubin; matching by index", file=sys.stderr)
    si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
    for k, r in enumerate(rows):
        key = table[k][0] if k < len(table) else ("?", 0)
        a = agg[key]
        a[0] += int(r[si]); a[1] += int(r[ii]); a[2] += int(r[ti])
        for s in stalls:
            a[3][hdr[s]] += int(r[s] or 0)
    ts = sum(a[0] for a in agg.values()) or 1
    tinst = sum(a[1] for a in agg.values()) or 1
    tthr = sum(a[2] for a in agg.values())
    print(f"total samples {ts}, warp instructions {tinst}, thread instructions {tthr}, avg threads/inst {tthr / tinst:.2f}")
    # per-function roll-up: a source line belongs to the last function header above it
    funcs = {}
    def func_of(fn, ln):
        for base in ("ray_tracing_weekend_b200/csrc",):
            path = os.path.join(base, fn)
            if os.path.exists(path):
                if path not in funcs:
                    heads = []
                    for i, t in enumerate(open(path, errors="replace").read().splitlines(), 1):
                        m = re.match(r"^(?:template.*>\s*)?(?:RTW_D|RTW_HD|__global__|static RTW_HD|inline)\b.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", t.strip())
                        if m and not t.strip().startswith("//"):
                            heads.append((i, m.group(1)))
                    funcs[path] = heads
                name = "?"
                for i, nme in funcs[path]:
                    if i <= ln:
                        name = nme
                    else:
                        break
                return f"{fn}:{name}"
        return fn
    roll = defaultdict(lambda: [0, 0, 0])
    for key, a in agg.items():
        f = func_of(*key)
        roll[f][0] += a[0]; roll[f][1] += a[1]; roll[f][2] += a[2]
    print("-- by function (samples%, warp-inst%, threads/inst) --")
    for f, a in sorted(roll.items(), key=lambda kv: -kv[1][0])[:22]:
        print(f"{a[0] / ts * 100:8.2f} {a[1] / tinst * 100:6.2f} {a[2] / max(a[1], 1):8.1f}  {f}")
    print("-- by source line --")
    print(f"{'samples%':>8} {'inst%':>6} {'thr/inst':>8}  top stalls | source line")
    src_cache = {}
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        fn, ln = key
        text = ""
        for base in ("ray_tracing_weekend_b200/csrc", "/usr/local/cuda/include", "/usr/local/cuda/include/crt"):
            p = os.path.join(base, fn)
            if os.path.exists(p):
                if p not in src_cache:
                    src_cache[p] = open(p, errors="replace").read().splitlines()
                if 0 < ln <= len(src_cache[p]):
                    text = src_cache[p][ln - 1].strip()[:90]
                break
        st = sorted(a[3].items(), key=lambda kv: -kv[1])[:2]
        sts = ",".join(f"{n[6:]}:{v * 100 // max(a[0], 1)}" for n, v in st)
        print(f"{a[0] / ts * 100:8.2f} {a[1] / tinst * 100:6.2f} {a[2] / max(a[1], 1):8.1f}  {sts:28s} | {fn}:{ln}  {text}")


if __name__ == "__main__":
    main()
