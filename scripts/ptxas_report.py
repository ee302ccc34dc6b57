#!/usr/bin/env python
"""Registers / spills / shared memory of every kernel, from `nvcc -Xptxas -v` (run HERE, no GPU needed).
usage: python scripts/ptxas_report.py [extra nvcc flags...]"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--fmad=false",
       "-std=c++17", "-Xptxas", "-v", "-I", str(ROOT / "include"), "-I", str(ROOT / "gradabm-june_b200/csrc"), "-c",
       "-o", "/tmp/gj_ptxas.o", str(ROOT / "gradabm-june_b200/csrc/gj_kernels.cu")] + sys.argv[1:]
out = subprocess.run(cmd, capture_output=True, text=True).stderr
name = None
rows = []
for line in out.splitlines():
    m = re.search(r"Compiling entry function '([^']+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void gj::", "")
        cur = {"name": name}
        rows.append(cur)
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and rows:
        rows[-1].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
    m = re.search(r"Used (\d+) registers", line)
    if m and rows:
        rows[-1]["regs"] = int(m.group(1))
        s = re.search(r"(\d+) bytes smem", line)
        rows[-1]["smem"] = int(s.group(1)) if s else 0
for r in sorted(rows, key=lambda r: r["name"]):
    print(f"{r['name']:<50} regs {r.get('regs', 0):>3}  spill st/ld {r.get('spill_st', 0):>4}/{r.get('spill_ld', 0):<4} "
          f"stack {r.get('stack', 0):>4}  smem {r.get('smem', 0)}")
