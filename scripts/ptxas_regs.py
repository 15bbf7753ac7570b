"""Compact view of `nvcc -Xptxas -v`: kernel, registers, spill bytes, shared memory.
    python scripts/ptxas_regs.py advchain_b200/csrc/advk_chain.cu [filter-substring]"""
import re
import subprocess
import sys

src = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
       "-Xptxas", "-v", "-c", src, "-o", "/dev/null"]
err = subprocess.run(cmd, capture_output=True, text=True).stderr
name = None
rows = []
for line in err.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        spill = None
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and name:
        spill = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
        continue
    m = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", line)
    if m and name:
        rows.append((name, int(m.group(1)), spill, m.group(2) or "0"))
        name = None
if not rows:
    sys.stderr.write(err)
for n, r, sp, sm in rows:
    if flt in n:
        print("%-90s regs %3d  stack/spill %s  smem %s" % (n[-90:], r, sp, sm))
