"""Registers / spills / shared memory of every kernel of libsaena_b200.so, from `nvcc -Xptxas -v` (no GPU needed).
usage: python tools/ptxas_report.py [> profiles/r02_ptxas.md]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from saena_b200 import build as b  # noqa: E402

rows = []
for src in b.SOURCES:
    cmd = ["nvcc", "-std=c++17", "-O3", "-lineinfo", *b.ARCH, "-Xptxas", "-v", "-I", os.path.join(ROOT, "include"), "-I", b.CSRC,
           *b._nccl_include(), "-c", os.path.join(b.CSRC, src), "-o", "/dev/null"]
    out = subprocess.run(cmd, capture_output=True, text=True).stderr
    name = None
    for line in out.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            spill = None
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            spill = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
            continue
        m = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", line)
        if m and name:
            smem = re.search(r"(\d+) bytes smem", line)
            rows.append((src, re.sub(r"\(.*", "", name), int(m.group(1)), spill, int(smem.group(1)) if smem else 0))
            name = None
print("| file | kernel | registers | stack / spill stores / spill loads (B) | static smem (B) |\n|---|---|---:|---|---:|")
for src, k, r, sp, sm in rows:
    k = k.replace("void ", "")
    print(f"| {src} | `{k}` | {r} | {sp[0]} / {sp[1]} / {sp[2]} | {sm} |")
