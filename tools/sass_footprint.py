#!/usr/bin/env python
"""SASS footprint of one MC kernel instance by source function (development tool).

    python tools/sass_footprint.py build/mc_philox_12.o 'mc_sim_kernelI9PhiloxRngLi1ELi1'

Uses nvdisasm --print-line-info (the build has -lineinfo) and attributes each SASS instruction
to the innermost source function whose line range holds it.  Prints bytes per function: what
has to fit the 32 KB L1.5 instruction cache while a block's warps work through one move type.
"""
import re
import subprocess
import sys
import tempfile
from collections import Counter
from pathlib import Path

obj, pat = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", str(Path(obj).resolve())], cwd=tmp, check=True, capture_output=True)
cubin = next(Path(tmp).glob("*.cubin"))
txt = subprocess.run(["nvdisasm", "--print-line-info", str(cubin)], capture_output=True, text=True).stdout.splitlines()

# function line ranges of the sources (crude: a line that starts a definition at any indent)
fn_re = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__|__host__)[^;{]*?\b([A-Za-z_]\w*)\s*\(")
ranges = {}
for src in Path("chromo_b200/csrc").glob("*.cu*"):
    lines = src.read_text().splitlines()
    starts = []
    for i, ln in enumerate(lines, 1):
        m = fn_re.match(ln)
        if m and "return" not in ln.split("(")[0]:
            starts.append((i, m.group(1)))
    ranges[src.name] = starts


def owner(fname, line):
    best = "?"
    for s, name in ranges.get(fname, []):
        if s <= line:
            best = name
        else:
            break
    return best


inside = False
cur = ("?", 0)
by_fn, by_line = Counter(), Counter()
total = 0
for ln in txt:
    if ln.startswith(".text."):
        inside = pat in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (Path(m.group(1)).name, int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        total += 16
        f = owner(*cur) if cur[0] in ranges else cur[0]
        by_fn[f] += 16
        by_line[cur] += 16
print(f"{pat}: {total} bytes of SASS")
for f, b in by_fn.most_common(40):
    print(f"  {b:8d}  {f}")
if len(sys.argv) > 3:
    print("top lines:")
    for (f, l), b in by_line.most_common(int(sys.argv[3])):
        print(f"  {b:6d}  {f}:{l}")
