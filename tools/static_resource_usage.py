"""Static resource usage of every kernel in libfsweep.so (cuobjdump -res-usage; no GPU needed): registers, spill stack,
static shared memory — the numbers to look at before spending GPU time on a kernel.
Usage: python tools/static_resource_usage.py [> profiles/<name>.md]"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "flamo_b200", "libfsweep.so")
out = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True, check=True).stdout
rows, name = [], None
for line in out.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        name = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and name:
        rows.append((name, *map(int, m.groups())))
        name = None
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
rows = [(re.sub(r"\(.*", "", n).replace("void ", ""), *r[1:]) for n, r in zip(names, rows)]
rows.sort(key=lambda r: (-r[2], -r[1], r[0]))
print("# Static resource usage of libfsweep.so (sm_100a, `cuobjdump -res-usage`)\n")
print(f"{len(rows)} kernels.  `stack` > 0 means register spills (or local arrays) — the first thing to look at.\n")
print("| kernel | regs | stack B | static smem B |")
print("|---|---:|---:|---:|")
for n, reg, stack, shared, local in rows:
    print(f"| `{n}` | {reg} | {stack} | {shared} |")
