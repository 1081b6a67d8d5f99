#!/usr/bin/env python3
"""Sum one source-page column per out-of-line device function.
usage: tools/ncu_metric_by_function.py rep lib "L2 Theoretical Sectors Local" warp_steps [scale]"""
import csv, subprocess, sys, io, re
rep, lib, col, wsteps = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
scale = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
secidx = None
for ln in elf.splitlines():
    t = ln.split()
    if len(t) >= 9 and t[-1] == ".text._Z16suews_run_kernelILb0ELi1024ELi1EEv9SuRunArgs" and "PROGBITS" in ln:
        secidx = "0x" + t[0]
funcs = []
for ln in elf.splitlines():
    t = ln.split()
    if len(t) >= 7 and t[5] == secidx and t[3] in ("0x2", "0x22", "0x12"):
        try:
            funcs.append((int(t[1], 16), int(t[2], 16), re.sub(r"_ZN\d+_INTERNAL_[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]{8}", "_Z", re.sub(r"^\$_Z16\w+?SuRunArgs\$", "", t[6]))))
        except ValueError:
            pass
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
base = int(rows[0]["Address"], 16)
agg = {}
for r in rows:
    try:
        v = int(r[col] or 0)
    except ValueError:
        continue
    if not v:
        continue
    off = int(r["Address"], 16) - base
    name = "kernel_body"
    for a, s, n in funcs:
        if a <= off < a + s and s < 400000:
            name = n
    agg[name] = agg.get(name, 0) + v
tot = sum(agg.values())
print("%s: %.1f per cell-step (x%g)" % (col, scale * tot / wsteps, scale))
for n, c in sorted(agg.items(), key=lambda kv: -kv[1])[:28]:
    print("%-64s %9.1f %5.1f%%" % (n[:64], scale * c / wsteps, 100.0 * c / tot))
