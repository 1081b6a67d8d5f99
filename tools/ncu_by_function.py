#!/usr/bin/env python3
"""Aggregate an ncu report's per-instruction samples by device function.
usage: tools/ncu_by_function.py report.ncu-rep library.so [kernel-substring]
Function address ranges come from the ELF symbol table of the kernel's .text section."""
import csv, subprocess, sys, io, re
rep, lib = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "suews_run_kernelILb0ELi1024ELi1"
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
secidx = None
for ln in elf.splitlines():
    t = ln.split()
    if len(t) >= 9 and t[-1] == ".text._Z16" + kern + "EEv9SuRunArgs" and "PROGBITS" in ln:
        secidx = "0x" + t[0]
funcs = []
for ln in elf.splitlines():
    t = ln.split()
    if len(t) >= 7 and t[5] == secidx and t[3] in ("0x2", "0x22", "0x12"):
        try:
            funcs.append((int(t[1], 16), int(t[2], 16), re.sub(r"_ZN\d+_INTERNAL_[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]{8}", "_Z", re.sub(r"^\$_Z16\w+?SuRunArgs\$", "", t[6]))))
        except ValueError:
            pass
funcs.sort()
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
rows = list(rd)
base = int(rows[0]["Address"], 16)
agg = {}
first = {}
for r in rows:
    off = int(r["Address"], 16) - base
    name = "kernel_body"
    for v, s, n in funcs:
        if v <= off < v + s and s < 400000:
            name = n
    if name not in first:
        first[name] = int(r["Instructions Executed"] or 0)   # executions of the entry instruction = warp-level calls
    a = agg.setdefault(name, [0, 0, 0, 0, 0])
    a[0] += int(r["# Samples"] or 0)
    a[1] += int(r["Instructions Executed"] or 0)
    a[2] += int(r["Thread Instructions Executed"] or 0)
    a[3] += int(r.get("stall_wait") or 0)
    a[4] += int(r.get("stall_long_sb") or 0)
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
wsteps = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0   # warp-steps of the launch (cells/32 * timesteps)
print("%-46s %7s %7s %6s %6s %6s %9s %9s" % ("function", "samp%", "inst%", "simt", "wait%", "lsb%", "calls/stp", "inst/call"))
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print("%-46s %6.2f%% %6.2f%% %6.1f %5.1f%% %5.1f%% %9.2f %9.1f" % (n[:46], 100.0 * a[0] / ts, 100.0 * a[1] / ti, a[2] / max(a[1], 1),
                                                   100.0 * a[3] / max(a[0], 1), 100.0 * a[4] / max(a[0], 1),
                                                   first[n] / wsteps if wsteps else 0.0, a[1] / max(first[n], 1)))
