#!/usr/bin/env python3
"""Who calls a given out-of-line device function, and how often per cell-step.
usage: tools/ncu_callers.py rep lib callee-substring warp_steps"""
import csv, subprocess, sys, io, re
rep, lib, callee, wsteps = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
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
target = [f for f in funcs if callee in f[2]][0]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
# cuda,sass view: source-line rows followed by their SASS rows -> attribute each CALL to the preceding source line
base = None
fpath, hdr, cur = "", None, ("?", 0, "")
agg = {}
recs = list(csv.reader(io.StringIO(out)))
for rec in recs:   # first pass: kernel base address = smallest SASS address
    if len(rec) > 3 and rec[2].startswith("0x"):
        a = int(rec[2], 16)
        base = a if base is None or a < base else base
for rec in recs:
    if len(rec) == 2 and rec[0] == "File Path":
        fpath = rec[1].split("/")[-1]; continue
    if rec and rec[0] == "Line No":
        hdr = rec; continue
    if hdr is None or len(rec) < 8: continue
    if rec[0].strip().isdigit():
        cur = (fpath, int(rec[0]), rec[1].strip()[:90]); continue
    if rec[2].startswith("0x") and "CALL" in rec[3]:
        m = re.search(r"CALL\S*\s+(?:R\d+\s+)?(0x[0-9a-f]+)", rec[3])
        if m and int(m.group(1), 16) - base == target[0]:
            try:
                n = int(rec[hdr.index("Instructions Executed")] or 0)
            except ValueError:
                n = 0
            agg[cur] = agg.get(cur, 0) + n
tot = sum(agg.values())
print("calls of %s: %.1f per cell-step" % (target[2][:50], tot / wsteps))
for k, c in sorted(agg.items(), key=lambda kv: -kv[1])[:40]:
    print("%-16s %4d %7.2f  %s" % (k[0][:16], k[1], c / wsteps, k[2]))
