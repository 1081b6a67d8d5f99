#!/usr/bin/env python3
"""Print the SASS of one out-of-line device function of suews_run_kernel<false>.  usage: tools/sass_of.py lib.so symbol-substring"""
import subprocess, sys, re
lib, want = sys.argv[1], sys.argv[2]
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
sec = None
for ln in elf.splitlines():
    t = ln.split()
    if len(t) >= 9 and t[-1] == ".text._Z16suews_run_kernelILb0ELi1024ELi1EEv9SuRunArgs" and "PROGBITS" in ln:
        sec = "0x" + t[0]
rng = None
for ln in elf.splitlines():
    t = ln.split()
    if len(t) >= 7 and t[5] == sec and t[3] in ("0x2", "0x22", "0x12") and want in t[6]:
        rng = (int(t[1], 16), int(t[2], 16), t[6])
        break
print(rng)
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
on = False
for ln in sass:
    if "Function : _Z16suews_run_kernelILb0" in ln:
        on = True
        continue
    if on and "Function :" in ln:
        break
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
        if m:
            a = int(m.group(1), 16)
            if rng[0] <= a < rng[0] + rng[1]:
                print("%06x  %s" % (a - rng[0], m.group(2).strip()))
