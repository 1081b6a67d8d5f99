#!/usr/bin/env python3
"""SASS opcode mix of an ncu report (executed warp-instructions per opcode).  usage: tools/ncu_opmix.py report.ncu-rep [warp_steps]"""
import csv, subprocess, sys, io, re, collections
rep = sys.argv[1]
wsteps = float(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
mix = collections.Counter()
tot = 0
for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
    n = int(r["Instructions Executed"] or 0)
    src = r["Source"].strip()
    src = re.sub(r"^@!?U?P\d+\s+", "", src)
    op = src.split()[0] if src else "?"
    op = ".".join(op.split(".")[:2]) if op.startswith(("MUFU", "LDL", "STL", "LDG", "STG", "LDS", "STS", "BAR", "CALL", "RET")) else op.split(".")[0]
    mix[op] += n
    tot += n
print("total warp-instructions %d%s" % (tot, (" = %.0f per cell-step" % (tot / wsteps)) if wsteps else ""))
for op, n in mix.most_common(40):
    print("%-14s %6.2f%% %s" % (op, 100.0 * n / tot, ("%9.1f /step" % (n / wsteps)) if wsteps else ""))
