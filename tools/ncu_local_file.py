#!/usr/bin/env python3
"""Local-memory sectors of every source line of one file, in line order.  usage: tools/ncu_local_file.py rep cell_steps file-substring [min_pct]"""
import csv, subprocess, io, sys
rep, cs, only = sys.argv[1], float(sys.argv[2]), sys.argv[3]
minp = float(sys.argv[4]) if len(sys.argv) > 4 else 0.05
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fpath, hdr, rows = "", None, []
for rec in csv.reader(io.StringIO(out)):
    if len(rec) == 2 and rec[0] == "File Path":
        fpath = rec[1].split("/")[-1]; continue
    if rec and rec[0] == "Line No":
        hdr = rec; continue
    if hdr is None or len(rec) < len(hdr) or not rec[0].strip().isdigit(): continue
    d = dict(zip(hdr[4:], rec[4:]))
    def g(k):
        try: return int(d.get(k) or 0)
        except ValueError: return 0
    rows.append((fpath, int(rec[0]), rec[1].strip()[:100], g("L2 Theoretical Sectors Local"), g("# Samples"), g("Instructions Executed")))
tot = sum(r[3] for r in rows); ts = sum(r[4] for r in rows); ti = sum(r[5] for r in rows)
sel = [r for r in rows if only in r[0]]
print("%s: %.1f%% of local sectors (%.0f B/cell-step of %.0f)" % (only, 100.0 * sum(r[3] for r in sel) / tot, sum(r[3] for r in sel) * 32 / cs, tot * 32 / cs))
for r in sorted(sel, key=lambda r: r[1]):
    if 100.0 * r[3] / tot >= minp:
        print("%4d  loc %5.2f%% (%4.0f B)  samp %5.2f%%  inst %5.2f%%  %s" % (r[1], 100.0 * r[3] / tot, r[3] * 32 / cs, 100.0 * r[4] / ts, 100.0 * r[5] / ti, r[2]))
