#!/usr/bin/env python3
"""Source lines / files ranked by executed instructions.  usage: tools/ncu_inst_by_line.py rep cell_steps [top]"""
import csv, subprocess, io, sys
rep, cs = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
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
    rows.append((fpath, int(rec[0]), rec[1].strip()[:90], g("Instructions Executed"), g("# Samples")))
ti = sum(r[3] for r in rows); ts = sum(r[4] for r in rows)
print("warp instructions %.4g = %.0f per cell-step" % (ti, ti * 32 / cs))
byfile = {}
for r in rows:
    a = byfile.setdefault(r[0], [0, 0]); a[0] += r[3]; a[1] += r[4]
for k, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print("  %-22s inst %5.1f%%  samples %5.1f%%" % (k, 100 * v[0] / ti, 100 * v[1] / ts))
# 20-line buckets per file
bk = {}
for r in rows:
    a = bk.setdefault((r[0], r[1] // 20 * 20), [0, 0]); a[0] += r[3]; a[1] += r[4]
print("-- 20-line buckets")
for k, v in sorted(bk.items(), key=lambda kv: -kv[1][0])[:top]:
    print("  %-20s %4d-%4d inst %5.2f%% (%6.0f/step) samples %5.2f%%" % (k[0], k[1], k[1] + 19, 100 * v[0] / ti, v[0] * 32 / cs, 100 * v[1] / ts))
print("-- lines")
for r in sorted(rows, key=lambda r: -r[3])[:top]:
    print("%-18s %4d inst %5.2f%% (samples %5.2f%%)  %s" % (r[0][:18], r[1], 100 * r[3] / ti, 100 * r[4] / ts, r[2]))
