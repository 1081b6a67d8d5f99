#!/usr/bin/env python3
"""Source lines ranked by local-memory sectors (stack traffic).  usage: tools/ncu_local_by_line.py rep cell_steps [top]"""
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
    rows.append((fpath, int(rec[0]), rec[1].strip()[:84], g("L2 Theoretical Sectors Local"), g("# Samples"), g("stall_long_sb")))
tot = sum(r[3] for r in rows); ts = sum(r[4] for r in rows)
print("local sectors %.3g = %.0f B per cell-step (L1-level)" % (tot, tot * 32 / cs))
byfile = {}
for r in rows: byfile[r[0]] = byfile.get(r[0], 0) + r[3]
print("  ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(byfile.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(rows, key=lambda r: -r[3])[:top]:
    print("%-16s %4d %5.2f%% (samples %5.2f%%)  %s" % (r[0][:16], r[1], 100 * r[3] / tot, 100 * r[4] / ts, r[2]))
