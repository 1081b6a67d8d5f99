#!/usr/bin/env python3
"""Extract the output-variable metadata table varListAll (name, unit, format, aggregation letter,
group, level) from SUEWS-SourceCode/suews_ctrl_output.f95:56-560.  It is what
output_name_n / output_size expose to supy (suews_ctrl_driver.f95:2778-2812) and what defines
the hourly aggregation of the golden text files (A = mean, S = sum, L = last, T = time).

Emits  suews_b200/varlist.json  (metadata only, no code).  Authoring-container tool.
"""
import json, os, re

SRC = "/root/reference/SUEWS-SourceCode/suews_ctrl_output.f95"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FMT = {"fy": "i0004,1X", "ft": "i0004,1X", "fd": "f08.4,1X", "f94": "f09.4,1X", "f104": "f10.4,1X",
       "f106": "f10.6,1X", "f146": "f14.6,1X"}
AGG = {"aT": "T", "aA": "A", "aS": "S", "aL": "L"}


def main():
    txt = open(SRC).read()
    txt = txt[:txt.index("CONTAINS")]
    rows = []
    pat = re.compile(r"varAttr\(\s*'([^']*)'\s*,\s*'([^']*)'\s*,\s*(\w+)\s*,\s*'([^']*)'\s*,\s*(\w+)\s*,\s*'([^']*)'\s*,\s*(\d+)\s*\)")
    for m in pat.finditer(txt):
        name, unit, fmt, longnm, agg, group, level = m.groups()
        rows.append(dict(name=name, unit=unit, fmt=FMT[fmt], long=longnm, aggreg=AGG[agg], group=group,
                         level=int(level)))
    counts = {}
    for r in rows:
        counts[r["group"]] = counts.get(r["group"], 0) + 1
    print(len(rows), counts)
    os.makedirs(os.path.join(ROOT, "suews_b200"), exist_ok=True)
    with open(os.path.join(ROOT, "suews_b200", "varlist.json"), "w") as fh:
        json.dump(rows, fh, indent=0)


if __name__ == "__main__":
    main()
