#!/usr/bin/env python
"""Generate suews_b200/csrc/suews_refargs.inc: the positional argument tables of the two reference entry
points the library replaces,

    SUEWS_cal_Main         SUEWS-SourceCode/suews_ctrl_driver.f95:45
    SUEWS_cal_multitsteps  SUEWS-SourceCode/suews_ctrl_driver.f95:2814

read from the reference's own SUBROUTINE statements and INTENT declarations (names, order, type, extent,
intent -- interface facts, no code).  Run in the build container only (needs /root/reference); the
generated table is committed so the GPU box never reads the reference.
"""
import re
import sys
from pathlib import Path

REF = Path("/root/reference/SUEWS-SourceCode/suews_ctrl_driver.f95")
OUT = Path(__file__).resolve().parent.parent / "suews_b200" / "csrc" / "suews_refargs.inc"

CONSTS = {
    "nsurf": 7, "nvegsurf": 3, "ndays": 366,
    "ncolumnsdataoutsuews": 87, "ncolumnsdataoutsnow": 103, "ncolumnsdataoutsol": 31,
    "ncolumnsdataoutestm": 32, "ncolumnsdataoutdailystate": 50, "ncolumnsdataoutrsl": 30 * 4 + 5 + 13,
}


def strip_comment(line):
    i = line.find("!")
    return line if i < 0 else line[:i]


def find_sub(lines, name):
    pat = re.compile(r"^\s*SUBROUTINE\s+%s\s*\(" % name, re.I)
    for i, l in enumerate(lines):
        if pat.match(l):
            j = i
            while not re.match(r"^\s*END\s+SUBROUTINE\s+%s" % name, lines[j], re.I):
                j += 1
            return i, j
    raise SystemExit("subroutine %s not found" % name)


def arg_list(lines, i0):
    text = ""
    i = i0
    while True:
        l = strip_comment(lines[i]).rstrip()
        cont = l.endswith("&")
        text += l.rstrip("&")
        i += 1
        if not cont:
            break
    inner = text[text.index("(") + 1:text.rindex(")")]
    return [a.strip() for a in inner.split(",") if a.strip()], i


def extent(expr, len_sim_ok):
    """Number of elements of DIMENSION(expr); -1 = len_sim-dependent rows (count = len_sim * cols); 0 = assumed shape."""
    parts = [p.strip() for p in split_top(expr)]
    n = 1
    rows_dep = False
    for p in parts:
        if p == ":":
            return 0, False
        q = p.lower()
        if ":" in q:   # lower:upper
            lo, hi = q.split(":")
            v = evalc(hi) - evalc(lo) + 1
        elif q == "len_sim" and len_sim_ok:
            rows_dep = True
            v = 1
        else:
            v = evalc(q)
        n *= v
    return n, rows_dep


def evalc(q):
    q = q.strip().lower()
    for k, v in sorted(CONSTS.items(), key=lambda kv: -len(kv[0])):
        q = re.sub(r"\b%s\b" % k, str(v), q)
    if not re.fullmatch(r"[0-9+\-* ()]+", q):
        raise SystemExit("cannot evaluate extent %r" % q)
    return int(eval(q))


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        if ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


DECL = re.compile(r"^\s*(INTEGER|REAL\s*\(\s*KIND\s*\(\s*1d0\s*\)\s*\))\s*(.*?)::\s*(.*)$", re.I)


def declarations(lines, i0, i1):
    d = {}
    for l in lines[i0:i1]:
        l = strip_comment(l)
        m = DECL.match(l)
        if not m or "intent" not in m.group(2).lower():
            continue
        typ = "int" if m.group(1).upper().startswith("INTEGER") else "double"
        attrs = m.group(2)
        im = re.search(r"intent\s*\(\s*(\w+)\s*\)", attrs, re.I)
        intent = im.group(1).lower()
        dm = re.search(r"dimension\s*\((.*)\)\s*,\s*intent|dimension\s*\((.*)\)\s*$|dimension\s*\((.*?)\)\s*,", attrs, re.I)
        dim = None
        if dm:
            dim = next(g for g in dm.groups() if g is not None)
        for ent in split_top(m.group(3)):
            ent = ent.strip()
            em = re.match(r"(\w+)\s*(\((.*)\))?$", ent)
            nm = em.group(1)
            edim = em.group(3) if em.group(3) else dim
            d[nm.lower()] = (typ, intent, edim)
    return d


def table(lines, name):
    i0, i1 = find_sub(lines, name)
    args, ibody = arg_list(lines, i0)
    decl = declarations(lines, ibody, i1)
    rows = []
    for a in args:
        typ, intent, dim = decl[a.lower()]
        if dim is None:
            n, rows_dep = 1, False
        else:
            n, rows_dep = extent(dim, True)
        rows.append((a, typ, intent, n, rows_dep))
    return rows, i0 + 1


def emit(rows, macro, f, line):
    f.write("/* %s, suews_ctrl_driver.f95:%d -- %d positional arguments */\n" % (macro, line, len(rows)))
    f.write("#define %s(X) \\\n" % macro)
    for k, (a, typ, intent, n, rows_dep) in enumerate(rows):
        f.write("   X(%s, %s, SU_ARG_%s, %d, %d)%s\n" % (a, "SU_ARG_INT" if typ == "int" else "SU_ARG_DBL", intent.upper(), n,
                                                      1 if rows_dep else 0, " \\" if k + 1 < len(rows) else ""))
    f.write("\n")
    # the same list as a C parameter list (every argument by reference, as gfortran passes them) and as
    # the matching pointer list, for the positional entry points
    f.write("#define %s_PARAMS \\\n" % macro[:-5])
    for k, (a, typ, intent, n, rows_dep) in enumerate(rows):
        const = "const " if intent == "in" else ""
        f.write("   %s%s *%s%s\n" % (const, typ, a, ", \\" if k + 1 < len(rows) else ""))
    f.write("\n#define %s_PTRS \\\n" % macro[:-5])
    for k, (a, typ, intent, n, rows_dep) in enumerate(rows):
        f.write("   (void *)%s%s\n" % (a, ", \\" if k + 1 < len(rows) else ""))
    f.write("\n")


def main():
    lines = REF.read_text(errors="replace").splitlines()
    with open(OUT, "w") as f:
        f.write("/* suews_refargs.inc -- GENERATED by tools/gen_refargs.py from the reference's SUBROUTINE statements;\n"
                " * do not edit.  X(name, type, intent, elements, per_row): elements = product of the declared extents\n"
                " * (0 = assumed-shape array the path never reads: AnOHM/ESTM inputs); per_row = 1 when the leading\n"
                " * extent is len_sim, so the element count is len_sim * elements. */\n\n")
        for sub, macro in (("SUEWS_cal_Main", "SUEWS_CAL_MAIN_ARGS"), ("SUEWS_cal_multitsteps", "SUEWS_CAL_MULTITSTEPS_ARGS")):
            rows, line = table(lines, sub)
            emit(rows, macro, f, line)
            print(sub, len(rows), "arguments")
    return 0


if __name__ == "__main__":
    sys.exit(main())
