#!/usr/bin/env python
"""Attribute an ncu capture to source lines (no GPU needed).

    python tools/ncu_lines.py <report.ncu-rep> <object.o> <kernel name> [top]

Joins the per-instruction counters of `ncu --page source` (SASS order) with the line table of the same kernel from
`nvdisasm -g` (compile with -lineinfo) by instruction index, checks that the opcodes agree, and prints the source
lines that execute the most warp instructions and collect the most stall samples, plus the opcode mix.
The object file must be the build that was profiled.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(obj, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    out, on, cur = [], False, ("?", 0)
    for l in txt.splitlines():
        if l.startswith("//---") and ".text." in l:
            on = (".text." + kernel + " ") in l + " " or l.strip().endswith(".text." + kernel + " " + "-" * 26)
            on = (".text.%s " % kernel) in l
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return out


def ncu_rows(rep):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        out.append({"text": r[ix["Source"]], "n": int(float(r[ix["Instructions Executed"]] or 0)),
                    "smp": int(float(r[ix["# Samples"]] or 0)),
                    "stall": {c[6:]: int(float(r[ix[c]] or 0)) for c in stallcols}})
    return out


def opcode(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t.strip())
    return t.split()[0].split(".")[0] if t else "?"


def main():
    rep, obj, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    sass = sass_lines(obj, kernel)
    prof = ncu_rows(rep)
    if len(sass) != len(prof):
        print("WARNING: %d SASS instructions in the object vs %d in the report -- different builds?" % (len(sass), len(prof)))
    n = min(len(sass), len(prof))
    bad = sum(1 for i in range(n) if opcode(sass[i][1]) != opcode(prof[i]["text"]))
    if bad:
        print("WARNING: %d opcode mismatches" % bad)
    by = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot = sum(p["n"] for p in prof[:n])
    tots = sum(p["smp"] for p in prof[:n])
    for i in range(n):
        k = sass[i][2]
        by[k][0] += prof[i]["n"]
        by[k][1] += prof[i]["smp"]
        by[k][2][opcode(prof[i]["text"])] += prof[i]["n"]
    cache = {}

    def line_text(f, ln):
        if f not in cache:
            try:
                cache[f] = open(f).read().splitlines()
            except OSError:
                cache[f] = []
        L = cache[f]
        return L[ln - 1].strip()[:100] if 0 < ln <= len(L) else ""
    print("kernel %s: %d warp instructions, %d samples" % (kernel, tot, tots))
    print("%6s %6s  %-28s %s" % ("inst%", "smp%", "file:line", "source"))
    for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%6.2f %6.2f  %-28s %s" % (100.0 * v[0] / max(tot, 1), 100.0 * v[1] / max(tots, 1),
                                        "%s:%d" % (os.path.basename(k[0]), k[1]), line_text(k[0], k[1])))
    ops = collections.Counter()
    for p in prof[:n]:
        ops[opcode(p["text"])] += p["n"]
    print("opcode mix:", ", ".join("%s %.1f%%" % (o, 100.0 * c / max(tot, 1)) for o, c in ops.most_common(16)))


if __name__ == "__main__":
    main()
