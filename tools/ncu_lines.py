"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line.
usage: ncu_lines.py file.csv [kernel_index] [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
# split into kernels at "Function Name" rows; a kernel has several file sections
kernels, cur = [], None
fpath = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Function Name":
        if cur is None or cur["name"] != r[1] or fpath in cur["files"]:
            cur = {"name": r[1], "files": set(), "lines": []}; kernels.append(cur)
        cur["files"].add(fpath); continue
    if r[0] == "Line No": hdr = r; continue
    if cur is not None and len(r) > 8 and r[2] == "-":
        cur["lines"].append((fpath, r))
k = kernels[kidx]
ie = hdr.index("Instructions Executed"); te = hdr.index("Thread Instructions Executed"); sm = hdr.index("# Samples")
tot = sum(int(r[ie]) for _, r in k["lines"]); tots = sum(int(r[sm]) for _, r in k["lines"])
print(k["name"][:60], "total warp instr", tot, "samples", tots)
lines = sorted(k["lines"], key=lambda fr: -int(fr[1][ie]))
for f, r in lines[:topn]:
    print("%5.1f%% inst %5.1f%% smp  thr/inst %4.1f  %s:%s  %s" % (100.0 * int(r[ie]) / tot, 100.0 * int(r[sm]) / max(tots, 1),
          int(r[te]) / max(int(r[ie]), 1), f.split("/")[-1][:18], r[0], r[1].strip()[:90]))

# ---- per-section summary: a section starts at a source line containing "// ----" (same file); other files by name
import re
sec_of = {}
src_cache = {}
def sections(path):
    if path in src_cache: return src_cache[path]
    marks = []
    try:
        for n, line in enumerate(open(path), 1):
            if "// ----" in line: marks.append((n, line.strip()[:60]))
    except OSError:
        pass
    src_cache[path] = marks
    return marks
agg = collections.Counter(); sagg = collections.Counter()
for f, r in k["lines"]:
    ln = int(r[0]); name = f.split("/")[-1]
    lab = name
    for n, text in sections(f):
        if n <= ln: lab = name + ": " + text
    if name.endswith(".hpp") or name.endswith(".h"): lab = name + ":" + r[0]
    agg[lab] += int(r[ie]); sagg[lab] += int(r[sm])
print("---- sections")
for lab, v in agg.most_common(14):
    print("%5.1f%% inst %5.1f%% smp  %s" % (100.0 * v / tot, 100.0 * sagg[lab] / max(tots, 1), lab))
