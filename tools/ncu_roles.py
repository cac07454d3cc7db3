"""Per-role instruction / stall breakdown of a warp-specialised kernel from an ncu source page.

    python tools/ncu_roles.py <rep> <lib.so> <kernel-substring> <file.cu> name:lo-hi ...  [wait:lo-hi]

Every SASS instruction is attributed to the role whose source-line range (in <file.cu>) was seen last in address order, so
inlined header code (cuda_fp16.hpp, ...) and PTX wrappers count for the role that uses them; the range named `wait` is the
mbarrier wait loop and is reported separately per role."""
import csv, io, subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_lines import sass_lines
rep, lib, kernel, fname = sys.argv[1:5]
ranges, wait = [], None
for a in sys.argv[5:]:
    n, r = a.split(":"); lo, hi = r.split("-")
    if n == "wait": wait = (int(lo), int(hi))
    else: ranges.append((n, int(lo), int(hi)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi_ = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi_]; ix = {h: j for j, h in enumerate(hdr)}
inst = [r for r in rows[hi_ + 1:] if len(r) == len(hdr)]
sl = sass_lines(lib, kernel)
agg, cur = {}, "prologue"
for k, r in enumerate(inst):
    line = sl[k][1] if k < len(sl) else None
    key = cur
    if line and line[0] == fname:
        hit = [n for n, lo, hi in ranges if lo <= line[1] <= hi]
        if hit: cur = hit[0]; key = cur
        elif wait and wait[0] <= line[1] <= wait[1]: key = cur + ".wait"
    a = agg.setdefault(key, {})
    for h in hdr:
        try: v = float(r[ix[h]])
        except Exception: continue
        a[h] = a.get(h, 0.0) + v
tot = sum(a.get("Instructions Executed", 0) for a in agg.values())
for role, a in sorted(agg.items()):
    print("== %-14s instr %6.1f%% (%.0fM)  samples %.0f" % (role, 100 * a.get("Instructions Executed", 0) / tot, a.get("Instructions Executed", 0) / 1e6, a.get("# Samples", 0)))
    st = sorted(((v, h) for h, v in a.items() if h.startswith("stall") and "Not Issued" not in h), reverse=True)[:6]
    print("      ", ", ".join("%s %.0f" % (h[6:], v) for v, h in st))
