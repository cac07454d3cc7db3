"""Timeline of the column-marching CReFF kernel's three roles for one CTA.  Needs a library built with
ARSEG_NVCC_EXTRA=-DARSEG_XTRACE python -m arseg_b200.build --force  (debug only; the product build has no trace)."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from arseg_b200 import _lib as L  # noqa: E402

subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "prof_creff.py"), "--frames", "11", "--iters", "1"]) if False else None
import runpy
sys.argv = ["prof_creff.py", "--frames", "11", "--iters", os.environ.get("TRACE_ITERS", "1")] + os.environ.get("TRACE_ARGS", "").split()
lib = L.load()
buf = (C.c_longlong * 16384)()
lib.arseg_debug_creff_trace.argtypes = [C.c_void_p, C.c_int]
runpy.run_path(os.path.join(ROOT, "tools", "prof_creff.py"), run_name="__main__")
n = lib.arseg_debug_creff_trace(buf, 16384)
names = {0: "G", 1: "D", 2: "C"}
rows = {}
for r in range(3):
    for st in range(256):
        d = {tag: buf[(r * 256 + st) * 8 + tag] for tag in range(8) if buf[(r * 256 + st) * 8 + tag]}
        if d:
            rows[(r, st - 1)] = d
t0 = min(min(d.values()) for d in rows.values() if d)
steps = sorted({s for (_, s) in rows})
print("cycles relative to the CTA's first event; per role: tags -> time")
for s in steps:
    if not (20 <= s <= 28):
        continue
    for r in (0, 1, 2):
        d = rows.get((r, s))
        if d:
            print("step %3d %s  " % (s, names[r]) + "  ".join("t%d=%7d" % (tag, d[tag] - t0) for tag in sorted(d)))
# average durations over steady-state steps
import statistics
def dur(r, a, b):
    xs = [rows[(r, s)][b] - rows[(r, s)][a] for s in steps if 10 <= s <= 70 and (r, s) in rows and a in rows[(r, s)] and b in rows[(r, s)]]
    return statistics.mean(xs) if xs else float("nan")
def period(r, tag):
    xs = [rows[(r, s + 1)][tag] - rows[(r, s)][tag] for s in steps if 10 <= s <= 70 and (r, s) in rows and (r, s + 1) in rows and tag in rows[(r, s)] and tag in rows[(r, s + 1)]]
    return statistics.mean(xs) if xs else float("nan")
print("period (cycles/step): G %.0f  D %.0f  C %.0f" % (period(0, 0), period(1, 0), period(2, 0)))
print("G: wait ddone %.0f | gather %.0f | nbar+arrive %.0f" % (dur(0, 0, 1), dur(0, 1, 2), dur(0, 2, 3)))
print("D: wait gfull %.0f | wait cdone %.0f | KV conv %.0f | wait qlempty %.0f | Q conv %.0f" % (dur(1, 0, 1), dur(1, 1, 2), dur(1, 2, 3), dur(1, 3, 4), dur(1, 4, 5)))
print("C: wait ddone %.0f | QK+resid %.0f | softmax %.0f | PV %.0f | resid-add %.0f | cls+lse %.0f | stores %.0f" % (dur(2, 0, 1), dur(2, 1, 2), dur(2, 2, 3), dur(2, 3, 4), dur(2, 4, 5), dur(2, 5, 6), dur(2, 6, 7)))
