"""Attribute samples of an ncu source page to code regions given as source-line ranges, and list the
mbarrier wait sites (consecutive SASS with source lines in [lo,hi]) with the line that follows them.

    python tools/ncu_waits.py <report.ncu-rep> <lib.so> <kernel-substr> <file.cu> <wait_lo> <wait_hi>
"""
import csv, io, subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as nl

rep, lib, kern, fname, lo, hi = sys.argv[1:7]
lo, hi = int(lo), int(hi)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[h]; ix = {x: j for j, x in enumerate(hdr)}
inst = [r for r in rows[h + 1:] if len(r) == len(hdr)]
sl = nl.sass_lines(lib, kern)
tot = sum(float(r[ix['# Samples']] or 0) for r in inst)
toti = sum(float(r[ix['Instructions Executed']] or 0) for r in inst)
k = 0
while k < len(inst):
    line = sl[k][1] if k < len(sl) else None
    if line and line[0] == fname and lo <= line[1] <= hi:
        k0 = k; smp = ie = 0.0
        while k < len(inst) and sl[k][1] and sl[k][1][0] == fname and lo <= sl[k][1][1] <= hi:
            smp += float(inst[k][ix['# Samples']] or 0); ie += float(inst[k][ix['Instructions Executed']] or 0); k += 1
        nxt = None
        for kk in range(k, min(k + 30, len(sl))):
            if sl[kk][1] and sl[kk][1][0] == fname:
                nxt = sl[kk][1][1]; break
        if smp / tot > 0.002:
            print("sass#%5d  samples %5.1f%%  instr %5.1f%%  next line %s" % (k0, 100 * smp / tot, 100 * ie / toti, nxt))
    else:
        k += 1
