"""Per-object SASS opcode histograms of the built library (run after `python -m arseg_b200.build`; needs no GPU).

    python tools/sass_hist.py > profiles/<tag>_sass_histograms.txt

For every object of arseg_b200/_build: instruction count, the 18 most frequent opcodes, and the counts of the mnemonics that
prove which hardware path a kernel uses (tcgen05: UTCHMMA / UTCQMMA / UTCCP, tensor memory: LDTM / STTM, TMA: UTMALDG / UBLKCP,
mbarrier: SYNCS, legacy tensor cores: HMMA, mixed-precision FMA: FHFMA, packed fp32: FFMA2)."""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROOF = ["UTCHMMA", "UTCQMMA", "UTCCP", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA", "FHFMA", "FFMA2", "FFMA", "MUFU", "LDSM", "LDGSTS", "DFMA"]
for obj in sorted(glob.glob(os.path.join(ROOT, "arseg_b200", "_build", "*.o"))):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    ops = collections.Counter()
    funcs = 0
    for l in txt.splitlines():
        if l.lstrip().startswith("Function :"):
            funcs += 1
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            ops[m.group(1)] += 1
    tot = sum(ops.values())
    print("== %s: %d kernels, %d SASS instructions" % (os.path.basename(obj), funcs, tot))
    print("   top: " + ", ".join("%s %d" % kv for kv in ops.most_common(18)))
    print("   path: " + ", ".join("%s %d" % (k, sum(v for o, v in ops.items() if o.startswith(k) and (k != "FFMA" or o == "FFMA"))) for k in PROOF))
