"""Aggregate an ncu SASS-level source page by CUDA source line.

    python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-mangled-substring> [topN]

ncu's CSV source page is per SASS instruction; nvdisasm -g gives the instruction -> source-line map of the
same cubin.  Instructions are matched by their order inside the function.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, line, inl = None, None, None
        for l in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
            if m:
                cur = m.group(1)
                continue
            if cur is None or kernel not in cur:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                out.append((int(m.group(1), 16), line, m.group(2).strip()))
        if out:
            break
    return out


def main():
    rep, lib, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ix = {h: j for j, h in enumerate(hdr)}
    inst = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    sl = sass_lines(lib, kernel)
    if len(sl) != len(inst):
        sys.stderr.write("warning: %d SASS instructions in the report vs %d in the cubin\n" % (len(inst), len(sl)))
    agg = {}
    tot_s = tot_i = 0.0
    for k, r in enumerate(inst):
        smp = float(r[ix["# Samples"]] or 0)
        ie = float(r[ix["Instructions Executed"]] or 0)
        line = sl[k][1] if k < len(sl) else None
        a = agg.setdefault(line, [0.0, 0.0, 0])
        a[0] += smp; a[1] += ie; a[2] += 1
        tot_s += smp; tot_i += ie
    src = {}
    print("total samples %.0f, warp instructions %.0f, SASS instructions %d" % (tot_s, tot_i, len(inst)))
    for line, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if line:
            path = None
            for root, _, files in os.walk(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))):
                if line[0] in files and "csrc" in root:
                    path = os.path.join(root, line[0]); break
            if path:
                src.setdefault(path, open(path).read().splitlines())
                if line[1] - 1 < len(src[path]):
                    text = src[path][line[1] - 1].strip()[:100]
        print("%5.1f%% samples %5.1f%% instr %4d sass  %s:%s  %s" % (100 * a[0] / max(tot_s, 1), 100 * a[1] / max(tot_i, 1), a[2],
                                                                      line[0] if line else "?", line[1] if line else "?", text))


if __name__ == "__main__":
    main()
