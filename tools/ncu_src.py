"""Aggregate `ncu --page source --print-source cuda,sass --csv` by CUDA source line (exact correlation from -lineinfo).

    python tools/ncu_src.py <report.ncu-rep> [topN]
"""
import csv
import subprocess
import sys
from collections import defaultdict


def num(v):
    try:
        return float(v)
    except (TypeError, ValueError):
        return 0.0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur_file, hdr, cur_line, cur_src = None, None, None, ""
    agg = defaultdict(lambda: defaultdict(float))
    srcs = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None:
            continue
        if r[0] != "":
            cur_line, cur_src = int(r[0]), r[1].strip()
            srcs[(cur_file, cur_line)] = cur_src
            continue
        d = dict(zip(hdr[2:], r[2:]))
        key = (cur_file, cur_line)
        a = agg[key]
        a["samples"] += num(d.get("# Samples"))
        a["inst"] += num(d.get("Instructions Executed"))
        a["sass"] += 1
        a["smem_wave"] += num(d.get("L1 Wavefronts Shared"))
        a["smem_excess"] += num(d.get("L1 Wavefronts Shared Excessive"))
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k:
                a[k] += num(v)
    tot_s = sum(a["samples"] for a in agg.values())
    tot_i = sum(a["inst"] for a in agg.values())
    tot_w = sum(a["smem_wave"] for a in agg.values())
    print("total samples %d, warp instructions %d, smem wavefronts %d" % (tot_s, tot_i, tot_w))
    st = defaultdict(float)
    for a in agg.values():
        for k, v in a.items():
            if k.startswith("stall_"):
                st[k] += v
    print("stalls: " + ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot_s) for k, v in sorted(st.items(), key=lambda x: -x[1])[:10]))
    for key, a in sorted(agg.items(), key=lambda x: -x[1]["samples"])[:top]:
        tops = sorted(((k[6:], v) for k, v in a.items() if k.startswith("stall_")), key=lambda x: -x[1])[:3]
        print("%5.1f%% smp %5.1f%% inst %5.1f%% smemw %4d sass  %s:%d  [%s]  %s" % (
            100 * a["samples"] / tot_s, 100 * a["inst"] / tot_i, 100 * a["smem_wave"] / max(tot_w, 1), a["sass"], key[0], key[1],
            " ".join("%s %.0f%%" % (k, 100 * v / max(a["samples"], 1)) for k, v in tops), srcs.get(key, "")[:90]))


if __name__ == "__main__":
    main()
