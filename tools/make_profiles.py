"""Summarise a GPU session's raw artefacts (gpurun_out/) into the tracked profiles/ directory.

    python tools/make_profiles.py <tag>      e.g. r01_d

  profiles/<tag>_bench_*.json          copies of the bench lines
  profiles/<tag>_ncu_launches.txt      per-kernel totals / shares from the ncu launch list (gpurun_out/launches.csv)
  profiles/<tag>_creff_ncu.txt         key metrics of the `ncu --set full` capture of the CReFF kernel + hottest source lines
  profiles/<tag>_creff_sweep.md        BASELINE config 5 table
  profiles/ncu_traffic.json            dram bytes per launch of the dominant kernel (read by bench.py for roofline.traffic)
"""
import csv
import json
import os
import re
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

KEYS = ["gpu__time_duration.sum", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_not_selected_per_warp_active.pct"]


def launches(tag):
    p = os.path.join(OUT, "launches.csv")
    if not os.path.exists(p):
        return
    agg = defaultdict(lambda: [0, 0.0])
    for r in csv.reader(l for l in open(p) if l.startswith('"')):
        if len(r) < 15 or not r[0].isdigit():
            continue
        name = re.sub(r"\(.*", "", r[4])
        agg[name][0] += 1
        agg[name][1] += float(r[-1]) / 1e3
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(PROF, tag + "_ncu_launches.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --alt-precision none\n"
                "# (cold-cache, serialised launches: the kernels' SHARES are what must agree with bench.py's live event timings)\n")
        for name, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%6.2f%% %5d launches %11.1f us total  %s\n" % (100 * us / tot, n, us, name))


def creff(tag, rep, key, cmd="-k regex:creff_march -c 1 python bench.py --steps 1 --warmup 1 (f16 plan: 11 frames per launch)"):
    p = os.path.join(OUT, rep)
    if not os.path.exists(p):
        return None
    txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    with open(os.path.join(PROF, "%s_%s_ncu.txt" % (tag, key)), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on %s\n# kernel: %s\n" % (cmd, d.get("Kernel Name", "")))
        for k in KEYS:
            if k in d:
                f.write("%-90s %s %s\n" % (k, d[k], u.get(k, "")))
        src = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_src.py"), p, "28"], capture_output=True, text=True).stdout
        f.write("\n# hottest CUDA source lines (warp-state samples; tools/ncu_src.py)\n" + src)

    def num(k):
        v = float(d[k].replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u[k]]
    return int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum"))


def main():
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    for f in os.listdir(OUT):
        if f.startswith("bench_") and f.endswith(".json") and os.path.getsize(os.path.join(OUT, f)) > 0:
            shutil.copy(os.path.join(OUT, f), os.path.join(PROF, "%s_%s" % (tag, f)))
    if os.path.exists(os.path.join(OUT, "creff_sweep.md")):
        shutil.copy(os.path.join(OUT, "creff_sweep.md"), os.path.join(PROF, tag + "_creff_sweep.md"))
    launches(tag)
    for f in ("configs.md",):
        if os.path.exists(os.path.join(OUT, f)):
            shutil.copy(os.path.join(OUT, f), os.path.join(PROF, "%s_%s" % (tag, f)))
    creff(tag, "conv_halo_f16.ncu-rep", "conv_halo", "-k regex:conv_tc_halo_kernel.*256 -s 14 -c 1 python bench.py --steps 1 --warmup 1 (one 3x3 layer of the f16 plan)")
    tp = os.path.join(PROF, "ncu_traffic.json")
    cur = json.load(open(tp)) if os.path.exists(tp) else {}
    tc = creff(tag, "creff_tc_f16.ncu-rep", "creff_tc", "-k regex:creff_tc_kernel -c 1 python bench.py --steps 1 --warmup 1 (f16 plan: 11 frames per launch)")
    tw = creff(tag, "creff_tc_warp_f16.ncu-rep", "creff_tc_warp", "-k regex:creff_tc_warp_kernel -c 1 python bench.py --steps 1 --warmup 1 (f16 plan: the MV-warp pre-pass of 11 frames)")
    tm = creff(tag, "creff_march_tf32.ncu-rep", "creff_march", "-k regex:creff_march -c 1 python bench.py --precision tf32 --steps 1 --warmup 1 (tf32 plan: 11 frames per launch)")
    if tc and tw:
        cur.update({"creff_f16": tc + tw, "creff_f16_parts": {"creff_tc_kernel": tc, "creff_tc_warp_kernel": tw},
                    "source_f16": "%s_creff_tc_ncu.txt + %s_creff_tc_warp_ncu.txt (dram__bytes_read.sum + dram__bytes_write.sum, one launch each = 11 frames)" % (tag, tag)})
    if tm:
        cur.update({"creff_tf32": tm, "source_tf32": "%s_creff_march_ncu.txt (march engine, fp32 LR feature, one launch = 11 frames)" % tag})
    cur.pop("source", None)
    json.dump(cur, open(tp, "w"), indent=1)
    print("profiles written for", tag)


if __name__ == "__main__":
    main()
