#!/usr/bin/env python
"""Read ncu reports brought back from the GPU box (no GPU needed: `ncu -i`) and write the numbers the judge reads:
profiles/rNN_traffic.json (per-launch DRAM bytes of the two hot kernels, used by bench.py's `roofline*.traffic`) and a
markdown table of the headline counters.

  python scripts/ncu_extract.py --round 2 --n 50000 --snps 8192 gemm=/path/f_prof_gemm.ncu-rep lmm=/path/f_prof_lmm.ncu-rep
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("gpc__cycles_elapsed.avg.per_second", "SM clock during the capture"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (shared memory)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
]
UNIT_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def read_report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            d[h] = (v, u)
        res.append(d)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", type=int, default=2)
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--snps", type=int, default=8192)
    ap.add_argument("--tag", default="")
    ap.add_argument("reports", nargs="+", help="name=path.ncu-rep")
    a = ap.parse_args()
    traffic_path = os.path.join(ROOT, "profiles", "r%02d_traffic.json" % a.round)
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    md = ["# ncu counters, round %d%s (n = %d, %d SNPs per launch; `--clock-control none`, one launch per kernel)\n" % (a.round, " " + a.tag if a.tag else "", a.n, a.snps)]
    for spec in a.reports:
        name, path = spec.split("=", 1)
        for d in read_report(path):
            kname = d["Kernel Name"][0].split("(")[0].replace("void ", "")
            md.append("\n## %s (`%s`, from `%s`)\n\n| counter | value |\n|---|---|" % (name, kname, os.path.basename(path)))
            for key, label in KEYS:
                if key in d and d[key][0] not in ("", "n/a"):
                    md.append("| %s (`%s`) | %s %s |" % (label, key, d[key][0], d[key][1]))
            try:
                rd = float(d["dram__bytes_read.sum"][0].replace(",", "")) * UNIT_BYTES[d["dram__bytes_read.sum"][1]]
                wr = float(d["dram__bytes_write.sum"][0].replace(",", "")) * UNIT_BYTES[d["dram__bytes_write.sum"][1]]
                base = kname.split("<")[0].split("::")[-1]
                traffic[base] = {"n": a.n, "snps_per_launch": a.snps, "dram_read_bytes": rd, "dram_write_bytes": wr,
                                 "duration_ms": float(d["gpu__time_duration.sum"][0].replace(",", "")) * (1e-3 if d["gpu__time_duration.sum"][1] == "us" else 1.0),
                                 "source": os.path.basename(path)}
            except Exception as ex:                                    # a capture without memory counters
                md.append("\n(no DRAM counters: %r)" % ex)
    json.dump(traffic, open(traffic_path, "w"), indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
