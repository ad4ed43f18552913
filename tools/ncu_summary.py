#!/usr/bin/env python
"""Summarise ncu captures into small text files under profiles/ (the .ncu-rep files themselves stay in gpurun_out/).

  python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
  python tools/ncu_summary.py full     gpurun_out/prof.ncu-rep  > profiles/rNN_full.txt
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    iname, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= ival or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        v = float(r[ival].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iunit], 1e-6)
        name = re.sub(r"\(.*", "", r[iname]).replace("void ", "")
        tot[name] += v * scale
        cnt[name] += 1
    total = sum(tot.values())
    print("# per-kernel device time over the captured launches (cold-cache, serialised: compare SHARES, not absolutes)")
    print("%-48s %8s %12s %10s %8s" % ("kernel", "launches", "total_ms", "avg_ms", "share"))
    for k in sorted(tot, key=tot.get, reverse=True):
        print("%-48s %8d %12.3f %10.4f %7.1f%%" % (k, cnt[k], tot[k], tot[k] / cnt[k], 100 * tot[k] / total))
    print("%-48s %8d %12.3f" % ("TOTAL", sum(cnt.values()), total))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==== %s" % r[hdr.index("Kernel Name")])
        for h, u, v in zip(hdr, units, r):
            if h in KEYS:
                print("%-82s %18s %s" % (h, v, u))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
