"""Summarises an `ncu --set full` capture for profiles/: one block per launch with the counters DESIGN.md quotes.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > gpurun_out/x_raw.csv
    python scripts/ncu_summary.py gpurun_out/x_raw.csv [--hbm-peak 6548.5] > profiles/rN_ncu_x_summary.txt

For every launch it also derives the achieved DRAM bandwidth (dram bytes read + written / duration) and its fraction
of the measured HBM copy peak (MEASURED_PEAKS.json), which is what the HBM-bound SIMT kernels are judged against.
"""
import argparse
import csv
import json
import os
import re

KEYS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TO_SEC = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--hbm-peak", type=float, default=None, help="GB/s; default MEASURED_PEAKS.json hbm copy peak")
    opt = ap.parse_args()
    peak = opt.hbm_peak
    if peak is None:
        try:
            mp = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
            peak = float(next(v for k, v in mp.items() if "hbm" in k.lower() and isinstance(v, (int, float))))
        except Exception:
            peak = 6548.5
    rows = list(csv.reader(l for l in open(opt.csv) if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    col = {}
    for i, h in enumerate(hdr):
        col.setdefault(h, i)
        col.setdefault(h.split("TriageCompute.")[-1], i)      # full sets prefix some metrics with their section
    name_i = hdr.index("Kernel Name")
    for n, r in enumerate(rows[2:]):
        print(f"--- launch {n}")
        print("  Kernel Name =", re.sub(r"\(.*", "", r[name_i]))
        vals = {}
        for k in KEYS:
            if k in col and r[col[k]] != "":
                vals[k] = (r[col[k]], units[col[k]])
                print(f"  {k} = {r[col[k]]} {units[col[k]]}")
        try:
            rd, ru = vals["dram__bytes_read.sum"]
            wr, wu = vals["dram__bytes_write.sum"]
            t, tu = vals["gpu__time_duration.sum"]
            nbytes = float(rd.replace(",", "")) * TO_BYTES[ru] + float(wr.replace(",", "")) * TO_BYTES[wu]
            sec = float(t.replace(",", "")) * TO_SEC[tu]
            gbs = nbytes / sec / 1e9
            print(f"  derived: DRAM traffic {nbytes / 1e6:.1f} MB in {sec * 1e6:.1f} us = {gbs:.0f} GB/s "
                  f"= {gbs / peak:.3f} of the measured HBM copy peak ({peak:.1f} GB/s)")
        except (KeyError, ValueError):
            pass
        print()


if __name__ == "__main__":
    main()
