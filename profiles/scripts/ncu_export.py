#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full of one bench step) into the two tracked files bench.py and the docs read:
  profiles/r02/ncu_raw_<tag>.csv   the raw-page metrics we quote, one row per kernel launch
  profiles/r02/ncu_kernels.csv     kernel, frames it processed, dram bytes, dram_bytes_per_frame, ms (bench.py: roofline.traffic)
usage: python profiles/scripts/ncu_export.py gpurun_out/x.ncu-rep <tag> <frames_all_tracks> <frames_multiband_tracks>"""
import csv
import os
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
ALL = ("k_eq", "k_kweight_energy", "k_apply_gain", "k_tail_peak", "k_block_hist", "k_finalize", "k_limiter", "k_true_peak")


def main():
    rep, tag, frames_all, frames_mb = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(here, "r02"), exist_ok=True)
    ki = hdr.index("Kernel Name")
    cols = [(k, hdr.index(k)) for k in KEEP if k in hdr]
    with open(os.path.join(here, "r02", f"ncu_raw_{tag}.csv"), "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel"] + [f"{k} [{units[i]}]" for k, i in cols])
        for r in data:
            w.writerow([r[ki].split("(")[0].replace("ame::", "").replace("void ", "")] + [r[i] for _, i in cols])
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    agg, seen = {}, set()
    for r in data:
        name = r[ki].split("(")[0].replace("ame::", "").replace("void ", "").split("<")[0]
        if name in seen:                                               # the capture may run into the next step: one launch per kernel
            continue
        seen.add(name)
        name = {"k_compact": "k_window_flag", "k_tile_prefix": "k_window_flag"}.get(name, name)   # timed (and counted) with k_window_flag
        rd = float(r[hdr.index("dram__bytes_read.sum")]) * scale[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(r[hdr.index("dram__bytes_write.sum")]) * scale[units[hdr.index("dram__bytes_write.sum")]]
        ms = float(r[hdr.index("gpu__time_duration.sum")]) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[units[hdr.index("gpu__time_duration.sum")]]
        a = agg.setdefault(name, [0.0, 0.0])
        a[0] += rd + wr
        a[1] += ms
    with open(os.path.join(here, "r02", "ncu_kernels.csv"), "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel", "frames", "dram_bytes", "dram_bytes_per_frame", "ms", "capture"])
        for name, (b, ms) in agg.items():
            fr = frames_all if name in ALL else frames_mb
            w.writerow([name, int(fr), int(b), round(b / fr, 4), round(ms, 4), tag])
    print("wrote", os.path.join(here, "r02"))


if __name__ == "__main__":
    main()
