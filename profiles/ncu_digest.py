#!/usr/bin/env python
"""Digest an .ncu-rep into the few numbers we track (run here, no GPU needed).
usage: python profiles/ncu_digest.py gpurun_out/prof.ncu-rep [kernel-substring]"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
STALLS = "smsp__average_warps_issue_stalled_"


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    for r in rows[2:]:
        yield dict(zip(hdr, r))


def source(rep, kernel_id):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", kernel_id], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    for i, r in enumerate(rows):
        if r and r[0] == "Address":
            return r, rows[i + 1:]
    return None, []


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    for row in raw(rep):
        name = row.get("Kernel Name", "")
        if want not in name:
            continue
        print("==", name.split("(")[0], "id", row.get("ID"))
        for k in KEYS:
            if k in row:
                print(f"  {k:70s} {row[k]}")
        st = {k[len(STALLS):].replace("_per_issue_active.ratio", ""): float(v) for k, v in row.items()
              if k.startswith(STALLS) and k.endswith("_per_issue_active.ratio") and v}
        print("  stalls/issue:", ", ".join(f"{k}={v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))


if __name__ == "__main__":
    main()


def top_stalls(rep, kernel_sub, n=30):
    """Print the instructions with most stall samples for the kernel whose name contains kernel_sub."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    blocks, cur = {}, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = r[1]
            blocks[cur] = []
        elif cur is not None:
            blocks[cur].append(r)
    for name, rs in blocks.items():
        if kernel_sub not in name:
            continue
        hdr = rs[0]
        ix = {h: i for i, h in enumerate(hdr)}
        items, tot = [], 0
        ops = collections.Counter()
        for r in rs[1:]:
            try:
                smp, ie = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])
            except Exception:
                continue
            tot += smp
            src = r[ix["Source"]]
            op = [t for t in src.split() if not t.startswith("@")][0].split(".")[0] if src.split() else ""
            ops[op] += ie
            items.append((smp, ie, r[ix["Address"]][-5:], src[:110]))
        print("==", name.split("(")[0], "samples", tot, "warp-instr", sum(i[1] for i in items))
        print("   instr mix:", ", ".join(f"{k}={v / max(1, sum(ops.values())):.1%}" for k, v in ops.most_common(12)))
        for it in sorted(items, reverse=True)[:n]:
            print(f"   {it[0]:8d} smp {it[1]:10d} exe  {it[2]}  {it[3]}")


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "stalls":
    top_stalls(sys.argv[1], sys.argv[2])
