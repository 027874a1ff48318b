#!/usr/bin/env python
"""Warp-state sampling digest of an .ncu-rep captured with --set full --import-source on (run here, no GPU needed):
per kernel, the share of the samples by stall reason and by SASS opcode, and the executed warp instructions per opcode.
usage: python profiles/scripts/ncu_stall_digest.py gpurun_out/x.ncu-rep > profiles/r02/ncu_stall_digest.txt"""
import collections
import csv
import subprocess
import sys

STALLS = ["stall_selected", "stall_wait", "stall_long_sb", "stall_math", "stall_not_selected", "stall_dispatch", "stall_no_inst",
          "stall_short_sb", "stall_branch_resolving", "stall_mio", "stall_lg", "stall_barrier", "stall_membar", "stall_drain"]


def as_int(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    ki = rows[0].index("Kernel Name")
    seen = set()
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").split("<")[0]
        if name in seen:
            continue
        seen.add(name)
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", name,
                              "--launch-count", "1"], capture_output=True, text=True).stdout
        src = list(csv.reader(out.splitlines()))
        hdr = next((x for x in src if x and x[0] == "Address"), None)
        if hdr is None:
            continue
        ix = {h: i for i, h in enumerate(hdr)}
        data = [x for x in src if len(x) >= len(hdr) and x[0].startswith("0x")]
        total = sum(as_int(x[ix["# Samples"]]) for x in data)
        if total < 200:
            continue
        by_reason = collections.Counter()
        by_op, exec_op = collections.Counter(), collections.Counter()
        for x in data:
            tok = x[ix["Source"]].split()
            op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
            by_op[op] += as_int(x[ix["# Samples"]])
            exec_op[op] += as_int(x[ix["Instructions Executed"]])
            for s in STALLS:
                if s in ix:
                    by_reason[s] += as_int(x[ix[s]])
        n_exec = sum(exec_op.values())
        print(f"== {name}: {len(data)} SASS instructions, {total} samples, {n_exec / 1e9:.3f} G warp instructions (source-page count)")
        print("   by reason: " + ", ".join(f"{s[6:]} {v / total:.3f}" for s, v in by_reason.most_common(8)))
        print("   by opcode (share of samples | share of executed instructions):")
        for op, v in by_op.most_common(10):
            print(f"      {op:10s} {v / total:6.3f} | {exec_op[op] / max(n_exec, 1):6.3f}")


if __name__ == "__main__":
    main()
