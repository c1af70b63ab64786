#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample shares of one kernel in an .ncu-rep.

usage: tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [TOP] [inst|stall] [SKIP]
SKIP: matching launches to skip first (template variants share one function name).
Needs a capture made with --import-source on and code built with -lineinfo.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    key = sys.argv[4] if len(sys.argv) > 4 else "stall"
    skip = sys.argv[5] if len(sys.argv) > 5 else "0"
    txt = subprocess.run(
        ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "-k", "regex:" + rx, "--launch-skip", skip,
         "--launch-count", "1"],
        capture_output=True, text=True).stdout
    cur_file, hdr, lines = None, None, []
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file, hdr = r[1].split("/")[-1], None
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] == "":  # SASS rows have an empty line number
            continue
        i_inst = hdr.index("Instructions Executed")
        i_smp = hdr.index("Warp Stall Sampling (All Samples)")
        try:
            lines.append((cur_file, r[0], float(r[i_inst] or 0), float(r[i_smp] or 0), r[1].strip()))
        except ValueError:
            continue
    ti = sum(l[2] for l in lines) or 1.0
    ts = sum(l[3] for l in lines) or 1.0
    print(f"total warp instructions {ti:.3e}, samples {ts:.0f}")
    k = 3 if key == "stall" else 2
    for f, n, inst, smp, src in sorted(lines, key=lambda l: -l[k])[:top]:
        print(f"{f}:{n:>4} inst {100 * inst / ti:5.1f}%  stall {100 * smp / ts:5.1f}%  {src[:100]}")


if __name__ == "__main__":
    main()
