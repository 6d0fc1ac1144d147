#!/usr/bin/env python
"""Executed-instruction histogram of one kernel of an .ncu-rep (source page, SASS view):
opcode shares, warp-stall sample shares, and the split between CTA barriers.

    python tools/sass_opcode_hist.py report.ncu-rep kernel_regex
"""
import csv
import io
import subprocess
import sys
from collections import Counter


def main():
    rep, regex = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                          "--kernel-name", "regex:" + regex], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][1][:100])
    hdr, rows = rows[1], rows[2:]
    for i, r in enumerate(rows):  # several launches match: keep the first one
        if r and r[0] == "Kernel Name":
            rows = rows[:i]
            break
    rows = [r for r in rows if len(r) == len(hdr)]
    ia, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
    ist = hdr.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(r[ia]) for r in rows)
    ts = max(1, sum(int(r[ist]) for r in rows))
    print("warp instructions executed: %d, stall samples: %d" % (tot, ts))
    c, s = Counter(), Counter()
    for r in rows:
        t = r[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        c[op] += int(r[ia])
        s[op] += int(r[ist])
    for op, n in c.most_common(24):
        print("%-10s %12d %5.1f %%   stall samples %5.1f %%" % (op, n, 100.0 * n / tot, 100.0 * s[op] / ts))
    seg_i = seg_s = k = 0
    for r in rows:
        seg_i += int(r[ia])
        seg_s += int(r[ist])
        if "BAR.SYNC" in r[isrc] or "EXIT" in r[isrc]:
            if seg_i * 200 > tot:
                print("segment %2d (up to %-28s): %5.1f %% of instructions, %5.1f %% of samples"
                      % (k, r[isrc].strip()[:28], 100.0 * seg_i / tot, 100.0 * seg_s / ts))
            seg_i = seg_s = 0
            k += 1


if __name__ == "__main__":
    main()
