#!/usr/bin/env python
"""Top source lines by warp-stall samples for the kernels matching a regex in an `ncu --set full --import-source on`
report.  Usage: python tools/ncu_source_top.py report.ncu-rep <kernel regex> [n_lines]"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}",
                          "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    hdr, cur = None, None
    seen = set()
    agg, inst, src = collections.Counter(), collections.Counter(), {}
    for r in csv.reader(io.StringIO(raw)):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) == 2:
            if r[0] == "Function Name" and r[1] not in seen:
                seen.add(r[1])
                print("#", r[1])
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        try:
            ln, s, n = int(r[0]), int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        agg[(cur, ln)] += s
        inst[(cur, ln)] += n
        src[(cur, ln)] = r[1].strip()
    tot = sum(agg.values()) or 1
    print(f"# total stall samples {tot}, warp instructions executed {sum(inst.values()) / 1e6:.0f} M")
    print("# share of samples | warp instructions (M) | file:line | source")
    for (f, ln), v in agg.most_common(top):
        print(f"{100 * v / tot:5.2f}% {inst[(f, ln)] / 1e6:8.1f}  {f}:{ln}  {src[(f, ln)][:110]}")


if __name__ == "__main__":
    main()
