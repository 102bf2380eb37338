#!/usr/bin/env python
"""SASS opcode census of the main kernels of sasktran2_b200/libsasktran2_b200.so (cuobjdump -sass): evidence of which
hardware paths a kernel uses - UBLKCP / SYNCS = TMA bulk copies completing on mbarriers, LDGSTS = cp.async, REDUX = warp
reductions of the pivot search, DMMA = FP64 tensor-core MMA (none: DESIGN.md section 4).
Usage: python tools/sass_census.py > profiles/sass_census_<tag>.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEEP = ("k_twostream", "k_bvp_v2<8, true", "k_wf_layer_fast<8, 1>", "k_bvp_tsolve<8>", "k_limb_coef_lanes<8>", "k_limb_integrate",
        "k_limb_table", "k_limb_phase", "k_eig_jacobi<8>", "k_eig_setup<8>", "k_layer_post<8>", "k_bvp_multi<8, 2>", "k_surface_general",
        "k_brdf_expand_snow", "k_radiance", "k_wf_chain_warp")
COLS = ("DFMA", "DMUL", "DADD", "UBLKCP", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "SHFL", "REDUX", "DMMA", "MUFU")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(ROOT / "sasktran2_b200" / "libsasktran2_b200.so")], capture_output=True,
                          text=True, check=True).stdout
    cur, counts = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            counts[cur][m.group(1).split(".")[0]] += 1
    rows = []
    for f, c in counts.items():
        name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().replace("disco::", "")
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        if any(k in name for k in KEEP):
            rows.append((name, sum(c.values())) + tuple(c[k] for k in COLS))
    print(("%-34s %7s" + " %6s" * len(COLS)) % (("kernel", "instr") + COLS))
    for r in sorted(rows):
        print(("%-34s %7d" + " %6d" * len(COLS)) % r)


if __name__ == "__main__":
    sys.exit(main())
