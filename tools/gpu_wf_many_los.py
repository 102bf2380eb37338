"""Weighting functions with many lines of sight: per-kernel device time of one solve (2 000 wavelengths, 16 streams,
20 layers, O3 + NO2 + aerosol-extinction mappings) for a given LOS count.  Run once per process:
    python tools/gpu_wf_many_los.py 40                      # k_wf_layer_fast in LOS tiles
    SK_B200_WF_TILE=-1 python tools/gpu_wf_many_los.py 40   # generic thread-per-problem kernel (the former fallback)
Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import sasktran2_b200 as sk  # noqa: E402
from sasktran2_b200 import scenarios  # noqa: E402

nlos = int(sys.argv[1]) if len(sys.argv) > 1 else 40
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
sc = scenarios.config2(nwavel=nw, nlayers=20, nstr=16, nlos=nlos, with_wf=True)
_, _, _, eng, atm = sk.engine_for_scenario(sc)
eng.set_workspace_gb(48.0)
eng.stage(atm)
eng.solve_staged()
eng.solve_staged()
t = {k: round(v, 3) for k, v in eng.timings_ms().items() if v > 0}
res = eng.fetch()
chk = {k: float(np.abs(np.asarray(v)).sum()) for k, v in res.items() if not k.startswith("_")}
print(json.dumps({"nlos": nlos, "nwavel": nw, "wf_tile_env": os.environ.get("SK_B200_WF_TILE", ""), "kernel_ms": t,
                  "abs_sums": chk}))
