"""gl_gadget_run_steps alone (bench.py's resident_loop.through_structs leg) with the phase times of GL_DEBUG_E2E."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gadget_leicester_b200 import ics  # noqa: E402
from gadget_leicester_b200.gadget import ForcePath  # noqa: E402

ngas = int(sys.argv[1]) if len(sys.argv) > 1 else 1600000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
par = ics.default_params()
ps = ics.disc(ngas, seed=42)
ps["hsml"] = ics.guess_hsml(ps, par["DesNumNgb"])
ps["hsml"][ps["type"] == 5] = par["Softening"][5]
n = ps["pos"].shape[0]
fp = ForcePath(par)
fp.load(ps)
order = fp.peano_hilbert_order(int((ps["type"] == 0).sum()))
fp.close()
ps = {a: (v[order] if hasattr(v, "shape") and v.shape[0] == n else v) for a, v in ps.items()}
print(json.dumps(bench.resident_block(ForcePath, par, ps, n, k)))
