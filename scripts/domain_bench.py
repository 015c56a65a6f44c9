"""domain_Decomposition() of the reference (oracle/_ref/ref_harness) next to the one-rank version of the glue
(oracle/_ref/ref_host_dropin: extent, keys, top-level tree, Peano-Hilbert order in the library on all host cores) on the
config-0 particle count.  Host only.  Prints one JSON line.    python scripts/domain_bench.py [--ngas 1600000]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util                                   # noqa: E402
from gadget_leicester_b200 import ics         # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ngas", type=int, default=1600000)
args = ap.parse_args()
ps = ics.disc(args.ngas, seed=42)
par = ics.default_params()
par["PartAllocFactor"] = 2.5
out = dict(particles=int(ps["pos"].shape[0]), cores=os.cpu_count())
for name, b in (("reference", util.REF_HARNESS), ("glue", os.path.join(ROOT, "oracle", "_ref", "ref_host_dropin"))):
    o = util.run_oracle(b, ps, par, "K")
    out[name] = dict(domain_Decomposition_s=float(o["time_domain"][0]), NTopnodes=int(o["top_n"][0]), NTopleaves=int(o["top_n"][1]))
print(json.dumps(out))
