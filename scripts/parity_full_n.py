#!/usr/bin/env python
"""Parity of the CUDA path against the compiled reference AT THE BENCHMARK'S OWN SIZE: BASELINE config 0, the 1.6 M-gas
disc + 2 sinks (tests/parity_bench.py does the comparing; bench.py keeps the 200 k record in every line).

The reference (oracle/_ref/ref_harness, one core) needs a few minutes for the sequence
  cold-start density, Barnes-Hut pass, relative pass, settled density, hmax, hydro, pair counts, one more full step,
so this is a script to run on the GPU box once per round, not a test; its record goes to profiles/.
Also answers a question of the round-1 review: how many passes does density() need on the SETTLED state, here and in
the reference ("ngb iteration k" lines of density.c:664), and which particle keeps iterating.

  python scripts/parity_full_n.py [--ngas 1600000] [--out profiles/r02_parity_full_n.json]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ngas", type=int, default=1600000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_parity_full_n.json"))
    a = ap.parse_args()
    import parity_bench
    import util
    from gadget_leicester_b200 import ics
    from gadget_leicester_b200.gadget import ForcePath
    from oracle import glio
    if not util.have_ref():
        raise SystemExit("oracle/_ref/ref_harness is missing (it is built where /root/reference exists and travels with the snapshot)")
    par = ics.default_params()
    ps = ics.disc(a.ngas, seed=42)
    n = ps["pos"].shape[0]
    ngas = int((ps["type"] == 0).sum())
    ops = parity_bench.OPS_SETTLE + " A"
    t0 = time.time()
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.glin"), os.path.join(td, "out.glou")
        glio.write_glin(fin, ps, par)
        r = subprocess.run([util.REF_HARNESS, fin, fout, ops], capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit("reference failed: " + r.stderr[-2000:])
        out = glio.read_glou(fout)
    t_ref = time.time() - t0
    # density() calls of the sequence, in order: op D (cold), op D (settled), the D inside op A; their "ngb iteration" lines
    iters, cur = [], None
    for ln in r.stdout.splitlines():
        if ln.startswith("ngb iteration"):
            cur = int(ln.split()[2].rstrip(":"))
            if cur == 1:
                iters.append(1)
            else:
                iters[-1] = cur
    R = parity_bench.golden_layout(out, n, ngas)
    rec = parity_bench.parity_record(ForcePath, ps, par, R)
    # pass counts on the settled state through the CUDA path, and who iterates
    psd = dict(ps, hsml=R["dens1__Hsml"].copy(), oldacc=R["grav1__OldAcc"].copy())
    fp = ForcePath(dict(par, ErrTolTheta=0.0))
    fp.load(psd)
    fp.gravity_tree()
    passes = []
    for _ in range(3):
        d = fp.density()
        st = fp.stats()
        passes.append(dict(passes=int(st["density_passes_max"]), slowest_index=int(st["density_slowest_index"]),
                           slowest_type=int(ps["type"][st["density_slowest_index"]]) if st["density_slowest_index"] >= 0 else None,
                           slowest_hsml=float(d["Hsml"][st["density_slowest_index"]]) if st["density_slowest_index"] >= 0 else None,
                           slowest_numngb=float(d["NumNgb"][st["density_slowest_index"]]) if st["density_slowest_index"] >= 0 else None))
        fp.set_TreeReconstructFlag(1)
        fp.gravity_tree()
    fp.close()
    rec["settled_density_passes_here"] = passes
    rec["reference_ngb_iteration_rounds_per_density_call"] = iters
    rec["reference_seconds_total"] = t_ref
    step = "%d_" % len(parity_bench.OPS_SETTLE.split())
    rec["reference_step_seconds"] = {op: float(out[step + op + "_time"][0]) for op in "GDUH" if (step + op + "_time") in out}
    rec["config"] = "BASELINE config 0: Pd_N16e5 protoplanetary disc, %d gas + 2 sinks (synthetic, seed 42)" % a.ngas
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
