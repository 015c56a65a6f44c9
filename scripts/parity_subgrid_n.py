#!/usr/bin/env python
"""Sub-grid loops against the compiled reference at a larger size than the test suite uses (tests/test_gpu_subgrid.py runs
20 k gas + 2 k grains): same comparison, same bar.   python scripts/parity_subgrid_n.py [--ngas 400000] [--ndust 40000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ngas", type=int, default=400000)
    ap.add_argument("--ndust", type=int, default=40000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_subgrid_parity_n.json"))
    a = ap.parse_args()
    import util
    import test_gpu_subgrid as T
    from gadget_leicester_b200 import ics
    par = ics.default_params(BlackHoleFeedbackFactor=1e-3, SinkBoundary=0.02, MinEgySpec=1e-6)
    ps = ics.disc_subgrid(a.ngas, a.ndust, seed=21)
    ps["timebin"][:] = 3 + (ps["id"] % 3)
    t0 = time.time()
    out = util.run_oracle(util.REF_HARNESS, ps, par, T.OPS)
    rec = dict(config="%d gas + %d grains + 4 sinks (ics.disc_subgrid, seed 21)" % (a.ngas, a.ndust), reference_seconds=time.time() - t0)
    ops = T.OPS.split()

    def rel(x, y):
        x, y = np.asarray(x, float), np.asarray(y, float)
        scale = np.abs(y).max() if y.size else 0.0
        e = np.abs(x - y) / np.where(np.abs(y) > 0, np.abs(y), 1.0)
        e = np.where(np.abs(x - y) <= 1e-9 * scale, 0.0, e)
        return float(e.max()) if e.size else 0.0
    # sinks
    pre, post, q = T._split(T._stage(out, ops.index("B"), "B"))
    n = pre["Pos"].shape[0]
    fp = T._forcepath(pre, q, par)
    S = dict(Vel=pre["Vel"].copy(), ID=pre["id"].copy(), SwallowID=pre["SwallowID"].copy(), BH_Mass=pre["BH_Mass"].copy(),
             BH_Mdot=pre["BH_Mdot"].copy(), AccDisc_Mass=pre["AccDisc_Mass"].copy(), Dust_Mass=pre["Dust_Mass"].copy(),
             Total_Mass=pre["Total_Mass"].copy(), Injected_BH_Energy=T._pad(pre["Injected_BH_Energy"], n), SinkDensity=pre["SinkDensity"].copy())
    mass, counts = fp.blackhole_accretion(S, active=pre["active"])
    fp.close()
    t = pre["Type"]
    gone = (pre["Mass"] > 0) & (post["Mass"] == 0)
    ng = post["Injected_BH_Energy"].shape[0]
    rec["blackhole_accretion"] = dict(swallowed_gas_sinks_dust=list(counts), reference=[int((gone & (t == k)).sum()) for k in (0, 5, 2)],
                                      swallowid_mismatches=int((S["SwallowID"] != post["SwallowID"]).sum()),
                                      mass_zero_mismatches=int(((mass == 0) != (post["Mass"] == 0)).sum()),
                                      max_rel=dict(Mass=rel(mass, post["Mass"]), Vel=rel(S["Vel"], post["Vel"]),
                                                   BH_Mass=rel(S["BH_Mass"], post["BH_Mass"]), Dust_Mass=rel(S["Dust_Mass"], post["Dust_Mass"]),
                                                   Injected_BH_Energy=rel(S["Injected_BH_Energy"][:ng], post["Injected_BH_Energy"])))
    # dust density
    pre, post, q = T._split(T._stage(out, ops.index("Y"), "Y"))
    fp = T._forcepath(pre, q, par)
    d7 = pre["DustParticleDensity"].copy()
    fp.dust_density(d7, active=pre["active"])
    fp.close()
    rec["dust_density"] = dict(nonzero=int((post["DustParticleDensity"] > 0).sum()), max_rel=rel(d7, post["DustParticleDensity"]))
    # dust drag
    pre, post, q = T._split(T._stage(out, ops.index("Z"), "Z"))
    fp = T._forcepath(pre, q, par)
    S = dict(Vel=pre["Vel"].copy(), GravAccel=pre["GravAccel"].copy(), GasDensity=pre["SinkDensity"].copy(), GasEntropy=pre["SinkEntropy"].copy(),
             GasVel=pre["SinkGasVel"].copy(), DustParticleDensity=pre["DustParticleDensity"].copy(), DustRadius=pre["DustRadius"].copy(),
             DustVcoll=pre["DustVcoll"].copy(), DeltaDustMomentum=pre["DeltaDustMomentum"].copy(), NewDragAcc=pre["NewDragAcc"].copy(),
             DeltaDragEnergy=pre["DeltaDragEnergy"].copy(), Entropy=T._pad(pre["Entropy"], n), DragHeating=T._pad(pre["DragHeating"], n))
    fp.dust_drag(S, active=pre["active"])
    fp.close()
    ng = post["Entropy"].shape[0]
    rec["dust_drag"] = dict(gas_touched=int((post["Entropy"] != pre["Entropy"]).sum()),
                            max_rel=dict(Vel=rel(S["Vel"], post["Vel"]), DeltaDustMomentum=rel(S["DeltaDustMomentum"], post["DeltaDustMomentum"]),
                                         DeltaDragEnergy=rel(S["DeltaDragEnergy"], post["DeltaDragEnergy"]), DustVcoll=rel(S["DustVcoll"], post["DustVcoll"]),
                                         Entropy=rel(S["Entropy"][:ng], post["Entropy"]), DragHeating=rel(S["DragHeating"][:ng], post["DragHeating"])))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
