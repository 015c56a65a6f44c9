#!/usr/bin/env python
"""The device-resident loop against the compiled reference's own run() loop (harness op L, eight iterations on individual
time steps) at a larger size than the test suite uses (tests/test_gpu_resident.py: 1540 particles): a disc with two
sinks, time bins 1..6.   python scripts/parity_resident_n.py [--ngas 100000]"""
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
    ap.add_argument("--ngas", type=int, default=100000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_resident_parity_n.json"))
    a = ap.parse_args()
    import util
    import test_gpu_resident as T
    from gadget_leicester_b200 import ics
    ps = ics.disc(a.ngas, seed=13)
    n = ps["pos"].shape[0]
    g = np.random.Generator(np.random.Philox(key=5))
    ps["timebin"] = g.integers(1, 7, size=n).astype(np.int32)
    ps["hsml"] = np.zeros(n)
    par = ics.default_params(Ti_Current=16, MinEgySpec=0.0)
    par["Timebase_interval"] = T.TBI
    par["PartAllocFactor"] = 20.0
    t0 = time.time()
    o = util.run_oracle(util.REF_HARNESS, ps, par, "T S D R G D U H L")
    t_ref = time.time() - t0
    st = {k[len("8_L_"):]: v for k, v in o.items() if k.startswith("8_L_")}
    b = {k[len("begin_"):]: v for k, v in st.items() if k.startswith("begin_")}
    e = {k[len("end_"):]: v for k, v in st.items() if k.startswith("end_")}
    par["MinEgySpec"] = 0.0
    fp, n, ng = T._start(par, b)
    ids = b["id"]
    t = T._tpar(fp, par, 64, 6)
    t0 = time.time()
    fp.advance_and_find_timesteps(t, commit_dtentropy=False)
    info = fp.run(t, 8)
    t_gpu = time.time() - t0
    s = fp.run_state()
    res = fp.results()

    def mine(x):
        out = np.zeros_like(x)
        out[ids[: x.shape[0]] - 1] = x
        return out

    gas = np.zeros(n, bool)
    gas[ids[:ng] - 1] = True
    rec = dict(config="%d gas + 2 sinks (ics.disc, seed 13), time bins 1..6, eight iterations of run()'s loop from Ti = 64" % a.ngas,
               reference_seconds=t_ref, device_loop_seconds=t_gpu, steps=int(info["steps_done"]), Ti_Current=[int(t.Ti_Current), int(st["loop"][0])],
               force_updates=int(info["active_total"]),
               timebin_mismatches=int((mine(s["TimeBin"]) != T._by_id(e, "TimeBin", n)).sum()),
               ti_begstep_mismatches=int((mine(s["Ti_begstep"]) != T._by_id(e, "Ti_begstep", n)).sum()))

    def stats(err):
        return dict(p50=float(np.median(err)), p99=float(np.percentile(err, 99)), max=float(err.max()))
    for key, arr in (("Pos", s["Pos"]), ("Vel", s["Vel"]), ("GravAccel", res["GravAccel"]), ("Hsml", s["Hsml"])):
        rec[key] = stats(util.rel_err_vec(mine(arr), T._by_id(e, key, n)))
    for key, arr in (("VelPred", s["VelPred"]), ("Entropy", s["Entropy"]), ("DtEntropy", s["DtEntropy"]), ("Density", res["Density"]),
                     ("Pressure", res["Pressure"]), ("HydroAccel", res["HydroAccel"]), ("MaxSignalVel", res["MaxSignalVel"])):
        ref = T._by_id(e, key, n)[gas]
        got = mine(T._pad(np.asarray(arr), n))[gas]
        if ref.ndim == 1:
            err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-6 * np.abs(ref).max())
        else:
            err = util.rel_err_vec(got, ref)
        rec[key] = stats(err)
    rec["ok"] = bool(rec["timebin_mismatches"] <= max(1, n // 100000) and rec["Pos"]["p99"] < 1e-9 and rec["GravAccel"]["p99"] < 1e-5)
    fp.close()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
