"""Generates the committed golden fixtures by running the UNMODIFIED reference (oracle/_ref/ref_harness,
built by oracle/refbuild from /root/reference) in the build container.

    python tests/golden/make_golden.py

Each fixture is one .npz holding the input particle set, the parameters and every array the reference
produced for the op sequence below, re-ordered to input order (the reference permutes P[] into
Peano-Hilbert order; ids make that reversible).  /root/reference is NOT needed to USE the fixtures.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gadget_leicester_b200 import ics  # noqa: E402
import util  # noqa: E402

OPS = "K T S D R G G D U H N C P"
# stage prefixes of OPS (position_op_)
STAGES = dict(keys="0_K_", tree0="1_T_", hsml0="2_S_", dens1="3_D_", grav1="5_G_", grav2="6_G_", dens2="7_D_",
              hydro="9_H_", nodes="10_N_", pairs="11_C_", pot="12_P_")

CASES = {
    "cube_2k": lambda: (ics.cube(2000, seed=11), ics.default_params()),
    "disc_3k_sinks": lambda: (ics.disc(3000, seed=12), ics.default_params()),
    "evrard_2k": lambda: (ics.evrard(2000, seed=13), ics.default_params()),
    "plummer_2k": lambda: (ics.plummer(2000, seed=14),
                           ics.default_params(Softening=[1e-5, ics.plummer_softening(2000), 5e-5, 1e-7, 0.01, 5e-5])),
    # ragged: 333 gas + one particle each of types 1 (halo), 2 (dust target), 5 (sink target)
    "mixed_336": lambda: (_mixed(), ics.default_params()),
    # individual time bins + entropy rates: exercises the viscosity limiter (hydra.c:1583-1597) and the pressure
    # prediction (density.c:490-496), both of which vanish for TimeBin == 0
    "cube_timebins_2k": lambda: (_timebins(), ics.default_params(Ti_Current=16)),
    # swallowed (Mass == 0) gas particles still sitting in the tree between two domain decompositions
    # (density.c:832, hydra.c:1236, forcetree.c:2214)
    "cube_zero_mass_2k": lambda: (_zero_mass(), ics.default_params()),
}


def _mixed():
    ps = ics.cube(333, seed=15)
    g = np.random.Generator(np.random.Philox(key=16))
    extra = 3
    out = {}
    n = 333 + extra
    out["pos"] = np.vstack([ps["pos"], g.random((extra, 3))])
    out["vel"] = np.vstack([ps["vel"], np.zeros((extra, 3))])
    out["mass"] = np.concatenate([ps["mass"], [0.05, 0.02, 0.01]])
    out["type"] = np.concatenate([ps["type"], [1, 2, 5]]).astype(np.int32)
    out["hsml"] = np.zeros(n)
    out["id"] = np.arange(1, n + 1, dtype=np.int32)
    out["timebin"] = np.zeros(n, dtype=np.int32)
    out["oldacc"] = np.zeros(n)
    out["entropy"] = np.concatenate([ps["entropy"], np.zeros(extra)])
    out["dtentropy"] = np.zeros(n)
    return out


def _timebins():
    ps = ics.cube(2000, seed=21)
    g = np.random.Generator(np.random.Philox(key=22))
    ps["timebin"] = g.integers(0, 7, size=2000).astype(np.int32)
    ps["dtentropy"] = (g.random(2000) - 0.5) * 2.0e3
    ps["vel"] *= 20.0           # strong compression/shear so that the viscosity and its limiter are active
    return ps


def _zero_mass():
    ps = ics.cube(2000, seed=23)
    g = np.random.Generator(np.random.Philox(key=24))
    ps["oldacc"][g.choice(2000, 100, replace=False)] = -1.0   # swallowed after the initial-Hsml step (see ref_harness.c op S)
    return ps


def main():
    if not util.have_ref():
        raise SystemExit("oracle/_ref/ref_harness missing: run `make -C oracle/refbuild` (needs /root/reference)")
    for name, mk in CASES.items():
        ps, par = mk()
        # the shipped run uses 10 (.txt PartAllocFactor); a tiny MaxPart trips a heap bug in the reference's top-node retry path
        par["PartAllocFactor"] = max(10.0, 40000.0 / ps["pos"].shape[0])
        n = ps["pos"].shape[0]
        ngas = int((ps["type"] == 0).sum())
        ref = util.run_oracle(util.REF_HARNESS, ps, par, OPS)
        out = {"in_" + k: v for k, v in ps.items()}
        out["params_json"] = np.frombuffer(json.dumps(par).encode(), dtype=np.uint8)
        for stage, pre in STAGES.items():
            for k, v in ref.items():
                if not k.startswith(pre) or k.endswith("_id") or k.endswith("_time"):
                    continue
                field = k[len(pre):]
                if (pre + "id") in ref and v.shape[0] in (n, ngas) and field not in ("nodes", "nodeflags", "domain"):
                    v = util.by_id(ref, pre, field, v.shape[0])
                out[stage + "__" + field] = v
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, n, "particles ->", os.path.getsize(path) // 1024, "KiB")


def drift_fixture():
    """State before / after the reference's own drift_particle (predict.c:129) for tests/test_drift.py: a settled
    force evaluation (so GravAccel, HydroAccel, DivVel, Density, Pressure are real), then harness op X."""
    ps, par = _timebins(), ics.default_params(Ti_Current=16)
    par["PartAllocFactor"] = 20.0
    ref = util.run_oracle(util.REF_HARNESS, ps, par, "T S D R G D U H X")
    out = {k[len("8_X_"):]: v for k, v in ref.items() if k.startswith("8_X_")}
    os.makedirs(os.path.join(HERE, "aux"), exist_ok=True)
    path = os.path.join(HERE, "aux", "drift_cube_timebins_2k.npz")
    np.savez_compressed(path, **out)
    print("drift fixture ->", os.path.getsize(path) // 1024, "KiB", sorted(out)[:6], "...")


def _kick_set():
    """1500 gas particles on individual time bins with entropy rates, plus 40 particles of the types the time-step
    code treats specially: 1 (plain), 2 (dust: no gravity kick, half the step), 4, 5 (sinks: accretion time step)."""
    n_gas, extra = 1500, 40
    ps = ics.cube(n_gas, seed=31)
    g = np.random.Generator(np.random.Philox(key=32))
    n = n_gas + extra
    out = {}
    out["pos"] = np.vstack([ps["pos"], g.random((extra, 3))])
    out["vel"] = np.vstack([ps["vel"] * 200.0, g.standard_normal((extra, 3)) * 0.1])
    out["mass"] = np.concatenate([ps["mass"], g.random(extra) * 0.002 + 1e-4])
    out["mass"][n_gas + 3] = 0.6            # one sink above All.SMBHmass = 0.5: the InnerBoundary branch
    out["type"] = np.concatenate([ps["type"], np.array([1, 2, 4, 5] * (extra // 4))]).astype(np.int32)
    out["hsml"] = np.zeros(n)
    out["id"] = np.arange(1, n + 1, dtype=np.int32)
    out["timebin"] = g.integers(1, 7, size=n).astype(np.int32)
    out["oldacc"] = np.zeros(n)
    out["entropy"] = np.concatenate([ps["entropy"], np.zeros(extra)])
    out["dtentropy"] = np.concatenate([(g.random(n_gas) - 0.5) * 40.0, np.zeros(extra)])
    return out


def kick_run(binary):
    """A settled force evaluation, then harness op V (advance_and_find_timesteps); returns the op's dump."""
    ps = _kick_set()
    par = ics.default_params(Ti_Current=16, MinEgySpec=1.8)
    par["Timebase_interval"] = 4.0e-5
    par["PartAllocFactor"] = 20.0
    ref = util.run_oracle(binary, ps, par, "T S D R G D U H V")
    return {k[len("8_V_"):]: v for k, v in ref.items() if k.startswith("8_V_")}


def kick_fixture():
    """State before / after the reference's own advance_and_find_timesteps (timestep.c:29) for tests/test_timestep.py."""
    out = kick_run(util.REF_HARNESS)
    os.makedirs(os.path.join(HERE, "aux"), exist_ok=True)
    path = os.path.join(HERE, "aux", "kick_mixed_1540.npz")
    np.savez_compressed(path, **out)
    old, new = out["pre_TimeBin"], out["post_TimeBin"]
    print("kick fixture ->", os.path.getsize(path) // 1024, "KiB; bins before", np.bincount(old, minlength=10)[:10],
          "after", np.bincount(new, minlength=10)[:10], "changed", int((old != new).sum()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "kick":
        kick_fixture()
    else:
        main()
        drift_fixture()
        kick_fixture()
