"""run()'s loop on the device-resident state (gl_run_begin / gl_run_kick / gl_run_steps, SURVEY.md 8(f) #2 device side)
against the compiled reference's own loop.

Harness op L (oracle/refbuild/ref_harness.c) puts a 1540-particle set on individual time steps (bins 1..6, sync point
Ti_Current = 64), calls advance_and_find_timesteps() and then runs eight iterations of run.c:63-110 -
find_next_sync_point_and_drift, domain_Decomposition, gravity / density / hydro for the ACTIVE particles,
advance_and_find_timesteps - with the unmodified reference.  It dumps the state it starts from ("begin") and the state
it arrives at ("end").  The CUDA path is given the begin state and has to arrive at the end state with everything
resident in HBM: the integer time line (bins, step starts, sync points, number of force updates) exactly, the floating-
point state within the tolerance of its FP32 force magnitudes (north_star: 1e-5; gravity agrees to ~3e-7 per
evaluation, everything else to rounding)."""
import importlib.util
import os

import numpy as np
import pytest

import util
from gadget_leicester_b200 import ics

pytestmark = pytest.mark.gpu

TBI = 4.0e-5


def _golden_module():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(util.ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg


def _reference_loop():
    mg = _golden_module()
    ps = mg._kick_set()
    par = ics.default_params(Ti_Current=16, MinEgySpec=1.8)
    par["Timebase_interval"] = TBI
    par["PartAllocFactor"] = 20.0
    o = util.run_oracle(util.REF_HARNESS, ps, par, "T S D R G D U H L")
    st = {k[len("8_L_"):]: v for k, v in o.items() if k.startswith("8_L_")}
    b = {k[len("begin_"):]: v for k, v in st.items() if k.startswith("begin_")}
    e = {k[len("end_"):]: v for k, v in st.items() if k.startswith("end_")}
    return par, b, e, st["loop"]


def _pad(a, n):
    out = np.zeros((n,) + a.shape[1:], dtype=a.dtype)
    out[: a.shape[0]] = a
    return out


def _tpar(fp, par, ti, active_upto):
    return fp.timestep_params(ti, [1 if b <= active_upto else 0 for b in range(29)], Timebase_interval=TBI, ErrTolIntAccuracy=0.025,
                              CourantFac=0.15, MaxSizeTimestep=200 * TBI, MinSizeTimestep=0.0, MinEgySpec=par["MinEgySpec"],
                              SofteningTable=par["Softening"], OuterBoundary=1.4e10, UnitVelocity_in_cm_per_s=2.98e6,
                              FeedBackVelocity=1.0, SMBHmass=0.5, InnerBoundary=2.0e-2, SinkBoundary=1.0e-4, AccDtBlackHole=0.007)


def _start(par, b):
    """a ForcePath holding the reference's begin state (in the reference's P[] order at that moment)"""
    from gadget_leicester_b200.gadget import ForcePath
    n, ng = b["Pos"].shape[0], b["VelPred"].shape[0]
    q = dict(par)
    q["ErrTolTheta"] = float(b["globals"][2])           # 0 after the first gravity_tree() (gravtree.c:396-397)
    q["Ti_Current"] = 64
    fp = ForcePath(q)
    fp.load(dict(pos=b["Pos"].copy(), vel=b["Vel"].copy(), mass=b["Mass"].copy(), type=b["Type"].copy(), hsml=b["Hsml"].copy(),
                 oldacc=b["OldAcc"].copy(), timebin=b["TimeBin"].copy(), ti_begstep=b["Ti_begstep"].copy(),
                 entropy=_pad(b["Entropy"], n), dtentropy=_pad(b["DtEntropy"], n)))
    fp.set_VelPred(b["VelPred"])
    fp.set_sph_state(_pad(b["Density"], n), _pad(b["Pressure"], n), _pad(b["DhsmlDensityFactor"], n), _pad(b["DivVel"], n),
                     _pad(b["CurlVel"], n))
    fp.set_accelerations(b["GravAccel"], _pad(b["HydroAccel"], n), _pad(b["MaxSignalVel"], n))
    fp.run_begin(64)
    return fp, n, ng


def _by_id(e, key, n):
    a = e[key]
    ids = e["id"][: a.shape[0]]
    out = np.zeros((n,) + a.shape[1:], dtype=a.dtype)
    out[ids - 1] = a
    return out


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref/ref_harness not built (needs /root/reference)")
def test_resident_loop_equals_the_reference_loop():
    par, b, e, loop = _reference_loop()
    fp, n, ng = _start(par, b)
    ids = b["id"]
    t = _tpar(fp, par, 64, 6)
    fp.advance_and_find_timesteps(t, commit_dtentropy=False)
    info = fp.run(t, 8)
    assert info["steps_done"] == 8
    assert t.Ti_Current == int(loop[0]) == 80
    st = fp.run_state(ng)
    res = fp.results()

    def mine(a):                                         # my arrays are in begin order: bring them to id order
        out = np.zeros_like(a)
        out[ids[: a.shape[0]] - 1] = a
        return out

    gas = np.zeros(n, bool)
    gas[ids[:ng] - 1] = True
    # ---- the integer time line, exactly ----
    assert np.array_equal(mine(st["TimeBin"]), _by_id(e, "TimeBin", n))
    assert np.array_equal(mine(st["Ti_begstep"]), _by_id(e, "Ti_begstep", n))
    # ---- the floating-point state ----
    worst = {}
    for key, mine_a, tol in (("Pos", st["Pos"], 1e-9), ("Vel", st["Vel"], 1e-6), ("GravAccel", res["GravAccel"], 1e-5),
                             ("Hsml", st["Hsml"], 1e-9)):
        err = util.rel_err_vec(mine(mine_a), _by_id(e, key, n))
        worst[key] = err.max()
        assert np.percentile(err, 99) < tol and err.max() < 30 * tol, util.summarize(key, err)
    for key, mine_a, tol in (("VelPred", st["VelPred"], 1e-6), ("Entropy", st["Entropy"][:ng], 1e-9), ("DtEntropy", st["DtEntropy"][:ng], 1e-6),
                             ("Density", res["Density"][:ng], 1e-9), ("Pressure", res["Pressure"][:ng], 1e-9),
                             ("HydroAccel", res["HydroAccel"][:ng], 1e-6), ("MaxSignalVel", res["MaxSignalVel"][:ng], 1e-9)):
        ref = _by_id(e, key, n)[gas]
        got = mine(_pad(np.asarray(mine_a), n))[gas]
        if ref.ndim == 1:
            scale = np.abs(ref).max()
            err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-6 * scale)
        else:
            err = util.rel_err_vec(got, ref)
        worst[key] = err.max()
        assert np.percentile(err, 99) < tol and err.max() < 100 * tol, util.summarize(key, err)
    print("resident loop vs reference loop, worst relative differences:", {k: float("%.2e" % v) for k, v in worst.items()})
    # the loop did something: eight different sync points, a changing active set
    assert 0 < info["active_last"] <= n and info["active_total"] < 8 * n
    assert not np.allclose(st["Pos"], b["Pos"])
    fp.close()


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref/ref_harness not built (needs /root/reference)")
def test_resident_kick_equals_host_kick():
    """gl_run_kick on the device against gl_gadget_advance_and_find_timesteps' statements on the host (which are
    bit-identical to the reference, tests/test_timestep.py): the first kick of op L, taken alone."""
    par, b, e, loop = _reference_loop()
    fp, n, ng = _start(par, b)
    t = _tpar(fp, par, 64, 6)
    fp.advance_and_find_timesteps(t, commit_dtentropy=False)
    st = fp.run_state(ng)
    # every particle is on an active bin at Ti = 64: all step starts move to 64, velocities change
    assert np.all(st["Ti_begstep"] == 64)
    assert np.all((st["TimeBin"] >= 1) & (st["TimeBin"] <= 6))
    # a bin may only grow onto a synchronised bin (timestep.c:147-154): at Ti = 64 that is every bin <= 6
    ty = b["Type"]
    moved = np.abs(st["Vel"] - b["Vel"]).max(axis=1) > 0
    assert moved[(ty != 2) & (ty != 3)].all() and not moved[ty == 2].any()      # dust is not kicked here (timestep.c:478-485)
    # second call at the same time: the context refuses a time that differs from the drift time
    t2 = _tpar(fp, par, 66, 6)
    with pytest.raises(Exception):
        fp.advance_and_find_timesteps(t2)
    fp.close()


def test_run_begin_needs_the_whole_state():
    from gadget_leicester_b200.gadget import ForcePath
    from gadget_leicester_b200 import lib as L
    ps = ics.cube(2000, seed=5)
    par = ics.default_params()
    ps["hsml"] = ics.guess_hsml(ps, par["DesNumNgb"])
    fp = ForcePath(par)
    fp.load(ps)
    with pytest.raises(L.GLError):
        fp.run_begin(0)          # no accelerations, no SphP state in the context yet
    fp.compute_accelerations()
    fp.run_begin(0)
    fp.close()
