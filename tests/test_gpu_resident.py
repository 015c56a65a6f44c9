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
    st = fp.run_state()
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
    st = fp.run_state()
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


def _loop_state(fp):
    st = fp.run_state()
    st.update(fp.results())
    return st


def ranks_vs_one(world, nsteps=8, n_gas=24000, exchange_after=None, skew=False, by_cost=True):
    """the loop on `world` in-process ranks and in one context from the same start; returns (errors, info of the one
    context, time-line mismatches)"""
    import threading
    from gadget_leicester_b200 import lib as L
    from gadget_leicester_b200.gadget import ForcePath
    ps = ics.disc(n_gas, seed=11)
    par = ics.default_params()
    par["Timebase_interval"] = 1.0e-4
    ps["hsml"] = ics.guess_hsml(ps, par["DesNumNgb"])
    ps["hsml"][ps["type"] == 5] = par["Softening"][5]
    n = ps["pos"].shape[0]
    ngas = int((ps["type"] == 0).sum())
    g = np.random.Generator(np.random.Philox(key=77))
    ps["timebin"] = g.integers(1, 7, size=n).astype(np.int32)
    ps["ti_begstep"] = (64 - (1 << ps["timebin"])).astype(np.int32)
    one = ForcePath(par)
    one.load(ps)
    order = one.peano_hilbert_order(ngas)
    pso = {k: (np.ascontiguousarray(v[order]) if hasattr(v, "shape") and v.shape[0] == n else v) for k, v in ps.items()}

    def tpar(fp):
        # (the accretion time-step criterion of the sinks is switched off: it wants steps far below this time line's tick)
        return fp.timestep_params(64, [1 if b <= 6 else 0 for b in range(29)], Timebase_interval=par["Timebase_interval"],
                                  ErrTolIntAccuracy=0.025, CourantFac=0.15, MaxSizeTimestep=64 * par["Timebase_interval"], MinSizeTimestep=0.0,
                                  MinEgySpec=0.0, SofteningTable=par["Softening"], OuterBoundary=1.4e10, UnitVelocity_in_cm_per_s=2.98e6,
                                  FeedBackVelocity=1.0, SMBHmass=0.5, InnerBoundary=0.0, SinkBoundary=0.0, AccDtBlackHole=0.007)

    moved = {}

    def drive(fp, p, ids=None, rank=None):
        fp.set_params(Ti_Current=64)
        fp.load(p)
        fp.compute_accelerations()          # Barnes-Hut pass + cold-start density: the state the loop starts from
        fp.set_TreeReconstructFlag(1)       # a new step: new tree (node maxsoft / hmax from the settled Hsml on ranks and in one context alike)
        fp.compute_accelerations()          # relative criterion
        fp.run_begin(64)
        t = tpar(fp)
        fp.advance_and_find_timesteps(t)
        if exchange_after is None or ids is None:
            info = fp.run(t, nsteps)
        else:
            fp.run_set_ids(ids)
            info = fp.run(t, exchange_after)
            moved[rank] = fp.domain_exchange(by_cost=by_cost) + (len(ids),)
            info2 = fp.run(t, nsteps - exchange_after)
            for k in ("steps_done", "active_total", "interactions"):
                info[k] += info2[k]
        st = _loop_state(fp)
        if ids is not None and exchange_after is not None:
            st["ids"] = fp.run_ids()
        return st, info, int(t.Ti_Current)

    ref, info1, ti1 = drive(one, pso)
    one.close()
    lib = L.load_library()
    group = lib.gl_comm_local_group_create(world)
    m = -(-n // world)
    # skew: a deliberately bad first assignment (rank 0 holds half of everything) for the exchange to repair
    cuts = [0] + [min(n, (r + 1) * m) for r in range(world)]
    if skew:
        cuts = [0, n // 2] + [n // 2 + (n - n // 2) * (r + 1) // (world - 1) for r in range(world - 1)]
    res, err = [None] * world, [None] * world

    def worker(r):
        try:
            lo, hi = cuts[r], cuts[r + 1]
            psr = {k: (np.ascontiguousarray(v[lo:hi]) if hasattr(v, "shape") and v.shape[0] == n else v) for k, v in pso.items()}
            fp = ForcePath(par)
            fp.comm_init_local(r, group, world)
            res[r] = drive(fp, psr, ids=np.arange(lo, hi, dtype=np.int32), rank=r)
            fp.close()
        except Exception as ex:   # noqa: BLE001
            err[r] = ex
    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    lib.gl_comm_local_group_destroy(group)
    for e in err:
        if e is not None:
            raise e
    if exchange_after is not None:
        ids = np.concatenate([r[0]["ids"] for r in res])
        assert np.array_equal(np.sort(ids), np.arange(n)), "every particle is held by exactly one rank"
        back = np.argsort(ids)                                     # rows of the concatenated ranks -> the order of the one context
        cat = lambda k: np.concatenate([r[0][k] for r in res])[back]       # noqa: E731
    else:
        cat = lambda k: np.concatenate([r[0][k] for r in res])       # noqa: E731
    line = dict(Ti_Current=[r[2] != ti1 for r in res].count(True), TimeBin=int((cat("TimeBin") != ref["TimeBin"]).sum()),
                Ti_begstep=int((cat("Ti_begstep") != ref["Ti_begstep"]).sum()),
                active_total=int(sum(r[1]["active_total"] for r in res) - info1["active_total"]))
    gas = pso["type"] == 0
    errs = {}
    for k in ("Pos", "Vel", "GravAccel", "Hsml"):
        errs[k] = util.rel_err_vec(cat(k), ref[k])
    for k in ("VelPred", "Entropy", "Density", "HydroAccel", "Pressure", "MaxSignalVel"):
        errs[k] = util.rel_err_vec(cat(k)[gas], ref[k][gas])
    if exchange_after is not None:
        return errs, info1, line, n, moved
    return errs, info1, line, n


@pytest.mark.parametrize("world", [2, 3, 4])
def test_resident_loop_on_ranks_equals_one_context(world):
    """The same loop in multi-rank mode (several contexts of this process as ranks, in-process transport): every rank
    integrates its own stretch of the Peano-Hilbert order, the TimeBinCount[] are all-gathered, the force path imports
    its locally-essential set in every step - also in steps where a rank has few or no active particles.  After eight
    steps on individual time bins the ranks hold what one context holds: time line exact, the floating-point state to
    the 1e-12 per evaluation of FP64 sums formed in another order."""
    errs, info1, line, n = ranks_vs_one(world)
    assert info1["steps_done"] == 8 and 0 < info1["active_total"] < 8 * n
    assert all(v == 0 for v in line.values()), line
    for k, tol in (("Pos", 1e-13), ("Vel", 1e-10), ("GravAccel", 1e-10), ("Hsml", 1e-10), ("VelPred", 1e-10), ("Entropy", 1e-10),
                   ("Density", 1e-10), ("HydroAccel", 1e-8), ("Pressure", 1e-10), ("MaxSignalVel", 1e-10)):
        assert errs[k].max() < tol, util.summarize(k, errs[k])


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref/ref_harness not built (needs /root/reference)")
def test_loop_on_the_reference_structs():
    """gl_gadget_run_steps: the same eight iterations, called the way a reference-side binding calls them - on P[] /
    SphP[] byte buffers laid out like the reference's structs (read once, written once).  The first kick of op L is
    taken with the host-side gl_gadget_advance_and_find_timesteps (bit-identical to the reference, tests/test_timestep.py);
    afterwards the structs must hold the reference's end state."""
    import ctypes as C
    import test_timestep as TT
    from gadget_leicester_b200 import aos, lib as L
    from gadget_leicester_b200.gadget import ForcePath
    par, b, e, loop = _reference_loop()
    n, ng = b["Pos"].shape[0], b["VelPred"].shape[0]
    z = {"pre_" + k: v for k, v in b.items()}
    a = TT._arrays(z, "pre")
    a.pview("p_OldAcc", np.float64)[:] = b["OldAcc"]
    a.pview("p_Ti_current", np.int32)[:] = 64
    a.sview("s_DhsmlDensityFactor", np.float64)[:] = b["DhsmlDensityFactor"]
    a.sview("s_DivVel", np.float64)[:] = b["DivVel"]
    a.sview("s_CurlVel", np.float64)[:] = b["CurlVel"]
    q = dict(par)
    q["ErrTolTheta"] = float(b["globals"][2])
    q["Ti_Current"] = 64
    fp = ForcePath(q)
    t = _tpar(fp, par, 64, 6)
    lay = aos.layout()
    t.p_DeltaDustMomentum, t.p_NewDensity = lay["ts_p_DeltaDustMomentum"], lay["io_p_NewDensity"]
    rc, msg = TT._call(a, t, np.arange(n, dtype=np.int32))
    assert rc == 0, msg
    lib = fp._lib
    ls = aos.layout_struct()
    fp._chk(lib.gl_gadget_bind(fp.ctx, C.byref(ls)))
    info = L.GLRunInfo()
    fp._chk(lib.gl_gadget_run_steps(fp.ctx, L._ptr(a.P), L._ptr(a.SphP), n, ng, C.byref(t), 8, C.byref(info)))
    assert info.steps_done == 8 and t.Ti_Current == 80
    ids = b["id"]

    def mine(x):
        out = np.zeros_like(x)
        out[ids[: x.shape[0]] - 1] = x
        return out

    gas = np.zeros(n, bool)
    gas[ids[:ng] - 1] = True
    assert np.array_equal(mine(a.pview("p_TimeBin", np.int16).astype(np.int32)), _by_id(e, "TimeBin", n))
    assert np.array_equal(mine(a.pview("p_Ti_begstep", np.int32).copy()), _by_id(e, "Ti_begstep", n))
    ty = b["Type"]
    assert np.all(a.pview("p_Ti_current", np.int32)[ty != 3] == 80)
    for key, f, c, tol in (("Pos", "p_Pos", 3, 1e-9), ("Vel", "p_Vel", 3, 1e-6), ("GravAccel", "p_GravAccel", 3, 1e-5), ("Hsml", "p_Hsml", 1, 1e-9)):
        err = util.rel_err_vec(mine(a.pview(f, np.float64, c).copy()), _by_id(e, key, n))
        assert np.percentile(err, 99) < tol and err.max() < 30 * tol, util.summarize(key, err)
    for key, f, c, tol in (("VelPred", "s_VelPred", 3, 1e-6), ("Entropy", "s_Entropy", 1, 1e-9), ("DtEntropy", "s_DtEntropy", 1, 1e-6),
                           ("Density", "s_Density", 1, 1e-9), ("Pressure", "s_Pressure", 1, 1e-9), ("HydroAccel", "s_HydroAccel", 3, 1e-6),
                           ("MaxSignalVel", "s_MaxSignalVel", 1, 1e-9)):
        ref = _by_id(e, key, n)[gas]
        got = mine(_pad(a.sview(f, np.float64, c).copy(), n))[gas]
        if ref.ndim == 1:
            err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-6 * np.abs(ref).max())
        else:
            err = util.rel_err_vec(got, ref)
        assert np.percentile(err, 99) < tol and err.max() < 100 * tol, util.summarize(key, err)
    # a particle that is not at Ti_Current is refused (the loop drifts everybody from ONE time)
    a.pview("p_Ti_current", np.int32)[5] = 79
    assert lib.gl_gadget_run_steps(fp.ctx, L._ptr(a.P), L._ptr(a.SphP), n, ng, C.byref(t), 1, C.byref(info)) != 0
    assert "not drifted" in lib.gl_last_error(fp.ctx).decode()
    fp.close()


REF_DROPIN = os.path.join(util.ROOT, "oracle", "_ref", "ref_dropin")


@pytest.mark.skipif(not (util.have_ref() and os.path.exists(REF_DROPIN)), reason="oracle/_ref binaries not built (needs /root/reference)")
def test_reference_harness_with_the_glue_resident_loop():
    """The drop-in proof for the loop: the reference harness and the reference's own objects, linked with
    integration/gl_gadget_glue.c (-DGL_RESIDENT_RUN).  Op M is op L with the eight iterations handed to
    gl_glue_run_steps() - P[] / SphP[] read once, the loop in the library on the GPU, the structs written once, the host's
    time line (All.Ti_Current, TimeBinActive[], the time-bin lists via the reference's reconstruct_timebins()) brought up
    to date.  The all-reference binary's op L is the checker."""
    mg = _golden_module()
    ps = mg._kick_set()
    par = ics.default_params(Ti_Current=16, MinEgySpec=1.8)
    par["Timebase_interval"] = TBI
    par["PartAllocFactor"] = 20.0
    ref = util.run_oracle(util.REF_HARNESS, ps, par, "T S D R G D U H L")
    got = util.run_oracle(REF_DROPIN, ps, par, "T S D R G D U H M")
    r = {k[len("8_L_end_"):]: v for k, v in ref.items() if k.startswith("8_L_end_")}
    g = {k[len("8_M_end_"):]: v for k, v in got.items() if k.startswith("8_M_end_")}
    assert set(r) == set(g) and len(r) > 25
    n = r["Pos"].shape[0]
    assert np.array_equal(ref["8_L_loop"], got["8_M_loop"])                  # Ti_Current 80, All.Time, everybody current, 8 steps
    for k in ("TimeBin", "Ti_begstep", "Type"):
        assert np.array_equal(_by_id(r, k, n), _by_id(g, k, n)), k
    for k in ("TimeBinCount", "TimeBinCountSph", "TimeBinActive"):           # the host's time line after reconstruct_timebins()
        assert np.array_equal(r[k], g[k]), k
    ng = r["VelPred"].shape[0]
    gas = np.zeros(n, bool)
    gas[r["id"][:ng] - 1] = True
    for k, tol in (("Pos", 1e-9), ("Vel", 1e-6), ("GravAccel", 1e-5), ("Hsml", 1e-9), ("Mass", 1e-15)):
        err = util.rel_err_vec(_by_id(g, k, n), _by_id(r, k, n))
        assert np.percentile(err, 99) < tol and err.max() < 30 * tol, util.summarize(k, err)
    # (here the state the loop STARTS from is the CUDA path's as well - ops G D U H of the drop-in binary - so the
    #  velocities carry the 1e-7 of its FP32 force magnitudes into every SPH quantity that reads them)
    for k, tol in (("VelPred", 1e-6), ("Entropy", 1e-8), ("DtEntropy", 1e-6), ("Density", 1e-8), ("Pressure", 1e-8), ("HydroAccel", 1e-6),
                   ("MaxSignalVel", 1e-7), ("DivVel", 1e-5), ("CurlVel", 1e-5), ("DhsmlDensityFactor", 1e-8)):
        a, b_ = _by_id(g, k, n)[gas], _by_id(r, k, n)[gas]
        if a.ndim == 1:
            err = np.abs(a - b_) / np.maximum(np.abs(b_), 1e-6 * np.abs(b_).max())
        else:
            err = util.rel_err_vec(a, b_)
        assert np.percentile(err, 99) < tol and err.max() < 100 * tol, util.summarize(k, err)


@pytest.mark.parametrize("by_cost", [False, True])
def test_domain_exchange_of_the_resident_state(by_cost):
    """gl_domain_exchange (domain_exchange, domain.c:665, for the state in HBM): three ranks start from a deliberately
    bad assignment (rank 0 holds half of all particles), run three steps, exchange - by equal counts, or by the reference's
    cost (1 + GravCost) / 2^TimeBin - and run five more.  Every particle ends on exactly one rank, the loads are balanced,
    and - because results do not depend on who owns what - the eight steps still end in the state of one context: time
    line exact, gravity bitwise, the rest to FP64 summation order."""
    errs, info1, line, n, moved = ranks_vs_one(3, nsteps=8, exchange_after=3, skew=True, by_cost=by_cost)
    assert all(v == 0 for v in line.values()), line
    for k, tol in (("Pos", 1e-13), ("Vel", 1e-10), ("GravAccel", 1e-10), ("Hsml", 1e-10), ("VelPred", 1e-10), ("Entropy", 1e-10),
                   ("Density", 1e-10), ("HydroAccel", 1e-8), ("Pressure", 1e-10), ("MaxSignalVel", 1e-10)):
        assert errs[k].max() < tol, util.summarize(k, errs[k])
    held = [moved[r][0] for r in range(3)]
    sent = [moved[r][1] for r in range(3)]
    before = [moved[r][2] for r in range(3)]
    assert sum(held) == n and before[0] == n // 2 and sum(sent) > 0
    if not by_cost:
        assert max(abs(h - n / 3) for h in held) < 0.02 * n, held      # equal counts (within the 2048-sample resolution)
        assert sent[0] > 0.15 * n                                        # rank 0 gave a third of its half away
