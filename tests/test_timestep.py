"""CPU suite: gl_gadget_advance_and_find_timesteps (host code of libgl_b200.so, no GPU involved) against the
reference's OWN advance_and_find_timesteps (timestep.c:29 -> get_timestep :607, do_the_kick :364).  The fixture
tests/golden/aux/kick_mixed_1540.npz was written by oracle/_ref/ref_harness op V (tests/golden/make_golden.py): 1500
gas particles + 40 particles of types 1, 2, 4, 5 on individual time bins at a synchronisation point, after a settled
force evaluation; the state the function reads and the state the unmodified reference left behind, including its
time-bin lists.  Same statements, same libm: compared bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

import util
from gadget_leicester_b200 import aos
from gadget_leicester_b200 import lib as L

FIX = os.path.join(util.ROOT, "tests", "golden", "aux", "kick_mixed_1540.npz")
TIMEBINS = 29

P_FIELDS = (("Pos", "p_Pos", np.float64, 3), ("Vel", "p_Vel", np.float64, 3), ("GravAccel", "p_GravAccel", np.float64, 3),
            ("Mass", "p_Mass", np.float64, 1), ("Hsml", "p_Hsml", np.float64, 1), ("Type", "p_Type", np.int16, 1),
            ("TimeBin", "p_TimeBin", np.int16, 1), ("Ti_begstep", "p_Ti_begstep", np.int32, 1),
            ("DeltaDustMomentum", "ts_p_DeltaDustMomentum", np.float64, 3), ("NewDensity", "io_p_NewDensity", np.float64, 1))
S_FIELDS = (("VelPred", "s_VelPred", np.float64, 3), ("HydroAccel", "s_HydroAccel", np.float64, 3), ("DragAccel", "s_DragAccel", np.float64, 3),
            ("Density", "s_Density", np.float64, 1), ("Entropy", "s_Entropy", np.float64, 1), ("DtEntropy", "s_DtEntropy", np.float64, 1),
            ("Pressure", "s_Pressure", np.float64, 1), ("MaxSignalVel", "s_MaxSignalVel", np.float64, 1))


def _arrays(z, tag="pre"):
    n = z[tag + "_Pos"].shape[0]
    ps = dict(pos=z[tag + "_Pos"], vel=z[tag + "_Vel"], mass=z[tag + "_Mass"], id=z[tag + "_id"], hsml=z[tag + "_Hsml"],
              oldacc=np.zeros(n), type=z[tag + "_Type"], timebin=z[tag + "_TimeBin"], entropy=z[tag + "_Entropy"],
              dtentropy=z[tag + "_DtEntropy"])
    a = aos.GadgetArrays(ps, pinned=False)
    for name, f, dt, c in P_FIELDS:
        a.pview(f, dt, c)[:] = z[tag + "_" + name]
    for name, f, dt, c in S_FIELDS:
        a.sview(f, dt, c)[:] = z[tag + "_" + name]
    return a


def _params(z):
    q = z["params"]
    lay = aos.layout()
    t = L.GLTimestepParams()
    t.Ti_Current = int(q[0])
    for b in range(TIMEBINS):
        t.TimeBinActive[b] = int(z["pre_TimeBinActive"][b])
    (t.Timebase_interval, t.ErrTolIntAccuracy, t.CourantFac, t.MaxSizeTimestep, t.MinSizeTimestep, t.MinEgySpec) = [float(x) for x in q[1:7]]
    t.GammaMinus1 = 7. / 5. - 1
    for k in range(6):
        t.SofteningTable[k] = float(z["SofteningTable"][k])
    (t.OuterBoundary, t.UnitVelocity_in_cm_per_s, t.FeedBackVelocity, t.SMBHmass, t.InnerBoundary, t.SinkBoundary,
     t.AccDtBlackHole) = [float(x) for x in q[7:14]]
    t.p_DeltaDustMomentum, t.p_NewDensity = lay["ts_p_DeltaDustMomentum"], lay["io_p_NewDensity"]
    return t


def _call(a, t, active, binold=None):
    lib = L.load_library()
    lay = aos.layout_struct()
    err = C.create_string_buffer(512)
    act = np.ascontiguousarray(active, dtype=np.int32)
    rc = lib.gl_gadget_advance_and_find_timesteps(C.byref(lay), C.byref(t), L._ptr(a.P), L._ptr(a.SphP), L._ptr(act), act.size,
                                                  L._ptr(binold), err, len(err))
    return rc, err.value.decode()


def _relink(z, binold, binnew, active):
    """the serial bookkeeping the caller keeps (timestep.c:170-233), restated on the reference's list arrays"""
    first, last = z["pre_FirstInTimeBin"].copy(), z["pre_LastInTimeBin"].copy()
    nxt, prv = z["pre_NextInTimeBin"].copy(), z["pre_PrevInTimeBin"].copy()
    cnt, cnts = z["pre_TimeBinCount"].copy(), z["pre_TimeBinCountSph"].copy()
    types = z["pre_Type"]
    for k, i in enumerate(active):
        bo, bn = int(binold[k]), int(binnew[i])
        if bo == bn:
            continue
        cnt[bo] -= 1
        if types[i] == 0:
            cnts[bo] -= 1
        p, q = prv[i], nxt[i]
        if first[bo] == i:
            first[bo] = q
        if last[bo] == i:
            last[bo] = p
        if p >= 0:
            nxt[p] = q
        if q >= 0:
            prv[q] = p
        if cnt[bn] > 0:
            prv[i] = last[bn]
            nxt[last[bn]] = i
            nxt[i] = -1
            last[bn] = i
        else:
            first[bn] = last[bn] = i
            prv[i] = nxt[i] = -1
        cnt[bn] += 1
        if types[i] == 0:
            cnts[bn] += 1
    return first, last, nxt, prv, cnt, cnts


def test_library_kick_equals_reference_advance_and_find_timesteps():
    z = np.load(FIX)
    a, t = _arrays(z), _params(z)
    active = z["active"]
    assert active.size == a.NumPart == int(z["params"][14])
    binold = np.full(active.size, -9, np.int32)
    rc, msg = _call(a, t, active, binold)
    assert rc == 0, msg
    assert np.array_equal(binold, z["pre_TimeBin"][active])
    for name, f, dt, c in P_FIELDS:
        assert (a.pview(f, dt, c) == z["post_" + name]).all(), name
    for name, f, dt, c in S_FIELDS:
        assert (a.sview(f, dt, c) == z["post_" + name]).all(), name
    # the fixture is not a trivial one: steps shrink and grow, some want more than the synchronised bins allow, both
    # entropy limiters, the energy floor and the dust / sink rules all fire
    old, new = z["pre_TimeBin"].astype(int), z["post_TimeBin"].astype(int)
    assert (new < old).sum() > 100 and (new > old).sum() > 100 and new.max() == 6
    gas = z["pre_Type"] == 0
    assert (z["post_DtEntropy"] == 0).sum() > 10                              # floored at MinEgySpec
    assert (z["post_Entropy"] == 0.5 * z["pre_Entropy"]).sum() > 0 or (z["post_DtEntropy"] != z["pre_DtEntropy"]).sum() > 50
    dust = z["pre_Type"] == 2
    assert np.array_equal(z["post_Vel"][dust], z["pre_Vel"][dust]) and not np.array_equal(z["post_Vel"][gas], z["pre_Vel"][gas])
    assert not z["post_DragAccel"].any() and z["pre_DragAccel"].any()
    # and the caller's share: relinking the changed particles in active-list order gives the reference's lists
    first, last, nxt, prv, cnt, cnts = _relink(z, binold, a.pview("p_TimeBin", np.int16), active)
    for got, name in ((first, "FirstInTimeBin"), (last, "LastInTimeBin"), (cnt, "TimeBinCount"), (cnts, "TimeBinCountSph"),
                      (nxt, "NextInTimeBin"), (prv, "PrevInTimeBin")):
        want = z["post_" + name]
        if name in ("FirstInTimeBin", "LastInTimeBin"):
            used = z["post_TimeBinCount"] > 0                     # heads of empty bins are stale in the reference too
            assert np.array_equal(got[used], want[used]), name
        else:
            assert np.array_equal(got, want), name


def _tiled(z, reps):
    """the fixture `reps` times over (gas blocks first, as P[] wants them): enough particles for several work items"""
    ng = int((z["pre_Type"] == 0).sum())
    out = {}
    for k in z.files:
        v = z[k]
        if k.startswith(("pre_", "post_")) and v.shape[0] == z["pre_Type"].shape[0]:
            out[k] = np.concatenate([np.concatenate([v[:ng]] * reps), np.concatenate([v[ng:]] * reps)])
        elif k.startswith(("pre_", "post_")) and v.shape[0] == ng:
            out[k] = np.concatenate([v] * reps)
        else:
            out[k] = v
    return out


def test_kick_over_many_work_items():
    z = np.load(FIX)
    zt = _tiled(z, 12)                                      # 18 480 particles: five work items of 4 096
    a, t = _arrays(zt), _params(z)
    rc, msg = _call(a, t, np.arange(a.NumPart, dtype=np.int32)[::-1])          # any order of the active list will do
    assert rc == 0, msg
    for name, f, dt, c in P_FIELDS:
        assert (a.pview(f, dt, c) == zt["post_" + name]).all(), name
    for name, f, dt, c in S_FIELDS:
        assert (a.sview(f, dt, c) == zt["post_" + name]).all(), name


def test_subset_of_active_particles_leaves_the_others_alone():
    z = np.load(FIX)
    a, t = _arrays(z), _params(z)
    lay = aos.layout()
    before_P = np.array(a.P, copy=True).reshape(a.NumPart, lay["p_stride"])
    before_S = np.array(a.SphP, copy=True)[: a.N_gas * lay["sph_stride"]].reshape(a.N_gas, lay["sph_stride"])
    active = z["active"][::3]
    rc, msg = _call(a, t, active)
    assert rc == 0, msg
    idle = np.ones(a.NumPart, bool)
    idle[active] = False
    after_P = np.asarray(a.P).reshape(a.NumPart, lay["p_stride"])
    after_S = np.asarray(a.SphP)[: a.N_gas * lay["sph_stride"]].reshape(a.N_gas, lay["sph_stride"])
    assert (after_P[idle] == before_P[idle]).all() and (after_S[idle[: a.N_gas]] == before_S[idle[: a.N_gas]]).all()
    for name, f, dt, c in (("Vel", "p_Vel", np.float64, 3), ("TimeBin", "p_TimeBin", np.int16, 1), ("Ti_begstep", "p_Ti_begstep", np.int32, 1)):
        assert (a.pview(f, dt, c)[active] == z["post_" + name][active]).all(), name


def test_failures_where_the_reference_would_endrun():
    z = np.load(FIX)
    active = z["active"]
    a, t = _arrays(z), _params(z)
    t.MinSizeTimestep = 1.0                      # endrun(888): step below the limit
    rc, msg = _call(a, t, active)
    assert rc != 0 and "MinSizeTimestep" in msg
    a, t = _arrays(z), _params(z)
    t.Timebase_interval = 1.0                    # endrun(818): integer step of size zero
    rc, msg = _call(a, t, active)
    assert rc != 0 and "size zero" in msg
    a, t = _arrays(z), _params(z)
    a.pview("p_Type", np.int16)[-1] = 3          # a virtual particle whose criterion is undefined in the reference
    a.pview("io_p_NewDensity", np.float64)[-1] = 1.0
    rc, msg = _call(a, t, active)
    assert rc != 0 and "uninitialised" in msg


REF_HOST_DROPIN = os.path.join(util.ROOT, "oracle", "_ref", "ref_host_dropin")


@pytest.mark.skipif(not os.path.exists(REF_HOST_DROPIN), reason="oracle/_ref/ref_host_dropin not built (needs /root/reference)")
def test_reference_host_with_the_glue_advance_and_find_timesteps():
    """The reference harness (its own CPU force path) linked with advance_and_find_timesteps() AND peano_hilbert_order()
    from integration/gl_gadget_glue.c (-DGL_TIMESTEP -DGL_PEANO_ORDER): every array of the committed all-reference
    fixture comes out the same - the particle order domain_Decomposition() leaves behind (the dumps are in P[] order,
    ids included) and the time-bin lists as well."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(util.ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    got = mg.kick_run(REF_HOST_DROPIN)
    z = np.load(FIX)
    assert set(got) == set(z.files)
    used = z["post_TimeBinCount"] > 0
    for k in z.files:
        if k in ("post_FirstInTimeBin", "post_LastInTimeBin"):
            assert np.array_equal(got[k][used], z[k][used]), k
        else:
            assert np.array_equal(got[k], z[k]), k


@pytest.mark.skipif(not (util.have_ref() and os.path.exists(REF_HOST_DROPIN)), reason="oracle/_ref binaries not built (needs /root/reference)")
def test_integration_loop_with_the_glue_host_pieces_is_bit_identical():
    """Eight iterations of the reference's run() loop (harness op L: find_next_sync_point_and_drift -> domain_Decomposition
    -> gravity / density / hydro for the ACTIVE particles -> advance_and_find_timesteps, individual time steps, a new
    decomposition and tree in every step) through the all-reference harness and through the one linked with the glue's
    domain_Decomposition(), peano_hilbert_order() and advance_and_find_timesteps(): the end state - positions,
    velocities, entropies, bins, time-bin lists, in P[] order - is the same bit for bit."""
    import importlib.util
    from gadget_leicester_b200 import ics
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(util.ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    ps = mg._kick_set()
    par = ics.default_params(Ti_Current=16, MinEgySpec=1.8)
    par["Timebase_interval"] = 4.0e-5
    par["PartAllocFactor"] = 20.0
    out = {}
    for name, binary in (("ref", util.REF_HARNESS), ("glue", REF_HOST_DROPIN)):
        o = util.run_oracle(binary, ps, par, "T S D R G D U H L")
        out[name] = {k[len("8_L_"):]: v for k, v in o.items() if k.startswith("8_L_")}
    ref, glue = out["ref"], out["glue"]
    assert ref["loop"][3] == 8 and ref["loop"][0] == 80                     # eight sync points: 66 ... 80
    assert set(ref) == set(glue) and len(ref) > 25
    for k in ref:
        assert np.array_equal(ref[k], glue[k]), k
    # the loop did something: particles moved, several bins were active along the way
    z = np.load(FIX)
    assert not np.array_equal(np.sort(ref["end_Pos"], axis=0), np.sort(z["pre_Pos"], axis=0))
