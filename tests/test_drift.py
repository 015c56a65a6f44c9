"""CPU suite: gl_gadget_drift (host code of libgl_b200.so, no GPU involved) against the reference's OWN drift_particle
(predict.c:129).  The fixture tests/golden/aux/drift_cube_timebins_2k.npz was written by oracle/_ref/ref_harness op X
(tests/golden/make_golden.py): the state the drift reads, and the state the unmodified reference left behind, for
2000 gas particles on individual time bins after a settled force evaluation; every third particle is already current
and must not move.  Same statements, same libm: the results are compared bit for bit."""
import ctypes as C
import os

import numpy as np

import util
from gadget_leicester_b200 import aos
from gadget_leicester_b200 import lib as L

FIX = os.path.join(util.ROOT, "tests", "golden", "aux", "drift_cube_timebins_2k.npz")

P_FIELDS = (("Pos", "p_Pos", np.float64, 3), ("Vel", "p_Vel", np.float64, 3), ("GravAccel", "p_GravAccel", np.float64, 3),
            ("Mass", "p_Mass", np.float64, 1), ("Hsml", "p_Hsml", np.float64, 1), ("Type", "p_Type", np.int16, 1),
            ("TimeBin", "p_TimeBin", np.int16, 1), ("Ti_begstep", "p_Ti_begstep", np.int32, 1), ("Ti_current", "p_Ti_current", np.int32, 1))
S_FIELDS = (("VelPred", "s_VelPred", np.float64, 3), ("HydroAccel", "s_HydroAccel", np.float64, 3), ("DragAccel", "s_DragAccel", np.float64, 3),
            ("DivVel", "s_DivVel", np.float64, 1), ("Density", "s_Density", np.float64, 1), ("Entropy", "s_Entropy", np.float64, 1),
            ("DtEntropy", "s_DtEntropy", np.float64, 1), ("Pressure", "s_Pressure", np.float64, 1))


def _arrays(z, tag):
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


def test_library_drift_equals_reference_drift_particle():
    z = np.load(FIX)
    t0, t1, tbi = z["times"]
    a = _arrays(z, "pre")
    before = np.array(a.P, copy=True)
    lib = L.load_library()
    lay = aos.layout_struct()
    nd = C.c_longlong(0)
    rc = lib.gl_gadget_drift(C.byref(lay), L._ptr(a.P), L._ptr(a.SphP), a.NumPart, a.N_gas, int(t1), float(tbi), float(z["mingashsml"][0]),
                             C.byref(nd))
    assert rc == 0
    stale = z["pre_Ti_current"] != int(t1)
    assert nd.value == int(stale.sum()) and 0 < nd.value < a.NumPart
    for name, f, dt, c in P_FIELDS:
        assert (a.pview(f, dt, c) == z["post_" + name]).all(), name
    for name, f, dt, c in S_FIELDS:
        assert (a.sview(f, dt, c) == z["post_" + name]).all(), name
    # particles that were already current did not change at all
    lay_d = aos.layout()
    Pb = np.frombuffer(before, np.uint8).reshape(a.NumPart, lay_d["p_stride"])
    Pa = np.frombuffer(np.array(a.P, copy=True), np.uint8).reshape(a.NumPart, lay_d["p_stride"])
    assert (Pb[~stale] == Pa[~stale]).all()
    # a particle ahead of the target time is the reference's endrun(12)
    assert lib.gl_gadget_drift(C.byref(lay), L._ptr(a.P), L._ptr(a.SphP), a.NumPart, a.N_gas, int(t1) - 1, float(tbi), 0.0, None) != 0


def test_drift_over_many_work_items():
    """the same fixture twelve times over: 24 000 particles, several work items of the host pool"""
    z0 = np.load(FIX)
    n0 = z0["pre_Pos"].shape[0]
    z = {k: (np.concatenate([z0[k]] * 12) if (k.startswith(("pre_", "post_")) and z0[k].shape[0] == n0) else z0[k]) for k in z0.files}
    t0, t1, tbi = z["times"]
    a = _arrays(z, "pre")
    lib = L.load_library()
    lay = aos.layout_struct()
    nd = C.c_longlong(0)
    assert lib.gl_gadget_drift(C.byref(lay), L._ptr(a.P), L._ptr(a.SphP), a.NumPart, a.N_gas, int(t1), float(tbi),
                               float(z["mingashsml"][0]), C.byref(nd)) == 0
    assert nd.value == int((z["pre_Ti_current"] != int(t1)).sum())
    for name, f, dt, c in P_FIELDS:
        assert (a.pview(f, dt, c) == z["post_" + name]).all(), name
    for name, f, dt, c in S_FIELDS:
        assert (a.sview(f, dt, c) == z["post_" + name]).all(), name


def test_reorder_permutes_structs_like_the_reference():
    """gl_gadget_reorder = reorder_gas + reorder_particles (peano.c:106, :148): new P[k] = old P[order[k]], SphP[] with
    the gas block, every byte of every struct."""
    rng = np.random.default_rng(0)
    lay_d = aos.layout()
    n, ngas = 1003, 800
    ps_, ss_ = lay_d["p_stride"], lay_d["sph_stride"]
    Pb = rng.integers(0, 256, size=n * ps_, dtype=np.uint8)
    Sb = rng.integers(0, 256, size=ngas * ss_, dtype=np.uint8)
    order = np.concatenate([rng.permutation(ngas), ngas + rng.permutation(n - ngas)]).astype(np.int32)
    want_P = Pb.reshape(n, ps_)[order].copy()
    want_S = Sb.reshape(ngas, ss_)[order[:ngas]].copy()
    lib = L.load_library()
    lay = aos.layout_struct()
    assert lib.gl_gadget_reorder(C.byref(lay), L._ptr(Pb), L._ptr(Sb), n, ngas, L._ptr(order)) == 0
    assert (Pb.reshape(n, ps_) == want_P).all() and (Sb.reshape(ngas, ss_) == want_S).all()
    bad = order.copy()
    bad[0], bad[-1] = bad[-1], bad[0]          # a non-gas particle into the gas block: refused, nothing touched
    assert lib.gl_gadget_reorder(C.byref(lay), L._ptr(Pb), L._ptr(Sb), n, ngas, L._ptr(bad)) != 0
    assert (Pb.reshape(n, ps_) == want_P).all()


def test_host_peano_hilbert_order_is_the_reference_stable_sort():
    """gl_gadget_peano_hilbert_order = peano_hilbert_order() (peano.c:25-92) from given keys: gas block and the rest sorted
    separately, equal keys keep their order (the reference's merge takes the left element, peano.c:551), structs
    permuted.  Many ties on purpose; sizes around the parallel sort's stretch boundaries."""
    lib = L.load_library()
    lay = aos.layout_struct()
    lay_d = aos.layout()
    ps_, ss_ = lay_d["p_stride"], lay_d["sph_stride"]
    for n, ngas, seed in ((1, 1, 0), (5000, 0, 1), (70001, 50000, 2), (20000, 20000, 3)):
        rng = np.random.default_rng(seed)
        Pb = rng.integers(0, 256, size=n * ps_, dtype=np.uint8)
        Sb = rng.integers(0, 256, size=max(ngas, 1) * ss_, dtype=np.uint8)
        keys = rng.integers(0, max(2, n // 7), size=n).astype(np.uint64) << np.uint64(40)
        want = np.concatenate([np.argsort(keys[:ngas], kind="stable"), ngas + np.argsort(keys[ngas:], kind="stable")]).astype(np.int32)
        want_P = Pb.reshape(n, ps_)[want].copy()
        want_S = Sb[: ngas * ss_].reshape(ngas, ss_)[want[:ngas]].copy()
        order = np.full(n, -1, np.int32)
        assert lib.gl_gadget_peano_hilbert_order(C.byref(lay), L._ptr(Pb), L._ptr(Sb), n, ngas, L._ptr(keys), L._ptr(order)) == 0
        assert np.array_equal(order, want)
        assert (Pb.reshape(n, ps_) == want_P).all() and (Sb[: ngas * ss_].reshape(ngas, ss_) == want_S).all()


def test_drift_of_a_list_touches_only_the_listed_particles():
    """gl_gadget_drift_list = the drift loop of find_next_sync_point_and_drift (run.c:325-329): the listed particles end
    up exactly where the reference's drift_particle puts them, everything else keeps every byte."""
    z = np.load(FIX)
    t0, t1, tbi = z["times"]
    a = _arrays(z, "pre")
    lay_d = aos.layout()
    before_P = np.array(a.P, copy=True).reshape(a.NumPart, lay_d["p_stride"])
    before_S = np.array(a.SphP, copy=True).reshape(-1, lay_d["sph_stride"])[: a.N_gas]
    stale = np.flatnonzero(z["pre_Ti_current"] != int(t1))
    chosen = np.ascontiguousarray(stale[::2][::-1], dtype=np.int32)          # every other stale particle, in reverse order
    lib = L.load_library()
    lay = aos.layout_struct()
    nd = C.c_longlong(0)
    rc = lib.gl_gadget_drift_list(C.byref(lay), L._ptr(a.P), L._ptr(a.SphP), L._ptr(chosen), chosen.size, a.N_gas, int(t1), float(tbi),
                                  float(z["mingashsml"][0]), C.byref(nd))
    assert rc == 0 and nd.value == chosen.size
    for name, f, dt, c in P_FIELDS:
        assert (a.pview(f, dt, c)[chosen] == z["post_" + name][chosen]).all(), name
    for name, f, dt, c in S_FIELDS:
        assert (a.sview(f, dt, c)[chosen] == z["post_" + name][chosen]).all(), name
    rest = np.ones(a.NumPart, bool)
    rest[chosen] = False
    after_P = np.asarray(a.P).reshape(a.NumPart, lay_d["p_stride"])
    after_S = np.asarray(a.SphP).reshape(-1, lay_d["sph_stride"])[: a.N_gas]
    assert (after_P[rest] == before_P[rest]).all() and (after_S[rest[: a.N_gas]] == before_S[rest[: a.N_gas]]).all()
