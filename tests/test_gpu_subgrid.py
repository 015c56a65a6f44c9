"""The sub-grid neighbour loops behind the force path against the compiled reference:
blackhole_accretion (blackhole.c:70), dust_density (dust.c:60), dust_drag (dust.c:263).

The reference (oracle/_ref/ref_harness ops B / Y / Z) dumps everything each routine reads ("pre") and everything it
leaves behind ("post") in its own P[] order; the CUDA path is given exactly the pre state in that order and has to arrive
at the post state: marks and counts (SwallowID, which particles lose their mass, how many of each type) bit-exact,
floating-point fields within 1e-9 relative (they are FP64 sums in another order; north_star's bar is 1e-5)."""
import numpy as np
import pytest

import util
from gadget_leicester_b200 import ics

pytestmark = pytest.mark.gpu

OPS = "T S D R G G D U H B Y Z"
RTOL = 1e-9


def _stage(out, k, op):
    pre = "%d_%s_" % (k, op)
    d = {}
    for key, v in out.items():
        if key.startswith(pre):
            d[key[len(pre):]] = v
    return d


def _split(st):
    a = {k[4:]: v for k, v in st.items() if k.startswith("pre_")}
    b = {k[5:]: v for k, v in st.items() if k.startswith("post_")}
    return a, b, st["params"]


def _pad(a, n):
    out = np.zeros((n,) + a.shape[1:], dtype=a.dtype)
    out[: a.shape[0]] = a
    return out


def _forcepath(pre, q, par):
    """a ForcePath holding the reference's pre state, in the reference's P[] order, with the tree built"""
    from gadget_leicester_b200.gadget import ForcePath
    n = pre["Pos"].shape[0]
    ps = dict(pos=pre["Pos"].copy(), vel=pre["Vel"].copy(), mass=pre["Mass"].copy(), type=pre["Type"].copy(), hsml=pre["Hsml"].copy(),
              id=pre["id"].copy(), timebin=pre["TimeBin"].copy(), oldacc=np.zeros(n), entropy=_pad(pre["Entropy"], n), dtentropy=np.zeros(n))
    fp = ForcePath(par)
    fp.load(ps)
    fp.force_treebuild()
    fp.set_subgrid_params(SMBHmass=q[0], InnerBoundary=q[1], SinkBoundary=q[2], SofteningBndry=q[3], ViscousTimescale=q[4], UnitTime_in_s=q[5],
                          UnitMass_in_g=q[6], UnitEnergy_in_cgs=q[7], BlackHoleFeedbackFactor=q[8], UnitLength_in_cm=q[10],
                          UnitDensity_in_cgs=q[11], MeanWeight=q[12], MinEgySpec=q[13], UnitVelocity_in_cm_per_s=q[14])
    return fp


def _close(a, b, what, rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.abs(b).max() if b.size else 0.0
    err = np.abs(a - b) / np.where(np.abs(b) > 1e-300, np.abs(b), 1.0)
    # entries that are tiny next to the field's scale are compared absolutely (sums that cancel)
    bad = (err > rtol) & (np.abs(a - b) > rtol * scale)
    assert not bad.any(), "%s: %d entries differ, worst rel %.3e" % (what, int(bad.sum()), float(err[bad].max()))


@pytest.fixture(scope="module")
def ref_run():
    if not util.have_ref():
        pytest.skip("oracle/_ref/ref_harness not built")
    par = ics.default_params(BlackHoleFeedbackFactor=1e-3, SinkBoundary=0.02, MinEgySpec=1e-6)
    ps = ics.disc_subgrid(20000, 2000, seed=7)
    ps["timebin"][:] = 3 + (ps["id"] % 3)
    out = util.run_oracle(util.REF_HARNESS, ps, par, OPS)
    return par, out


def test_blackhole_accretion(ref_run):
    par, out = ref_run
    pre, post, q = _split(_stage(out, OPS.split().index("B"), "B"))
    n = pre["Pos"].shape[0]
    fp = _forcepath(pre, q, par)
    S = dict(Vel=pre["Vel"].copy(), ID=pre["id"].copy(), SwallowID=pre["SwallowID"].copy(), BH_Mass=pre["BH_Mass"].copy(),
             BH_Mdot=pre["BH_Mdot"].copy(), AccDisc_Mass=pre["AccDisc_Mass"].copy(), Dust_Mass=pre["Dust_Mass"].copy(),
             Total_Mass=pre["Total_Mass"].copy(), Injected_BH_Energy=_pad(pre["Injected_BH_Energy"], n), SinkDensity=pre["SinkDensity"].copy())
    mass, counts = fp.blackhole_accretion(S, active=pre["active"])
    fp.close()
    gone_ref = (pre["Mass"] > 0) & (post["Mass"] == 0)
    t = pre["Type"]
    want = tuple(int((gone_ref & (t == k)).sum()) for k in (0, 5, 2))
    assert sum(want) > 100 and min(want) > 0, "the test set must make every branch swallow something: %r" % (want,)
    assert counts == want
    assert np.array_equal(S["SwallowID"], post["SwallowID"])
    assert np.array_equal(mass == 0, post["Mass"] == 0)
    _close(mass, post["Mass"], "Mass")
    _close(S["Vel"], post["Vel"], "Vel")
    for k in ("BH_Mass", "BH_Mdot", "AccDisc_Mass", "Dust_Mass", "Total_Mass"):
        _close(S[k][t == 5], post[k][t == 5], k)
    ng = post["Injected_BH_Energy"].shape[0]
    assert (post["Injected_BH_Energy"] != 0).sum() > 50
    _close(S["Injected_BH_Energy"][:ng], post["Injected_BH_Energy"], "Injected_BH_Energy")


def test_dust_density(ref_run):
    par, out = ref_run
    pre, post, q = _split(_stage(out, OPS.split().index("Y"), "Y"))
    fp = _forcepath(pre, q, par)
    d7 = pre["DustParticleDensity"].copy()
    fp.dust_density(d7, active=pre["active"])
    fp.close()
    assert (post["DustParticleDensity"] > 0).sum() > 1000
    _close(d7, post["DustParticleDensity"], "DUST_particle_density")


def test_dust_drag(ref_run):
    par, out = ref_run
    pre, post, q = _split(_stage(out, OPS.split().index("Z"), "Z"))
    n = pre["Pos"].shape[0]
    fp = _forcepath(pre, q, par)
    S = dict(Vel=pre["Vel"].copy(), GravAccel=pre["GravAccel"].copy(), GasDensity=pre["SinkDensity"].copy(), GasEntropy=pre["SinkEntropy"].copy(),
             GasVel=pre["SinkGasVel"].copy(), DustParticleDensity=pre["DustParticleDensity"].copy(), DustRadius=pre["DustRadius"].copy(),
             DustVcoll=pre["DustVcoll"].copy(), DeltaDustMomentum=pre["DeltaDustMomentum"].copy(), NewDragAcc=pre["NewDragAcc"].copy(),
             DeltaDragEnergy=pre["DeltaDragEnergy"].copy(), Entropy=_pad(pre["Entropy"], n), DragHeating=_pad(pre["DragHeating"], n))
    fp.dust_drag(S, active=pre["active"])
    fp.close()
    ng = post["Entropy"].shape[0]
    assert (post["Entropy"] != pre["Entropy"]).sum() > 10000 and (post["DragHeating"] != 0).sum() > 10000
    _close(S["Vel"], post["Vel"], "Vel")
    for k in ("DustVcoll", "DeltaDustMomentum", "NewDragAcc", "DeltaDragEnergy"):
        _close(S[k], post[k], k)
    _close(S["Entropy"][:ng], post["Entropy"], "Entropy")
    _close(S["DragHeating"][:ng], post["DragHeating"], "DragHeating")


REF_DROPIN_SUBGRID = util.ROOT + "/oracle/_ref/ref_dropin_subgrid"


def test_reference_host_with_subgrid_dropin_symbols(ref_run):
    """The reference's own harness linked against integration/gl_gadget_glue.c -DGL_SUBGRID: gravity_tree / density /
    hydro_force AND blackhole_accretion / dust_density / dust_drag are the CUDA path working on the reference's P[] /
    SphP[] (only the particles the loops touch are gathered and scattered), everything else is the unmodified reference.
    Same op sequence as the all-CPU binary, same answers (the inputs of the sub-grid loops differ by the round-off of the
    force path in front of them, hence 1e-6 instead of 1e-9)."""
    import os
    if not os.path.exists(REF_DROPIN_SUBGRID):
        pytest.skip("oracle/_ref/ref_dropin_subgrid did not travel")
    par, cpu = ref_run
    ps = ics.disc_subgrid(20000, 2000, seed=7)
    ps["timebin"][:] = 3 + (ps["id"] % 3)
    gpu = util.run_oracle(REF_DROPIN_SUBGRID, ps, par, OPS)
    n = ps["pos"].shape[0]

    def both(op, field, cnt=n):
        k = OPS.split().index(op)
        pre = "%d_%s_post_" % (k, op)
        return util.by_id(gpu, pre, field, cnt), util.by_id(cpu, pre, field, cnt)
    a, b = both("B", "SwallowID")
    assert np.array_equal(a, b)
    a, b = both("B", "Mass")
    assert np.array_equal(a == 0, b == 0)
    _close(a, b, "Mass", 1e-6)
    for k in ("Vel", "BH_Mass", "BH_Mdot", "AccDisc_Mass", "Dust_Mass", "Total_Mass"):
        _close(*both("B", k), k, 1e-6)
    ng = 20000
    _close(*both("B", "Injected_BH_Energy", ng), "Injected_BH_Energy", 1e-6)
    _close(*both("Y", "DustParticleDensity"), "DUST_particle_density", 1e-6)
    for k in ("Vel", "DustVcoll", "DeltaDustMomentum", "DeltaDragEnergy"):
        _close(*both("Z", k), k, 1e-6)
    _close(*both("Z", "Entropy", ng), "Entropy", 1e-6)
    _close(*both("Z", "DragHeating", ng), "DragHeating", 1e-6)


def test_sinks_competing_for_the_same_particles():
    """Two sinks of nearly equal mass inside each other's accretion boundary, dust between them, the active list in an
    order that is NOT the index order: who swallows what depends on the order the reference visits the sinks in
    (blackhole.c:991, :1030 - the last claim wins; :508 - a sink that is itself marked swallows nothing)."""
    if not util.have_ref():
        pytest.skip("oracle/_ref/ref_harness not built")
    par = ics.default_params(BlackHoleFeedbackFactor=2e-3, SinkBoundary=0.03, SMBHmass=10.0, MinEgySpec=1e-6,
                             Softening=[1.0e-5, 0.01, 5.0e-5, 1.0e-7, 0.01, 0.03])
    ps = ics.disc_subgrid(8000, 1500, seed=11)
    t = ps["type"]
    s = np.nonzero(t == 5)[0]
    # every sink is planet-class here (SMBHmass = 10): move the last two next to the planet, similar masses
    ps["pos"][s[2]] = ps["pos"][s[1]] + [0.004, 0.0, 0.0]
    ps["pos"][s[3]] = ps["pos"][s[1]] + [0.0, 0.005, 0.0]
    ps["mass"][s[1:]] = [4.0e-3, 4.0e-3, 3.9e-3]
    ps["vel"][s[2]] = ps["vel"][s[1]]
    ps["vel"][s[3]] = ps["vel"][s[1]]
    # time bins make the active list visit the sinks in the order s[3], s[1], s[2], s[0] (lists are per bin, run.c:190)
    ps["timebin"][:] = 4
    ps["timebin"][s] = [6, 3, 5, 2]
    out = util.run_oracle(util.REF_HARNESS, ps, par, OPS)
    pre, post, q = _split(_stage(out, OPS.split().index("B"), "B"))
    n = pre["Pos"].shape[0]
    fp = _forcepath(pre, q, par)
    S = dict(Vel=pre["Vel"].copy(), ID=pre["id"].copy(), SwallowID=pre["SwallowID"].copy(), BH_Mass=pre["BH_Mass"].copy(),
             BH_Mdot=pre["BH_Mdot"].copy(), AccDisc_Mass=pre["AccDisc_Mass"].copy(), Dust_Mass=pre["Dust_Mass"].copy(),
             Total_Mass=pre["Total_Mass"].copy(), Injected_BH_Energy=_pad(pre["Injected_BH_Energy"], n), SinkDensity=pre["SinkDensity"].copy())
    mass, counts = fp.blackhole_accretion(S, active=pre["active"])
    fp.close()
    t = pre["Type"]
    marked = post["SwallowID"][t == 5]
    assert (marked != 0).sum() >= 2, "the set-up must make sinks claim each other: %r" % (marked,)
    assert np.array_equal(S["SwallowID"], post["SwallowID"])
    assert np.array_equal(mass == 0, post["Mass"] == 0)
    _close(mass, post["Mass"], "Mass")
    _close(S["Vel"], post["Vel"], "Vel")
    for k in ("BH_Mass", "AccDisc_Mass", "Dust_Mass", "Total_Mass"):
        _close(S[k][t == 5], post[k][t == 5], k)
    ng = post["Injected_BH_Energy"].shape[0]
    _close(S["Injected_BH_Energy"][:ng], post["Injected_BH_Energy"], "Injected_BH_Energy")


def test_active_subset_and_empty_cases(ref_run):
    """Only the particles on the active list are targets; a list without sinks / dust does nothing and costs nothing."""
    par, out = ref_run
    pre, post, q = _split(_stage(out, OPS.split().index("Y"), "Y"))
    fp = _forcepath(pre, q, par)
    t = pre["Type"]
    dust = np.nonzero(t == 2)[0]
    some = dust[::3].astype(np.int32)
    d7 = np.full(t.shape[0], -1.0)
    fp.dust_density(d7, active=some)
    assert (d7[np.setdiff1d(np.arange(t.shape[0]), some)] == -1.0).all()
    _close(d7[some], post["DustParticleDensity"][some], "DUST_particle_density (subset)")
    gas_only = np.nonzero(t == 0)[0][:100].astype(np.int32)
    d7[:] = -1.0
    fp.dust_density(d7, active=gas_only)
    assert (d7 == -1.0).all()
    n = t.shape[0]
    S = dict(Vel=pre["Vel"].copy(), ID=pre["id"].copy(), SwallowID=np.zeros(n, np.int32), BH_Mass=pre["BH_Mass"].copy(),
             BH_Mdot=pre["BH_Mdot"].copy(), AccDisc_Mass=pre["AccDisc_Mass"].copy(), Dust_Mass=pre["Dust_Mass"].copy(),
             Total_Mass=pre["Total_Mass"].copy(), Injected_BH_Energy=np.zeros(n), SinkDensity=pre["SinkDensity"].copy())
    mass, counts = fp.blackhole_accretion(S, active=gas_only)
    assert counts == (0, 0, 0) and np.array_equal(mass, pre["Mass"]) and not S["SwallowID"].any()
    fp.close()


def test_subgrid_calls_fail_loudly_without_their_inputs():
    from gadget_leicester_b200.gadget import ForcePath
    from gadget_leicester_b200.lib import GLError
    ps = ics.disc_subgrid(2000, 200, seed=3)
    fp = ForcePath(ics.default_params())
    fp.load(ps)
    with pytest.raises(GLError, match="gl_set_subgrid_params"):
        fp.dust_density(np.zeros(ps["pos"].shape[0]))
    fp.set_subgrid_params()
    with pytest.raises(GLError, match="no tree"):
        fp.dust_density(np.zeros(ps["pos"].shape[0]))
    fp.close()
