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
