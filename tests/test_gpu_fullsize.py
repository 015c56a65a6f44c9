"""GPU suite: BASELINE.json config 0 (1.6M gas + 2 sinks) and config 1 (Evrard sphere, 1M gas) at FULL size, checked
through size-independent properties because no CPU oracle finishes there in seconds; configs 3 and 4 (2^27 in
bench.py --config 3/4) are checked at 2^24 particles with the same device-side generators:
  * the interaction total equals the sum of per-particle GravCost (a checksum of checksums);
  * tree forces agree with an FP64 direct summation on a random sample (accuracy of the criterion);
  * every SPH target ends inside the neighbour window of density.c:559 (or at the Hsml floor);
  * density / pair counts of a sample equal a brute-force FP64 evaluation at the final Hsml, exactly;
  * hydro forces conserve momentum (the pair force is antisymmetric, hydra.c:1602-1618);
  * a second evaluation is idempotent (same Hsml in -> one pass, identical bits out).
"""
import numpy as np
import pytest

from gadget_leicester_b200 import ics

pytestmark = pytest.mark.gpu

N_GAS = 1600000


@pytest.fixture(scope="module", params=["config0_disc", "config1_evrard"])
def state(request):
    from gadget_leicester_b200.gadget import ForcePath
    if request.param == "config0_disc":
        ps = ics.disc(N_GAS, seed=42)
        par = ics.default_params()
    else:
        ps = ics.evrard(1000000, seed=42)
        par = ics.default_params(Softening=[1e-4, 5e-5, 5e-5, 1e-7, 0.01, 5e-5])
    ps["hsml"] = ics.guess_hsml(ps, par["DesNumNgb"])
    ps["hsml"][ps["type"] == 5] = par["Softening"][5]
    fp = ForcePath(par)
    fp.load(ps)
    g1 = fp.gravity_tree()      # Barnes-Hut pass (OldAcc = 0), builds the tree
    g2 = fp.gravity_tree()      # relative criterion
    st_g = fp.stats()
    d = fp.density()
    st_d = fp.stats()
    fp.force_update_hmax()
    h = fp.hydro_force()
    yield dict(fp=fp, ps=ps, par=par, g1=g1, g2=g2, d=d, h=h, st_g=st_g, st_d=st_d)
    fp.close()


def test_interaction_checksum(state):
    assert int(state["g2"]["GravCost"].astype(np.int64).sum()) == state["st_g"]["interactions"]
    assert state["st_g"]["interactions"] > state["ps"]["pos"].shape[0]  # at least the self term + root


def test_tree_force_vs_direct_sum(state):
    ps, g2 = state["ps"], state["g2"]
    n = ps["pos"].shape[0]
    rng = np.random.default_rng(3)
    sample = rng.choice(n, 48, replace=False)
    soft = np.where(ps["type"] == 0, state["d"]["Hsml"], 2.8 * np.asarray(state["par"]["Softening"])[ps["type"]])
    errs = []
    for i in sample:
        dx = ps["pos"] - ps["pos"][i]
        r2 = (dx * dx).sum(1)
        r = np.sqrt(r2)
        h = np.maximum(soft[i], soft)
        u = r / h
        with np.errstate(divide="ignore", invalid="ignore"):
            newton = ps["mass"] / (r2 * r)
            inner = ps["mass"] / h ** 3 * (10.666666666667 + u * u * (32.0 * u - 38.4))
            outer = ps["mass"] / h ** 3 * (21.333333333333 - 48.0 * u + 38.4 * u * u - 10.666666666667 * u ** 3
                                             - 0.066666666667 / u ** 3)
        fac = np.where(r >= h, newton, np.where(u < 0.5, inner, outer))
        fac[i] = 0.0
        acc = (dx * fac[:, None]).sum(0) * state["par"]["G"]
        errs.append(np.linalg.norm(g2["GravAccel"][i] - acc) / np.linalg.norm(acc))
    errs = np.array(errs)
    # note: the tree was built with the pre-density Hsml; softening only matters inside h, so tolerate that
    assert np.median(errs) < 5e-3 and errs.max() < 5e-2, errs


def test_neighbour_window(state):
    ps, par, d = state["ps"], state["par"], state["d"]
    sph = (ps["type"] == 0) | (ps["type"] == 5)
    des = np.where(ps["type"] == 5, par["DesNumNgb"] * par["BlackHoleNgbFactor"], par["DesNumNgb"])
    dev = np.abs(d["NumNgb"] - des)
    floor = 1.01 * par["MinGasHsmlFractional"] * 2.8 * par["Softening"][0]
    ok = (dev <= par["MaxNumNgbDeviation"]) | (d["Hsml"] <= floor)
    # density.c:566-573 also accepts a particle whose bracket closed to 1e-3; allow a handful
    assert (~ok[sph]).sum() <= 16, int((~ok[sph]).sum())


def test_density_sample_bruteforce(state):
    ps, d = state["ps"], state["d"]
    gas = np.where(ps["type"] == 0)[0]
    rng = np.random.default_rng(5)
    pos_g, m_g = ps["pos"][gas], ps["mass"][gas]
    for i in rng.choice(gas, 24, replace=False):
        h = d["Hsml"][i]
        dx = ps["pos"][i] - pos_g
        r2 = dx[:, 0] * dx[:, 0] + dx[:, 1] * dx[:, 1] + dx[:, 2] * dx[:, 2]   # density.c:853 operation order
        hit = r2 < h * h
        assert int(hit.sum()) == int(d["DensPairs"][i])
        u = np.sqrt(r2[hit]) / h
        hinv3 = 1.0 / h ** 3
        wk = np.where(u < 0.5, hinv3 * (2.546479089470 + 15.278874536822 * (u - 1) * u * u),
                      hinv3 * 5.092958178941 * (1.0 - u) ** 3)
        rho = (m_g[hit] * wk).sum()
        assert abs(rho - d["Density"][i]) <= 1e-12 * rho


def test_hydro_momentum_conservation(state):
    ps, h = state["ps"], state["h"]
    gas = ps["type"] == 0
    f = ps["mass"][gas, None] * h["HydroAccel"][gas]
    assert np.abs(f.sum(0)).max() <= 1e-9 * np.abs(f).sum()


def test_idempotent_second_evaluation(state):
    """Same Hsml in -> same neighbour sets and (to FP64 round-off) the same sums out.  The only particles allowed to
    move are those the reference accepts through its closed-bracket exit (density.c:566-573) with NumNgb outside
    the window: their Left/Right restart on every call, in the reference too."""
    fp, d, ps, par = state["fp"], state["d"], state["ps"], state["par"]
    d2 = fp.density()
    des = np.where(ps["type"] == 5, par["DesNumNgb"] * par["BlackHoleNgbFactor"], par["DesNumNgb"])
    inside = np.abs(d["NumNgb"] - des) <= par["MaxNumNgbDeviation"] * (1 - 1e-9)
    sph = (ps["type"] == 0) | (ps["type"] == 5)
    assert (sph & ~inside).sum() <= 16
    ok = sph & inside
    assert (d2["Hsml"][ok] == d["Hsml"][ok]).all()
    assert (d2["DensPairs"][ok] == d["DensPairs"][ok]).all()
    gas_in = (ps["type"] == 0) & inside
    assert np.abs(d2["Density"][gas_in] / d["Density"][gas_in] - 1).max() <= 1e-13


# ---- configs 3 and 4 at 2^24 particles (device-side generators, nothing of size N on the host) -------------------
N_BIG = 1 << 24


def _dfp(par):
    from gadget_leicester_b200.multigpu import DistributedForcePath
    return DistributedForcePath(par, rank=0, world=1, device=0)


def test_config3_plummer_gravity_large():
    """Collisionless Plummer sphere (BASELINE config 3 generator): interaction checksum, and the tree force of a sample
    against an FP64 direct sum over all 2^24 particles (on the GPU)."""
    import torch
    par = ics.default_params(Softening=[1e-5, ics.plummer_softening(N_BIG), 5e-5, 1e-7, 0.01, 5e-5])
    dev = torch.device("cuda", 0)
    ps = ics.plummer_torch(N_BIG, dev, seed=42)
    d = _dfp(par)
    order = d.decompose_device(ps)
    d.compute_accelerations(sph=False)          # Barnes-Hut pass
    d.compute_accelerations(sph=False)          # relative criterion
    st = d.fp.stats()
    assert int(d.res["gravcost"].double().sum().item()) == st["interactions"]
    pos = ps["pos"].index_select(0, order.long())         # Peano-Hilbert order = result order
    mass = ps["mass"].index_select(0, order.long())
    soft = 2.8 * par["Softening"][1]
    sample = torch.randint(0, N_BIG, (12,), device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    errs = []
    for i in sample.tolist():
        dx = pos - pos[i]
        r2 = (dx * dx).sum(1)
        r = torch.sqrt(r2)
        u = r / soft
        newton = mass / (r2 * r)
        inner = mass / soft ** 3 * (10.666666666667 + u * u * (32.0 * u - 38.4))
        outer = mass / soft ** 3 * (21.333333333333 - 48.0 * u + 38.4 * u * u - 10.666666666667 * u ** 3 - 0.066666666667 / u ** 3)
        fac = torch.where(r >= soft, newton, torch.where(u < 0.5, inner, outer))
        fac[i] = 0.0
        acc = (dx * fac[:, None]).sum(0) * par["G"]
        errs.append(float(torch.linalg.norm(d.res["acc"][i] - acc) / torch.linalg.norm(acc)))
    errs = np.array(errs)
    assert np.median(errs) < 5e-3 and errs.max() < 5e-2, errs
    d.close()


def test_config4_cube_sph_large():
    """Uniform random gas cube, Ngb = 64 (BASELINE config 4 generator): every target ends inside the neighbour window,
    a sample's pair counts and densities equal a brute-force FP64 evaluation at the final Hsml (on the GPU), the
    pair force is antisymmetric (momentum conservation)."""
    import torch
    par = ics.default_params(DesNumNgb=64, MaxNumNgbDeviation=0.5)
    dev = torch.device("cuda", 0)
    ps = ics.cube_torch(N_BIG, dev, seed=42)
    ps["hsml"].fill_((3.0 * 64 / (4 * np.pi * N_BIG)) ** (1.0 / 3.0) * 1.3)
    d = _dfp(par)
    order = d.decompose_device(ps)
    d.compute_accelerations(gravity=False)
    pos = ps["pos"].index_select(0, order.long())
    mass = ps["mass"].index_select(0, order.long())
    nn, hs, rho = d.res["numngb"], d.res["hsml"], d.res["density"]
    assert int(((nn - 64).abs() > 0.5).sum().item()) <= 16
    pairs = torch.zeros(N_BIG, dtype=torch.int32, device=dev)
    import ctypes as C
    d.fp._chk(d.fp._lib.gl_get_pair_counts(d.fp.ctx, C.c_void_p(pairs.data_ptr()), None, 1))
    sample = torch.randint(0, N_BIG, (12,), device=dev, generator=torch.Generator(device=dev).manual_seed(5))
    for i in sample.tolist():
        h = float(hs[i].item())
        dx = pos[i] - pos
        r2 = dx[:, 0] * dx[:, 0] + dx[:, 1] * dx[:, 1] + dx[:, 2] * dx[:, 2]      # density.c:853 operation order
        hit = r2 < h * h
        assert int(hit.sum().item()) == int(pairs[i].item())
        u = torch.sqrt(r2[hit]) / h
        hinv3 = 1.0 / h ** 3
        wk = torch.where(u < 0.5, hinv3 * (2.546479089470 + 15.278874536822 * (u - 1) * u * u), hinv3 * 5.092958178941 * (1.0 - u) ** 3)
        r_ref = float((mass[hit] * wk).sum().item())
        assert abs(r_ref - float(rho[i].item())) <= 1e-12 * r_ref
    f = mass[:, None] * d.res["hacc"]
    assert float(f.sum(0).abs().max().item()) <= 1e-9 * float(f.abs().sum().item())
    d.close()
