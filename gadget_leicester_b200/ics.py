"""Synthetic particle sets for the five BASELINE.json configs.

The reference ships no initial-conditions file (its param file points at an absent
INIT_COND/Pd_N16e5_..., Pd_N16e5_*.txt:3), so every workload is generated here from a
Philox stream with key = seed.  All arrays are float64 / int32 in the P[]/SphP[] order the
reference expects: gas (Type 0) first, then the other types (allvars.h:1131, :1384).

Returned dict keys: pos[N,3] vel[N,3] mass[N] type[N] hsml[N] id[N] timebin[N] oldacc[N]
entropy[N] dtentropy[N]   (entropy/dtentropy only meaningful for the leading gas block).
"""
import numpy as np

GAMMA = 7.0 / 5.0  # allvars.h:64 (compile-time constant of the reference build)


def default_params(**over):
    """Run-time parameters of the shipped run (Pd_N16e5_*.txt) that the hot path reads."""
    p = dict(
        ErrTolTheta=0.3,              # .txt:65
        TypeOfOpeningCriterion=1,     # .txt:66
        ErrTolForceAcc=0.003,         # .txt:67
        DesNumNgb=40,                 # .txt:72
        MaxNumNgbDeviation=0.5,       # .txt:128
        ArtBulkViscConst=1.0,         # .txt:73
        G=1.0,                        # units .txt:86-90 give G = 1.0000
        # SofteningGas, Halo, Disk, Bulge, Stars, Bndry  (.txt:97-102)
        Softening=[1.0e-5, 0.01, 5.0e-5, 1.0e-7, 0.01, 5.0e-5],
        MinGasHsmlFractional=1.0,     # .txt:116
        Timebase_interval=100.0 / (1 << 29),  # (TimeMax-TimeBegin)/TIMEBASE, begrun.c
        Ti_Current=0,
        BlackHoleNgbFactor=4.0,       # .txt
        PartAllocFactor=2.5,
    )
    p.update(over)
    return p


def _rng(seed):
    return np.random.Generator(np.random.Philox(key=seed))


def _finish(pos, vel, mass, ptype, entropy, hsml=None):
    n = pos.shape[0]
    order = np.argsort(ptype != 0, kind="stable")  # gas first, stable
    pos, vel, mass, ptype, entropy = pos[order], vel[order], mass[order], ptype[order], entropy[order]
    if hsml is not None:
        hsml = hsml[order]
    else:
        hsml = np.zeros(n)
    return dict(
        pos=np.ascontiguousarray(pos, dtype=np.float64),
        vel=np.ascontiguousarray(vel, dtype=np.float64),
        mass=np.ascontiguousarray(mass, dtype=np.float64),
        type=np.ascontiguousarray(ptype, dtype=np.int32),
        hsml=np.ascontiguousarray(hsml, dtype=np.float64),
        id=np.arange(1, n + 1, dtype=np.int32),
        timebin=np.zeros(n, dtype=np.int32),
        oldacc=np.zeros(n, dtype=np.float64),
        entropy=np.ascontiguousarray(entropy, dtype=np.float64),
        dtentropy=np.zeros(n, dtype=np.float64),
    )


def disc(n_gas, seed=42, with_sinks=True, m_disc=0.01, m_star=1.0, m_planet=4.77e-3,
         r_in=0.1, r_out=1.0, aspect=0.05):
    """Configs 0 and 2: protoplanetary disc, Sigma ~ 1/R, Keplerian, + star and planet sinks."""
    g = _rng(seed)
    R = r_in + (r_out - r_in) * g.random(n_gas)        # Sigma ~ 1/R  <=>  R uniform
    phi = 2 * np.pi * g.random(n_gas)
    z = g.standard_normal(n_gas) * aspect * R
    pos = np.stack([R * np.cos(phi), R * np.sin(phi), z], axis=1)
    vk = np.sqrt(m_star / R)
    vel = np.stack([-vk * np.sin(phi), vk * np.cos(phi), np.zeros(n_gas)], axis=1)
    mass = np.full(n_gas, m_disc / n_gas)
    sigma = m_disc / (2 * np.pi * (r_out - r_in) * R)
    rho = sigma / (np.sqrt(2 * np.pi) * aspect * R) * np.exp(-0.5 * (z / (aspect * R)) ** 2)
    rho = np.maximum(rho, 1e-12)
    entropy = (aspect * vk) ** 2 / rho ** (GAMMA - 1.0)
    ptype = np.zeros(n_gas, dtype=np.int32)
    if with_sinks:
        rp = 0.5
        pos = np.vstack([pos, [[0.0, 0.0, 0.0], [rp, 0.0, 0.0]]])
        vel = np.vstack([vel, [[0.0, 0.0, 0.0], [0.0, np.sqrt(m_star / rp), 0.0]]])
        mass = np.concatenate([mass, [m_star, m_planet]])
        ptype = np.concatenate([ptype, [5, 5]]).astype(np.int32)
        entropy = np.concatenate([entropy, [0.0, 0.0]])
    return _finish(pos, vel, mass, ptype, entropy)


def disc_subgrid(n_gas, n_dust, seed=42, m_disc=0.01, m_star=1.0, m_planet=4.77e-3, r_in=0.008, r_out=1.0, aspect=0.05, clump_frac=0.3):
    """The disc of config 0 arranged so that the sub-grid neighbour loops behind the force path have work
    (blackhole.c:794/:1201, dust.c:748/:889): gas reaching inside InnerBoundary of the star, dust grains (Type 2) mixed
    into the disc, around the star and around the planet, a small third sink next to the planet and a fourth next to
    the star (mergers).  Gas first, then the other types in the order dust, sinks."""
    g = _rng(seed)
    ps = disc(n_gas, seed=seed, with_sinks=False, m_disc=m_disc, m_star=m_star, r_in=r_in, r_out=r_out, aspect=aspect)
    rp = 0.5
    # dust: 70 % through the disc, 15 % within 0.03 of the star, 15 % within 0.03 of the planet (clump_frac = 0.3)
    na = int((1.0 - clump_frac) * n_dust)
    nb = (n_dust - na) // 2
    nc = n_dust - na - nb
    R = r_in + (r_out - r_in) * g.random(na)
    phi = 2 * np.pi * g.random(na)
    z = g.standard_normal(na) * 0.5 * aspect * R
    dpos = [np.stack([R * np.cos(phi), R * np.sin(phi), z], axis=1)]
    dpos.append(0.03 * (g.random((nb, 3)) - 0.5) * [1, 1, 0.3])
    dpos.append(np.array([rp, 0, 0]) + 0.03 * (g.random((nc, 3)) - 0.5) * [1, 1, 0.3])
    dpos = np.vstack(dpos)
    Rd = np.maximum(np.hypot(dpos[:, 0], dpos[:, 1]), 1e-3)
    vk = np.sqrt(m_star / Rd)
    dvel = np.stack([-vk * dpos[:, 1] / Rd, vk * dpos[:, 0] / Rd, np.zeros(n_dust)], axis=1) * (0.9 + 0.2 * g.random((n_dust, 1)))
    dmass = np.full(n_dust, 0.01 * m_disc / max(n_dust, 1))
    # sinks: star, planet, a light sink 2e-5 from the planet, a light sink 0.01 from the star
    spos = np.array([[0.0, 0.0, 0.0], [rp, 0.0, 0.0], [rp + 2.0e-5, 1.0e-5, 0.0], [0.01, 0.002, 0.0]])
    svel = np.array([[0.0, 0.0, 0.0], [0.0, np.sqrt(m_star / rp), 0.0], [0.0, np.sqrt(m_star / rp), 0.0], [0.0, 3.0, 0.0]])
    smass = np.array([m_star, m_planet, 1.0e-5, 2.0e-5])
    ng = n_gas
    pos = np.vstack([ps["pos"][:ng], dpos, spos])
    vel = np.vstack([ps["vel"][:ng], dvel, svel])
    mass = np.concatenate([ps["mass"][:ng], dmass, smass])
    ptype = np.concatenate([np.zeros(ng), np.full(n_dust, 2), np.full(4, 5)]).astype(np.int32)
    entropy = np.concatenate([ps["entropy"][:ng], np.zeros(n_dust + 4)])
    return _finish(pos, vel, mass, ptype, entropy)


def evrard(n_gas, seed=42, u0=0.05):
    """Config 1: Evrard sphere, M=1, R=1, rho ~ 1/r, at rest, u = 0.05."""
    g = _rng(seed)
    r = np.sqrt(g.random(n_gas))                         # M(<r) ~ r^2
    ct = 2 * g.random(n_gas) - 1
    st = np.sqrt(1 - ct * ct)
    phi = 2 * np.pi * g.random(n_gas)
    pos = np.stack([r * st * np.cos(phi), r * st * np.sin(phi), r * ct], axis=1)
    vel = np.zeros((n_gas, 3))
    mass = np.full(n_gas, 1.0 / n_gas)
    rho = 1.0 / (2 * np.pi * np.maximum(r, 1e-3))
    entropy = (GAMMA - 1.0) * u0 / rho ** (GAMMA - 1.0)
    return _finish(pos, vel, mass, np.zeros(n_gas, dtype=np.int32), entropy)


def plummer(n, seed=42, a=1.0, rmax=10.0):
    """Config 3: collisionless (Type 1) Plummer sphere, M=1, scale a, truncated at rmax*a."""
    g = _rng(seed)
    out = np.empty((0,))
    r = np.empty(0)
    while r.size < n:
        x = g.random(n - r.size + 16)
        x = x[(x > 0) & (x < 1)]
        rr = a / np.sqrt(x ** (-2.0 / 3.0) - 1.0)
        r = np.concatenate([r, rr[rr < rmax * a]])
    r = r[:n]
    ct = 2 * g.random(n) - 1
    st = np.sqrt(1 - ct * ct)
    phi = 2 * np.pi * g.random(n)
    pos = np.stack([r * st * np.cos(phi), r * st * np.sin(phi), r * ct], axis=1)
    # isotropic velocities with the local 1-D dispersion (good enough for a force benchmark)
    sig = np.sqrt(1.0 / (6.0 * np.sqrt(r * r + a * a)))
    vel = g.standard_normal((n, 3)) * sig[:, None]
    mass = np.full(n, 1.0 / n)
    return _finish(pos, vel, mass, np.ones(n, dtype=np.int32), np.zeros(n))


def plummer_softening(n, a=1.0):
    """SofteningHalo of config 3: 0.01 a (N/1e6)^(-1/3)."""
    return 0.01 * a * (n / 1.0e6) ** (-1.0 / 3.0)


def cube(n_gas, seed=42, vamp=0.05):
    """Config 4: uniform random gas cube [0,1)^3, equal mass, A = 1, velocities U(-vamp, vamp)."""
    g = _rng(seed)
    pos = g.random((n_gas, 3))
    vel = (2 * g.random((n_gas, 3)) - 1) * vamp
    mass = np.full(n_gas, 1.0 / n_gas)
    return _finish(pos, vel, mass, np.zeros(n_gas, dtype=np.int32), np.ones(n_gas))


def guess_hsml(ps, des_num_ngb):
    """A cheap initial Hsml for cold-start density tests: local mean spacing from a coarse grid.

    (The reference derives its guess from its tree, init.c:719-745; any guess works because
    density() iterates to the same fixed point window.)"""
    pos = ps["pos"]
    gas = ps["type"] == 0
    lo, hi = pos[gas].min(0), pos[gas].max(0)
    ext = float((hi - lo).max()) * 1.001 + 1e-300
    ng = int(gas.sum())
    m = max(2, int(round((ng / 8.0) ** (1.0 / 3.0))))
    idx = np.minimum(((pos[gas] - lo) / ext * m).astype(np.int64), m - 1)
    cell = (idx[:, 0] * m + idx[:, 1]) * m + idx[:, 2]
    cnt = np.bincount(cell, minlength=m ** 3)
    dens = cnt[cell] / (ext / m) ** 3
    h = (3.0 / (4 * np.pi) * des_num_ngb / dens) ** (1.0 / 3.0)
    out = np.zeros(pos.shape[0])
    out[gas] = h
    return out


# ---- device-side generators for the 2^27 configs (host numpy would need > 100 GB across 8 ranks) ----------------
def plummer_torch(n, device, seed=42, a=1.0, rmax=10.0):
    """Config 3 on the GPU: same distribution as plummer(), torch Philox stream."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = torch.rand(n, generator=g, device=device, dtype=torch.float64).clamp_(1e-12, 1 - 1e-12)
    # truncate at rmax by rescaling the mass fraction: M(<r)/M = r^3 / (r^2 + a^2)^1.5
    fmax = (rmax ** 3) / (rmax * rmax + 1.0) ** 1.5
    x = x * fmax
    r = a / torch.sqrt(x ** (-2.0 / 3.0) - 1.0)
    ct = 2 * torch.rand(n, generator=g, device=device, dtype=torch.float64) - 1
    st = torch.sqrt(1 - ct * ct)
    phi = 2 * np.pi * torch.rand(n, generator=g, device=device, dtype=torch.float64)
    pos = torch.stack([r * st * torch.cos(phi), r * st * torch.sin(phi), r * ct], dim=1).contiguous()
    sig = torch.sqrt(1.0 / (6.0 * torch.sqrt(r * r + a * a)))
    vel = (torch.randn(n, 3, generator=g, device=device, dtype=torch.float64) * sig[:, None]).contiguous()
    return _finish_torch(pos, vel, torch.full((n,), 1.0 / n, device=device, dtype=torch.float64),
                         torch.ones(n, device=device, dtype=torch.int32), torch.zeros(n, device=device, dtype=torch.float64))


def cube_torch(n, device, seed=42, vamp=0.05):
    """Config 4 on the GPU: uniform random gas cube."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    pos = torch.rand(n, 3, generator=g, device=device, dtype=torch.float64)
    vel = (2 * torch.rand(n, 3, generator=g, device=device, dtype=torch.float64) - 1) * vamp
    return _finish_torch(pos, vel, torch.full((n,), 1.0 / n, device=device, dtype=torch.float64),
                         torch.zeros(n, device=device, dtype=torch.int32), torch.ones(n, device=device, dtype=torch.float64))


def _finish_torch(pos, vel, mass, ptype, entropy):
    import torch
    n = pos.shape[0]
    dev = pos.device
    z = lambda dt: torch.zeros(n, device=dev, dtype=dt)
    return dict(pos=pos, vel=vel, mass=mass, type=ptype, hsml=z(torch.float64), oldacc=z(torch.float64), timebin=z(torch.int32),
                ti_begstep=z(torch.int32), entropy=entropy, dtentropy=z(torch.float64))
