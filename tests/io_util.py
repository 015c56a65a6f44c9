"""Shared pieces of the snapshot / IC file tests: a small particle set with every type and every file-only field
filled, and the driver of the reference's own writer / reader (oracle/_ref/ref_io: io.c + read_ic.c compiled
unmodified - TEST INFRASTRUCTURE, see oracle/refbuild/ref_io.c)."""
import os
import subprocess

import numpy as np

from gadget_leicester_b200 import aos, snapshot

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_IO = os.path.join(ROOT, "oracle", "_ref", "ref_io")
REF_IO_DROPIN = os.path.join(ROOT, "oracle", "_ref", "ref_io_dropin")      # same main, read_ic/savepositions from the glue

P_IO_FIELDS = ("io_p_DustRadius", "io_p_StellarAge", "io_p_BH_Mass", "io_p_BH_Mdot", "io_p_AccDisc_Mass", "io_p_BH_Mseed",
               "io_p_NewDensity", "io_p_DeltaPhotonMomentum", "io_p_OldPhotonMomentum", "io_p_BirthPhotonEnergy")


def particle_arrays(n=60, ngas=25, seed=11, zero_u=0):
    """GadgetArrays with gas first and the other five types interleaved behind it (the writer has to pick them out by
    Type, io.c:212-216), individual time bins, and accelerations so that the velocity record needs its kick."""
    g = np.random.default_rng(seed)
    types = np.concatenate([np.zeros(ngas, np.int32), (np.arange(n - ngas) % 5 + 1).astype(np.int32)])
    ps = dict(pos=g.standard_normal((n, 3)) * 3, vel=g.standard_normal((n, 3)), mass=g.random(n) + 0.01, type=types,
              id=g.permutation(n).astype(np.uint32) + 1000, hsml=g.random(n) + 0.05, oldacc=g.random(n),
              timebin=g.integers(0, 6, n).astype(np.int32), ti_begstep=(g.integers(0, 4, n) * 32).astype(np.int32),
              entropy=g.random(n) + 0.1, dtentropy=g.standard_normal(n))
    a = aos.GadgetArrays(ps, pinned=False)
    a.pview("p_GravAccel", np.float64, 3)[:] = g.standard_normal((n, 3))
    for f in P_IO_FIELDS:
        a.pview(f, np.float64)[:] = g.random(n) * 10
    if ngas:
        a.sview("s_HydroAccel", np.float64, 3)[:] = g.standard_normal((ngas, 3))
        a.sview("s_Density", np.float64)[:] = g.random(ngas) * 5 + 0.01
        a.sview("io_s_Ne", np.float64)[:] = g.random(ngas)
        a.sview("io_s_accreted_mass", np.float64)[:] = g.random(ngas)
        a.sview("io_s_accreted_momentum", np.float64, 3)[:] = g.standard_normal((ngas, 3))
        if zero_u:
            a.sview("s_Entropy", np.float64)[:zero_u] = 0.0
    return a


def derived_records(a):
    """NH and SFR as oracle/refbuild/ref_io.c's stand-ins for AbundanceRatios / get_starformation_rate produce them"""
    rho = a.sview("s_Density", np.float64)
    return 0.25 + 0.5 * rho / (1.0 + rho), 1.0e-3 * (np.arange(a.N_gas) % 17)


def common_args(par):
    args = ["format=%d" % par.format, "cooling=%d" % par.CoolingOn, "sfr=%d" % par.StarformationOn, "restart=%d" % par.RestartFlag,
            "minegy=%.17g" % par.MinEgySpec, "initgastemp=%.17g" % par.InitGasTemp, "ti_current=%d" % par.Ti_Current,
            "timebase=%.17g" % par.Timebase_interval, "time=%.17g" % par.Time, "box=%.17g" % par.BoxSize,
            "omega0=%.17g" % par.Omega0, "omegal=%.17g" % par.OmegaLambda, "hubble=%.17g" % par.HubbleParam]
    return args + ["mt%d=%.17g" % (t, par.MassTable[t]) for t in range(6)]


def reference_write(a, par, workdir, binary=REF_IO):
    """savepositions() of the reference on a's P[] / SphP[] images -> file bytes"""
    lay = aos.layout()
    pf, sf = os.path.join(workdir, "P.img"), os.path.join(workdir, "SphP.img")
    a.P[: a.NumPart * lay["p_stride"]].tofile(pf)
    a.SphP[: max(a.N_gas, 1) * lay["sph_stride"]].tofile(sf)
    subprocess.run([binary, "W", "p=" + pf, "sph=" + sf, "n=%d" % a.NumPart, "ngas=%d" % a.N_gas, "dir=" + workdir, "base=ref"]
                   + common_args(par), check=True, stdout=subprocess.DEVNULL)
    with open(os.path.join(workdir, "ref_000"), "rb") as f:
        return f.read()


READ_P = (("p_Pos", np.float64, 3), ("p_Vel", np.float64, 3), ("p_ID", np.uint32, 1), ("p_Mass", np.float64, 1), ("p_Type", np.int16, 1))
READ_P_SNAP = (("p_Hsml", np.float64, 1),) + tuple((f, np.float64, 1) for f in P_IO_FIELDS)
READ_S = (("s_Entropy", np.float64, 1),)
READ_S_SNAP = (("s_Density", np.float64, 1), ("io_s_Ne", np.float64, 1), ("io_s_accreted_mass", np.float64, 1),
               ("io_s_accreted_momentum", np.float64, 3))


def read_fields(a, restart, cooling=True):
    """the fields a reader sets, as plain arrays (everything else in the structs is left as it was)"""
    out = {}
    types = a.pview("p_Type", np.int16)
    only = {"p_Hsml": 0, "io_p_StellarAge": 4, "io_p_BH_Mass": 5, "io_p_BH_Mdot": 5, "io_p_AccDisc_Mass": 5, "io_p_BH_Mseed": 5,
            "io_p_NewDensity": 3, "io_p_DeltaPhotonMomentum": 3, "io_p_OldPhotonMomentum": 3, "io_p_BirthPhotonEnergy": 3}
    for f, dt, c in READ_P + (READ_P_SNAP if restart == 2 else ()):
        v = a.pview(f, dt, c)
        out[f] = (v[types == only[f]] if f in only else v).copy()      # a record only reaches the types that carry it
    for f, dt, c in READ_S + (READ_S_SNAP if restart == 2 else ()):
        if f == "io_s_Ne" and not cooling:
            continue
        out[f] = a.sview(f, dt, c).copy()
    if "io_s_accreted_momentum" in out:
        # the ACMO record holds one value per gas particle (get_bytes_per_blockelement, io.c:1184): the first N_gas
        # entries of the (particle, component) stream are real, the reference fills the rest from a stale buffer
        out["io_s_accreted_momentum"] = out["io_s_accreted_momentum"].reshape(-1)[: a.N_gas]
    return out


def reference_read(fname, par, workdir, binary=REF_IO):
    """read_ic() of the reference -> (fields dict, meta dict)"""
    lay = aos.layout()
    pf, sf, mf = (os.path.join(workdir, x) for x in ("Pr.img", "SphPr.img", "meta.txt"))
    subprocess.run([binary, "R", "file=" + str(fname), "p=" + pf, "sph=" + sf, "meta=" + mf] + common_args(par), check=True,
                   stdout=subprocess.DEVNULL)
    meta = {}
    with open(mf) as f:
        for line in f:
            k, v = line.split()
            meta[k] = float(v)
    n, ng = int(meta["NumPart"]), int(meta["N_gas"])
    a = aos.GadgetArrays.empty(n, max(ng, 1))
    a.NumPart, a.N_gas = n, ng
    a.P[: n * lay["p_stride"]] = np.fromfile(pf, np.uint8)
    if ng:
        a.SphP[: ng * lay["sph_stride"]] = np.fromfile(sf, np.uint8)
    return read_fields(a, par.RestartFlag, bool(par.CoolingOn)), meta
