"""Throughput of the host-side rows of SURVEY.md 8(f) - the kick (gl_gadget_advance_and_find_timesteps), the lazy drift
(gl_gadget_drift) and the snapshot writer / reader (gl_snapshot_write / gl_snapshot_read) - on the config-0 particle
count, next to the reference's own single-rank code for the file I/O (oracle/_ref/ref_io, when built).  These loops
are bound by host memory bandwidth (536 B + 264 B structs, a few fields of each touched), so the figure to compare
with is bytes of P[] / SphP[] swept per second.  Prints one JSON line.  No GPU involved.

    python scripts/host_rows_bench.py [--ngas 1600000] [--reps 5]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gadget_leicester_b200 import aos, ics, snapshot       # noqa: E402
from gadget_leicester_b200 import lib as L                 # noqa: E402


def best(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ngas", type=int, default=1600000)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    lib = L.load_library()
    lay_d, lay = aos.layout(), aos.layout_struct()
    ps = ics.disc(args.ngas, seed=42)
    n = ps["pos"].shape[0]
    g = np.random.default_rng(1)
    ps["timebin"] = g.integers(1, 7, n).astype(np.int32)
    a = aos.GadgetArrays(ps, pinned=False)
    ng = a.N_gas
    a.pview("p_GravAccel", np.float64, 3)[:] = g.standard_normal((n, 3))
    a.pview("p_Ti_begstep", np.int32)[:] = 64 - (1 << ps["timebin"])
    a.pview("p_Ti_current", np.int32)[:] = 0
    a.pview("p_Hsml", np.float64)[:] = g.random(n) * 0.04 + 0.01
    a.sview("s_HydroAccel", np.float64, 3)[:] = g.standard_normal((ng, 3))
    a.sview("s_Density", np.float64)[:] = g.random(ng) + 0.1
    a.sview("s_MaxSignalVel", np.float64)[:] = g.random(ng) * 50 + 10
    a.sview("s_DivVel", np.float64)[:] = g.standard_normal(ng)
    P0, S0 = a.P.copy(), a.SphP.copy()
    struct_bytes = n * lay_d["p_stride"] + ng * lay_d["sph_stride"]
    out = dict(particles=n, gas=ng, host_threads=int(os.environ.get("GL_HOST_THREADS", 0)) or (os.cpu_count() or 2) - 1,
               struct_bytes=struct_bytes)

    # drift
    def drift():
        a.P[:] = P0
        a.SphP[:] = S0
    nd = C.c_longlong(0)
    t_copy = best(drift, 2)

    def drift_run():
        drift()
        lib.gl_gadget_drift(C.byref(lay), L._ptr(a.P), L._ptr(a.SphP), n, ng, 64, 1.0e-4, 0.0, C.byref(nd))
    t = best(drift_run, args.reps) - t_copy
    out["drift"] = dict(ms=t * 1e3, particles_per_s=n / t, struct_GBps=struct_bytes / t / 1e9)

    # kick
    tp = L.GLTimestepParams()
    tp.Ti_Current = 64
    for b in range(29):
        tp.TimeBinActive[b] = 1 if b <= 6 else 0
    tp.Timebase_interval, tp.ErrTolIntAccuracy, tp.CourantFac, tp.MaxSizeTimestep, tp.MinSizeTimestep = 1.0e-5, 0.025, 0.15, 8.0e-3, 0.0
    tp.MinEgySpec, tp.GammaMinus1 = 1.0e-4, 7. / 5. - 1
    for k, s in enumerate((1e-5, 1e-5, 5e-5, 1e-7, 0.01, 5e-5)):
        tp.SofteningTable[k] = s
    tp.OuterBoundary, tp.UnitVelocity_in_cm_per_s, tp.FeedBackVelocity = 1.4e10, 2.98e6, 1.0
    tp.SMBHmass, tp.InnerBoundary, tp.SinkBoundary, tp.AccDtBlackHole = 0.5, 2.0e-2, 1.0e-4, 0.007
    tp.p_DeltaDustMomentum, tp.p_NewDensity = lay_d["ts_p_DeltaDustMomentum"], lay_d["io_p_NewDensity"]
    active = np.arange(n, dtype=np.int32)
    binold = np.zeros(n, np.int32)
    err = C.create_string_buffer(512)
    rcs = []

    def kick_run():
        drift()
        rcs.append(lib.gl_gadget_advance_and_find_timesteps(C.byref(lay), C.byref(tp), L._ptr(a.P), L._ptr(a.SphP), L._ptr(active), n,
                                                            L._ptr(binold), err, len(err)))
    t = best(kick_run, args.reps) - t_copy
    assert not any(rcs), err.value
    out["kick"] = dict(ms=t * 1e3, particles_per_s=n / t, struct_GBps=struct_bytes / t / 1e9,
                       bins_changed=int((binold != a.pview("p_TimeBin", np.int16)).sum()))

    # peano_hilbert_order from given keys (random 63-bit keys: the worst case for the permutation's locality)
    keys = g.integers(0, 1 << 62, size=n, dtype=np.int64).astype(np.uint64)

    def ph_run():
        drift()
        assert lib.gl_gadget_peano_hilbert_order(C.byref(lay), L._ptr(a.P), L._ptr(a.SphP), n, ng, L._ptr(keys), None) == 0
    t = best(ph_run, max(2, args.reps // 2)) - t_copy
    out["peano_hilbert_order"] = dict(ms=t * 1e3, particles_per_s=n / t, struct_GBps=2 * struct_bytes / t / 1e9)

    # snapshot files
    drift()
    a.sview("s_Density", np.float64)[:] = g.random(ng) + 0.1
    par = snapshot.io_params(format=1, CoolingOn=1, StarformationOn=1, Ti_Current=64, Timebase_interval=4.0e-5)
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        fn = os.path.join(td, "snap_000")
        t = best(lambda: snapshot.write_snapshot(fn, a, par), max(2, args.reps // 2))
        size = os.path.getsize(fn)
        out["snapshot_write"] = dict(ms=t * 1e3, file_bytes=size, file_GBps=size / t / 1e9, particles_per_s=n / t)
        rp = snapshot.io_params(format=1, CoolingOn=1, StarformationOn=1, RestartFlag=2)
        b = aos.GadgetArrays.empty(n + 16, ng + 16)                 # allocate_memory() is not part of the read
        b.P[:] = 1
        b.SphP[:] = 1
        io, hdr, nn, nng = snapshot.io_layout_struct(), L.GLIOHeader(), C.c_longlong(0), C.c_longlong(0)

        def rd():
            assert lib.gl_snapshot_read(fn.encode(), C.byref(lay), C.byref(io), C.byref(rp), L._ptr(b.P), L._ptr(b.SphP), b.MaxPart,
                                        b.MaxPartSph, C.byref(nn), C.byref(nng), C.byref(hdr), err, len(err)) == 0, err.value
        t = best(rd, max(2, args.reps // 2))
        assert nn.value == n and nng.value == ng
        out["snapshot_read"] = dict(ms=t * 1e3, file_GBps=size / t / 1e9, particles_per_s=n / t)
        ref_io = os.path.join(ROOT, "oracle", "_ref", "ref_io")
        if os.path.exists(ref_io):
            pf, sf = os.path.join(td, "P.img"), os.path.join(td, "S.img")
            a.P[: n * lay_d["p_stride"]].tofile(pf)
            a.SphP[: ng * lay_d["sph_stride"]].tofile(sf)
            common = ["format=1", "cooling=1", "sfr=1", "ti_current=64", "timebase=4e-5"]
            t0 = time.perf_counter()
            subprocess.run([ref_io, "W", "p=" + pf, "sph=" + sf, "n=%d" % n, "ngas=%d" % ng, "dir=" + td, "base=ref"] + common,
                           check=True, stdout=subprocess.DEVNULL)
            tw = time.perf_counter() - t0
            t0 = time.perf_counter()
            subprocess.run([ref_io, "R", "file=" + os.path.join(td, "ref_000"), "restart=2", "p=" + pf + ".o", "sph=" + sf + ".o",
                            "meta=" + os.path.join(td, "m.txt")] + common, check=True, stdout=subprocess.DEVNULL)
            tr = time.perf_counter() - t0
            out["reference_io"] = dict(write_ms=tw * 1e3, read_ms=tr * 1e3,
                                       note="whole process of oracle/_ref/ref_io (1 core), includes loading / dumping the raw struct images")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
