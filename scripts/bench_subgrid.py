#!/usr/bin/env python
"""Time the sub-grid neighbour loops (blackhole_accretion, dust_density, dust_drag) on the tree of a config-0 sized step,
inputs resident in HBM (CUDA events on the library's stream), next to the reference's own CPU routines on a bounded
sample (oracle/_ref/ref_harness ops B / Y / Z, one core).

  python scripts/bench_subgrid.py [--ngas 1600000] [--ndust 160000] [--ref-ngas 200000] [--out profiles/r02_subgrid.json]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ngas", type=int, default=1600000)
    ap.add_argument("--ndust", type=int, default=160000)
    ap.add_argument("--ref-ngas", type=int, default=200000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--clump-frac", type=float, default=0.02, help="fraction of the dust packed around the two main sinks (the rest follows the gas disc)")
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_subgrid.json"))
    a = ap.parse_args()
    import torch
    import util
    from gadget_leicester_b200 import ics, lib as L
    from gadget_leicester_b200.gadget import ForcePath
    par = ics.default_params(BlackHoleFeedbackFactor=1e-3, SinkBoundary=0.02, MinEgySpec=1e-6)
    rec = dict(config="disc of config 0 with dust: %d gas + %d dust + 4 sinks (ics.disc_subgrid, seed 42, clump_frac %g)" % (a.ngas, a.ndust, a.clump_frac))

    ps = ics.disc_subgrid(a.ngas, a.ndust, seed=42, clump_frac=a.clump_frac)
    ps["timebin"][:] = 3 + (ps["id"] % 3)
    ps["hsml"] = ics.guess_hsml(ps, par["DesNumNgb"])
    ps["hsml"][ps["type"] == 5] = par["Softening"][5]     # init.c:741-745
    ps["hsml"][ps["type"] == 2] = par["Softening"][2]
    n = ps["pos"].shape[0]
    fp = ForcePath(par)
    fp.load(ps)
    fp.gravity_tree()
    g = fp.gravity_tree()
    d = fp.density()
    fp.set_subgrid_params(BlackHoleFeedbackFactor=1e-3, SinkBoundary=0.02, MinEgySpec=1e-6)
    lib, ctx = fp._lib, fp._ctx
    stream = torch.cuda.ExternalStream(lib.gl_stream(ctx))
    dev = torch.device("cuda:0")

    def T(x, dt=torch.float64):
        return torch.as_tensor(np.ascontiguousarray(x), dtype=dt, device=dev)
    t = ps["type"]
    sink = t == 5
    base = dict(vel=T(ps["vel"]), id=T(ps["id"], torch.int32), swallow=torch.zeros(n, dtype=torch.int32, device=dev),
                bh_mass=T(np.where(sink, 0.8 * ps["mass"], 0)), bh_mdot=torch.zeros(n, dtype=torch.float64, device=dev),
                accdisc=T(np.where(sink, 0.1 * ps["mass"], 0)), dust_mass=T(np.where(sink, 0.01 * ps["mass"], 0)),
                total=T(np.where(sink, ps["mass"], 0)), inj=torch.zeros(n, dtype=torch.float64, device=dev))
    times = {}

    def timed(name, fn, reps):
        ms = []
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        times[name] = dict(ms_min=min(ms), ms_median=float(np.median(ms)), reps=reps)

    # dust loops first (they do not change the particle set), each repetition on the same state
    d7 = torch.zeros(n, dtype=torch.float64, device=dev)
    timed("dust_density", lambda: fp._chk(lib.gl_dust_density(ctx, None, 0, C.c_void_p(d7.data_ptr()), 1)), a.reps)
    rad = T(10.0 ** ((ps["id"] % 7) - 3.0))
    ent0 = T(ps["entropy"])
    out = {k: torch.zeros(n * c, dtype=torch.float64, device=dev) for k, c in (("vcoll", 1), ("ddm", 3), ("nda", 3), ("dde", 1), ("heat", 1))}

    def drag():
        st = L.GLDustState()
        vel = base["vel"].clone()
        ent = ent0.clone()
        st.vel = vel.data_ptr(); st.dust_particle_density = d7.data_ptr(); st.dust_radius = rad.data_ptr()
        st.dust_vcoll = out["vcoll"].data_ptr(); st.delta_dust_momentum = out["ddm"].data_ptr(); st.new_drag_acc = out["nda"].data_ptr()
        st.delta_drag_energy = out["dde"].data_ptr(); st.entropy = ent.data_ptr(); st.drag_heating = out["heat"].data_ptr()
        torch.cuda.synchronize()
        fp._chk(lib.gl_dust_drag(ctx, None, 0, C.byref(st), 1))
    timed("dust_drag", drag, a.reps)
    # sink accretion last: it swallows particles, so every repetition starts from a freshly loaded set (the first one,
    # which also pays for the library's one-time buffer allocations, is not timed)
    cnt = (C.c_longlong * 3)()

    def fresh():
        return dict(vel=T(ps["vel"]), id=T(ps["id"], torch.int32), swallow=torch.zeros(n, dtype=torch.int32, device=dev),
                    bh_mass=T(np.where(sink, 0.8 * ps["mass"], 0)), bh_mdot=torch.zeros(n, dtype=torch.float64, device=dev),
                    accdisc=T(np.where(sink, 0.1 * ps["mass"], 0)), dust_mass=T(np.where(sink, 0.01 * ps["mass"], 0)),
                    total=T(np.where(sink, ps["mass"], 0)), inj=torch.zeros(n, dtype=torch.float64, device=dev))
    ms = []
    for rep in range(1 + min(a.reps, 3)):
        if rep:
            fp.load(dict(ps, hsml=d["Hsml"]))
            fp.force_treebuild()
        b = fresh()
        st = L.GLSinkState()
        st.vel = b["vel"].data_ptr(); st.id = b["id"].data_ptr(); st.swallow_id = b["swallow"].data_ptr()
        st.bh_mass = b["bh_mass"].data_ptr(); st.bh_mdot = b["bh_mdot"].data_ptr(); st.accdisc_mass = b["accdisc"].data_ptr()
        st.dust_mass = b["dust_mass"].data_ptr(); st.total_mass = b["total"].data_ptr(); st.injected_bh_energy = b["inj"].data_ptr()
        rho = T(d["SinkDensity"])
        st.sink_density = rho.data_ptr()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fp._chk(lib.gl_blackhole_accretion(ctx, None, 0, C.byref(st), 1, cnt))
        e1.record(stream)
        torch.cuda.synchronize()
        if rep:
            ms.append(e0.elapsed_time(e1))
    times["blackhole_accretion"] = dict(ms_min=min(ms), ms_median=float(np.median(ms)), reps=len(ms))
    rec["gpu"] = times
    rec["gpu_swallowed_gas_sinks_dust"] = [int(x) for x in cnt]
    rec["gpu_note"] = "B200, arrays resident, CUDA events on the library's stream around the whole C call (search kernels, sort, apply)"
    fp.close()

    if util.have_ref() and not a.skip_ref:
        ps2 = ics.disc_subgrid(a.ref_ngas, a.ref_ngas // 10, seed=42, clump_frac=a.clump_frac)
        ps2["timebin"][:] = 3 + (ps2["id"] % 3)
        ops = "T S D R G G D U H B Y Z"
        o = util.run_oracle(util.REF_HARNESS, ps2, par, ops)
        k = ops.split()
        rec["reference_cpu"] = dict(sample="%d gas + %d dust + 4 sinks, one core (oracle/_ref/ref_harness)" % (a.ref_ngas, a.ref_ngas // 10),
                                    blackhole_accretion_ms=1e3 * float(o["%d_B_time" % k.index("B")][0]),
                                    dust_density_ms=1e3 * float(o["%d_Y_time" % k.index("Y")][0]),
                                    dust_drag_ms=1e3 * float(o["%d_Z_time" % k.index("Z")][0]))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
