#!/usr/bin/env python
"""bench.py - one JSON line per run (contract in the task prompt; tier section (4)).

A "step" is one full force evaluation (accel.c:61-106): tree rebuild (keys, sort, build, moments) +
gravity walk with the relative criterion + density (converged-Hsml pass) + hmax refresh + hydro, on the
configuration BASELINE.json's target is quoted on: config 0, the 1.6M-gas protoplanetary disc + 2 sinks.
With --gpus N (torchrun) the disc is scaled to N x 1.6M gas particles (BASELINE config 2's "scaled disc",
weak scaling): each rank owns one Peano-Hilbert segment and the library exchanges the locally-essential tree over
NCCL (csrc/gl_let.cu; --exchange replicate selects the round-1 all-gather of every source instead).  At N > 1 the
line also carries `ngpu_equals_1gpu` (the N-rank results against one GPU on the same particles) and, unless
--no-strong, `strong_2p27`: BASELINE configs 3 and 4 at 2^27 particles on 1 GPU (rank 0) and on N GPUs.

  value  = gravity particle-node interactions per second of whole-step time, inputs resident in HBM
  e2e    = the same metric through the drop-in boundary with HOST P[]/SphP[] buffers
           (gl_gadget_pull -> gl_compute_accelerations -> gl_gadget_push), wall clock
  --impl reference : the reference's own CPU code (oracle/_ref/ref_harness when it was compiled, else the
           C restatement oracle/gl_oracle) on the host cores, one single-rank instance per core.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
    os.environ["NCCL_DEBUG"] = "NONE"      # VERSION and WARN print NCCL's banner on stdout; rank 0 prints ONE JSON line

METRIC = "gravity particle-node interactions/s per full force step (gravity+density+hydro)"
UNIT = "interactions/s"
N_GAS_PER_GPU = 1600000
CPU_SAMPLE_NGAS = 200000
FLOP_PER_INTERACTION = 25.0
DENS_BYTES_PER_UPDATE = 84.0    # SURVEY.md 8(d), compulsory HBM bytes per update per pass
HYDRO_BYTES_PER_UPDATE = 96.0


def peaks():
    p = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
    except Exception:
        pass
    out = dict(hbm_gbs=p.get("hbm_gbs", 6650.0), sm_max_mhz=p.get("sm_max_mhz", 1965.0),
               source="MEASURED_PEAKS.json" if p else "fallback (B200_PROFILING.md)")
    # non-tensor FMA peaks measured on this pool's B200 by scripts/probes/fma_peaks.cu (newest committed record)
    import glob
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_fma_peaks.json"))):
        try:
            q = json.load(open(f))
            out.update(fp32_tflops=float(q["fp32_fma_tflops"]), fp64_tflops=float(q["fp64_fma_tflops"]), fma_source=os.path.basename(f))
        except Exception:
            pass
    return out


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the newest committed ncu --set full capture (profiles/*_traffic.json,
    written by scripts/ncu_summaries.py from the 1.6M-disc step); None when no capture names the kernel."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json"))):
        try:
            t = json.load(open(f)).get("bytes_per_launch", {})
            if kernel in t:
                best = float(t[kernel])
        except Exception:
            pass
    return best


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        import numpy as np
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unavailable"])
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(self.rows))


# ----------------------------------------------------------------------------------------- reference arm ----
def oracle_binary():
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if os.path.exists(ref):
        return ref, "reference"
    import __graft_entry__ as ge
    ge.build_oracle()
    return os.path.join(ROOT, "oracle", "gl_oracle"), "port"


CONFIG0 = "BASELINE config 0: Pd_N16e5 protoplanetary disc, %d gas + 2 sinks"
SETTLE_OPS = "T S D R G G D U H C"      # cold start, BH pass, relative pass, settled density / hmax / hydro, pair counts


def run_cpu_instances(ninst, nsteps, ngas, seed0=42, keep_dump=False):
    """ninst concurrent single-rank instances; returns (interactions per step per instance, per-step seconds, kind,
    [dump of instance 0])."""
    import numpy as np
    import tempfile
    from gadget_leicester_b200 import ics
    from oracle import glio
    binary, kind = oracle_binary()
    td = tempfile.mkdtemp(prefix="glbench_")
    par = ics.default_params()
    procs = []
    ops = SETTLE_OPS + " A" * nsteps      # settle (cold start), then nsteps timed force evaluations
    nset = len(SETTLE_OPS.split())
    inputs = []
    for k in range(ninst):
        ps = ics.disc(ngas, seed=seed0 + k)
        fin, fout = os.path.join(td, "in%d.glin" % k), os.path.join(td, "out%d.glou" % k)
        glio.write_glin(fin, ps, par)
        inputs.append(ps)
        procs.append((subprocess.Popen([binary, fin, fout, ops], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL), fout))
    inter, secs, dump = [], [], None
    for k, (p, fout) in enumerate(procs):
        if p.wait() != 0:
            raise RuntimeError("CPU reference instance failed")
        o = glio.read_glou(fout)
        if k == 0 and keep_dump:
            dump = (inputs[0], par, o)
        for s_ in range(nsteps):
            pre = "%d_" % (nset + s_)
            t = sum(float(o[pre + op + "_time"][0]) for op in "GDUH")
            inter.append(float(o[pre + "G_interactions"][0]))
            secs.append(t)
    inter = np.array(inter).reshape(ninst, nsteps)
    secs = np.array(secs).reshape(ninst, nsteps)
    return inter, secs, kind, dump


def reference_arm(args):
    """The reference's own CPU code on the box's host cores: `--steps K --warmup W` force evaluations per instance
    (honoured as given), one single-rank instance per core (the image has no MPI), each on a bounded sample of the
    configuration our arm names (the 1.6M disc at full size is about two minutes per evaluation on one core)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = max(1, min(os.cpu_count() or 1, 32))
    nwarm, nsteps = max(0, args.warmup), max(1, args.steps)
    inter, secs, kind, _ = run_cpu_instances(cores, nwarm + nsteps, args.cpu_sample_ngas)
    inter, secs = inter[:, nwarm:], secs[:, nwarm:]
    step_s = float(secs.max(axis=0).mean())          # slowest instance per step
    value = float(inter.sum(axis=0).mean() / step_s)
    ngas = N_GAS_PER_GPU * max(1, args.gpus)
    sample = ("%d concurrent single-rank instances (no MPI in the image) of the same generator at %d gas + 2 sinks each "
              "(bounded sample of the workload); step = gravity_tree + density + force_update_hmax + hydro_force after a "
              "settled start" % (cores, args.cpu_sample_ngas))
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=nsteps, warmup=nwarm, ms_per_step=step_s * 1e3,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                config=dict(workload=CONFIG0 % ngas + ("" if args.gpus <= 1 else " (weak scaling, %d per GPU)" % N_GAS_PER_GPU),
                            sample=sample, instances=cores, particles_per_instance=args.cpu_sample_ngas + 2),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind=kind, sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                sph_updates_per_s=float(cores * args.cpu_sample_ngas / step_s), gpu_launches=0)
    print(json.dumps(line))


# ------------------------------------------------------------------------------ the 2^27-particle configs ----
def big_config(args):
    """BASELINE configs 3 (collisionless Plummer sphere, gravity walk only) and 4 (uniform random gas cube,
    density + hydro only), strong scaling: the SAME N-particle set for any GPU count, generated on the device."""
    import numpy as np
    import torch
    from gadget_leicester_b200 import ics
    from gadget_leicester_b200.multigpu import DistributedForcePath
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.nparts or (1 << 27)
    dev = torch.device("cuda", local)
    warm = max(3, args.warmup)
    if args.config == 3:
        par = ics.default_params(Softening=[1e-5, ics.plummer_softening(n), 5e-5, 1e-7, 0.01, 5e-5])
        ps = ics.plummer_torch(n, dev, seed=42)
        kw = dict(sph=False, gravity=True)
        metric, unit = "gravity particle-node interactions/s (tree rebuild + walk, relative criterion)", "interactions/s"
        work = "BASELINE config 3: collisionless Plummer sphere, %d particles, gravity tree walk only" % n
    else:
        par = ics.default_params(DesNumNgb=64, MaxNumNgbDeviation=0.5)
        ps = ics.cube_torch(n, dev, seed=42)
        ps["hsml"].fill_((3.0 * 64 / (4 * np.pi * n)) ** (1.0 / 3.0) * 1.3)
        kw = dict(sph=True, gravity=False)
        metric, unit = "SPH particle updates/s (tree rebuild + density + hydro)", "updates/s"
        work = "BASELINE config 4: uniform random SPH cube, %d gas particles, Ngb=64, density+hydro only" % n
    use_let = world > 1 and args.exchange == "let"
    if use_let:
        from gadget_leicester_b200.multigpu import LetForcePath
        dfp = LetForcePath(par, rank=rank, world=world, device=local)
    else:
        dfp = DistributedForcePath(par, rank=rank, world=world, device=local)
    dfp.decompose_device(ps)
    del ps
    torch.cuda.empty_cache()
    dfp.compute_accelerations(**kw)          # settle: Barnes-Hut pass / cold-start density
    cold_passes = dfp.fp.stats()["density_passes_max"]
    if world > 1 and args.config == 3 and not args.no_rebalance and not use_let:
        # domain.c:1075: the next decomposition balances the measured work (GravCost), not the particle count
        for _ in range(2):
            dfp.rebalance(timed=True)
            dfp.compute_accelerations(**kw)

    def one_step():
        dfp.fp.set_TreeReconstructFlag(1)
        dfp.compute_accelerations(**kw)

    for _ in range(warm):
        one_step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    dfp.fp.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot_ms, inter, phase = 0.0, 0, {}
    for _ in range(args.steps):
        torch.cuda.synchronize()
        e0.record(torch.cuda.current_stream())
        one_step()
        torch.cuda.synchronize()
        e1.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        tot_ms += e0.elapsed_time(e1)
        st = dfp.fp.stats()
        inter += st["interactions"] if args.config == 3 else 0
        for k in ("ms_keys_sort", "ms_treebuild", "ms_moments", "ms_gravity_walk", "ms_density", "ms_hmax", "ms_hydro"):
            phase[k] = phase.get(k, 0.0) + st[k] / args.steps
    launches = dfp.fp.stats()["kernel_launches"]
    clocks = sampler.stop()
    t = torch.tensor([tot_ms], dtype=torch.float64, device="cuda")
    it = torch.tensor([float(inter)], dtype=torch.float64, device="cuda")
    rank_phase = None
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(it, op=dist.ReduceOp.SUM)
        keys = ("ms_gravity_walk", "ms_density", "ms_hydro")
        mine = torch.tensor([phase[k] for k in keys] + [float(dfp.hi - dfp.lo)], dtype=torch.float64, device="cuda")
        allr = torch.empty(world * mine.numel(), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.view(world, -1).cpu().numpy()
        rank_phase = {k: [round(float(x), 2) for x in allr[:, j]] for j, k in enumerate(keys + ("targets",))}
    ms_per_step = float(t.item()) / args.steps
    units = (float(it.item()) / args.steps) if args.config == 3 else float(n)
    pk = peaks()
    if args.config == 3:
        nsm = torch.cuda.get_device_properties(local).multi_processor_count
        peak = nsm * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
        ach = FLOP_PER_INTERACTION * (inter / args.steps) / (phase["ms_gravity_walk"] * 1e-3) / 1e12
        roof = dict(kernel="k_gravity_walk", bound="fp32", achieved=ach, peak=peak, unit="TFLOP/s", frac=ach / peak, traffic=None,
                    note="25 flop/interaction; this rank's launch; peak = SMs*128*2*sm_max_mhz (" + pk["source"] + ")")
    else:
        own = dfp.hi - dfp.lo
        dom = "hydro" if phase["ms_hydro"] >= phase["ms_density"] else "density"
        per = HYDRO_BYTES_PER_UPDATE if dom == "hydro" else DENS_BYTES_PER_UPDATE
        ach = per * own / (phase["ms_" + dom] * 1e-3) / 1e9
        roof = dict(kernel="k_" + dom, bound="hbm", achieved=ach, peak=pk["hbm_gbs"], unit="GB/s", frac=ach / pk["hbm_gbs"],
                    traffic=None, note="%g B/update; this rank's launch; peak from %s" % (per, pk["source"]))
    if rank == 0:
        print(json.dumps(dict(metric=metric, value=units / (ms_per_step * 1e-3), unit=unit, n_gpus=world, steps=args.steps, warmup=warm,
                              ms_per_step=ms_per_step, higher_is_better=True, scaling="strong", vs_baseline=None,
                              dtype="f64 predicates + f32 force magnitude" if args.config == 3 else "f64", data="synthetic",
                              config=dict(workload=work, particles=n, parallelism="ph-segments x%d" % world + ("" if world == 1 else ", exchange=" + args.exchange),
                                          l2="inputs larger than L2 (%.1f GB of tree records)" % (n * 150e-9)),
                              phase_ms=phase, density_passes=dict(cold_start=cold_passes), gpu_launches=int(launches), clocks=clocks,
                              roofline=roof, e2e=None, interactions_per_step=(units if args.config == 3 else None),
                              rank_phase_ms=rank_phase)))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def resident_loop(ForcePath, par, ps, n, ksteps):
    """run.c:63-110 on the device-resident state (gl_run_begin / gl_run_steps): every step = TimeBinCount histogram ->
    next sync point -> drift of all particles -> tree rebuild + gravity + density + hmax + hydro for the active ones ->
    advance_and_find_timesteps.  MaxSizeTimestep is two ticks of the integer time line, so every particle stays on bin 1
    and every step is a full step (all particles active) - the same force work as `value`, plus drift and kick."""
    import time
    import numpy as np
    import torch
    tbi = par["Timebase_interval"]
    ps = dict(ps)
    ps["timebin"] = np.ones(n, dtype=np.int32)
    ps["ti_begstep"] = np.zeros(n, dtype=np.int32)
    fp = ForcePath(dict(par, ErrTolTheta=0.0, Ti_Current=2))
    fp.load(ps)
    fp.compute_accelerations()
    fp.run_begin(2)
    t = fp.timestep_params(2, [1, 1] + [0] * 27, Timebase_interval=tbi, ErrTolIntAccuracy=0.025, CourantFac=0.15, MaxSizeTimestep=2 * tbi,
                           MinSizeTimestep=0.0, MinEgySpec=0.0, SofteningTable=par["Softening"], OuterBoundary=1.4e10,
                           UnitVelocity_in_cm_per_s=2.98e6, FeedBackVelocity=1.0, SMBHmass=0.5, InnerBoundary=2.0e-2, SinkBoundary=1.0e-4,
                           AccDtBlackHole=0.007)
    fp.advance_and_find_timesteps(t)
    fp.run(t, 3)                                          # warm-up
    torch.cuda.synchronize()
    fp.reset_launch_count()
    t0 = time.perf_counter()
    info = fp.run(t, ksteps)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / ksteps
    launches = fp.stats()["kernel_launches"]
    st = fp.run_state()
    out = dict(value=info["interactions"] / ksteps / wall, unit=UNIT, ms_per_step=wall * 1e3, device_ms_per_step=info["ms_total"] / ksteps,
               steps=ksteps, active_per_step=info["active_total"] / ksteps, particles=n, Ti_Current=int(t.Ti_Current),
               bins_in_use=sorted(int(b) for b in np.unique(st["TimeBin"])), gpu_launches=int(launches),
               h2d_bytes_per_step=0, d2h_bytes_per_step=128 + 3 * 24 + 8 + 8,
               api="gl_run_steps: find_next_sync_point_and_drift + compute_accelerations + advance_and_find_timesteps per step on the "
                   "state in HBM; per step only TimeBinCount[32] and the phases' counters come back (wall clock, host loop included)")
    fp.close()
    return out


def _block_on(fp, arr, par, ksteps, dist=None, world=1):
    """gl_gadget_run_steps on structs that hold a complete, settled state (all ranks together when dist is given)."""
    import ctypes as C
    import time
    import numpy as np
    import torch
    from gadget_leicester_b200 import aos, lib as L
    tbi = par["Timebase_interval"]
    n, ng = arr.NumPart, arr.N_gas
    fp.set_params(Ti_Current=2)
    arr.pview("p_TimeBin", np.int16)[:] = 1
    arr.pview("p_Ti_begstep", np.int32)[:] = 0
    arr.pview("p_Ti_current", np.int32)[:] = 2
    t = fp.timestep_params(2, [1, 1] + [0] * 27, Timebase_interval=tbi, ErrTolIntAccuracy=0.025, CourantFac=0.15, MaxSizeTimestep=2 * tbi,
                           MinSizeTimestep=0.0, MinEgySpec=0.0, SofteningTable=par["Softening"], OuterBoundary=1.4e10,
                           UnitVelocity_in_cm_per_s=2.98e6, FeedBackVelocity=1.0, SMBHmass=0.5, InnerBoundary=2.0e-2, SinkBoundary=1.0e-4,
                           AccDtBlackHole=0.007)
    lib, lay = fp._lib, aos.layout_struct()
    act = np.arange(n, dtype=np.int32)
    err = C.create_string_buffer(512)
    if lib.gl_gadget_advance_and_find_timesteps(C.byref(lay), C.byref(t), L._ptr(arr.P), L._ptr(arr.SphP), L._ptr(act), n, None, err, len(err)):
        raise RuntimeError(err.value.decode())
    info = L.GLRunInfo()
    fp._chk(lib.gl_gadget_run_steps(fp.ctx, L._ptr(arr.P), L._ptr(arr.SphP), n, ng, C.byref(t), 2, C.byref(info)))      # warm-up
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    fp._chk(lib.gl_gadget_run_steps(fp.ctx, L._ptr(arr.P), L._ptr(arr.SphP), n, ng, C.byref(t), ksteps, C.byref(info)))
    wall = time.perf_counter() - t0
    h2d = n * (16 * 8 + 3 * 4) + ng * 16 * 8
    d2h = n * (14 * 8 + 4 + 2 * 4) + ng * 16 * 8
    inter, active = float(info.interactions), float(info.active_total)
    if dist is not None:
        tt = torch.tensor([wall], dtype=torch.float64, device="cuda")
        ii = torch.tensor([inter, active, float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ii, op=dist.ReduceOp.SUM)
        wall, (inter, active, h2d, d2h) = float(tt.item()), [float(x) for x in ii.tolist()]
    return dict(value=inter / wall, unit=UNIT, ms_per_step=wall * 1e3 / ksteps, ms_per_call=wall * 1e3, steps_per_call=ksteps,
                device_ms_per_step=info.ms_total / ksteps, active_per_step=active / ksteps,
                h2d_bytes_per_step=int(h2d) // ksteps, d2h_bytes_per_step=int(d2h) // ksteps, h2d_bytes_per_call=int(h2d), d2h_bytes_per_call=int(d2h),
                api="gl_gadget_run_steps on %s pinned host P[]/SphP[]: structs gathered + H2D once, %d iterations of run()'s loop on the "
                    "device, D2H + scatter once (all inside the timed region%s)"
                    % ("every rank's own" if world > 1 else "the", ksteps, "; max over ranks" if world > 1 else ""))


def resident_block(ForcePath, par, ps, n, ksteps):
    """The resident loop through the reference-facing struct call (gl_gadget_run_steps): host P[] / SphP[] (pinned) are
    read once, ksteps iterations of run()'s loop run on the device, the structs are written once - gather, H2D, D2H and
    scatter are inside the timed region, paid once per call instead of twice per step."""
    import numpy as np
    from gadget_leicester_b200 import aos
    ps = dict(ps)
    ps["timebin"] = np.ones(n, dtype=np.int32)
    ps["ti_begstep"] = np.zeros(n, dtype=np.int32)
    fp = ForcePath(dict(par, ErrTolTheta=0.0, Ti_Current=2))
    arr = aos.GadgetArrays(ps, pinned=True)
    drop = aos.GadgetDropIn(fp, arr)
    drop.compute_accelerations()                          # the structs now hold a complete, settled state
    out = _block_on(fp, arr, par, ksteps)
    fp.close()
    return out


def resident_loop_ranks(dfp, par, dist, world, ksteps):
    """resident_loop() at N ranks: every rank integrates its own Peano-Hilbert segment inside libgl_b200 (gl_run_steps in
    multi-rank mode: TimeBinCount all-gather, locally-essential exchange and force path, kick), max over ranks."""
    import time
    import torch
    k = dfp.hi - dfp.lo
    r = dfp.fetch()
    if "hsml" in r:
        dfp.own["hsml"][:k] = r["hsml"]
    dfp.own["oldacc"][:k] = r["oldacc"]
    dfp.own["timebin"][:k] = 1
    dfp.own["ti_begstep"][:k] = 0
    fp = dfp.fp
    fp.set_params(Ti_Current=2)
    dfp._load_own()
    fp.set_TreeReconstructFlag(1)
    dfp.compute_accelerations()
    tbi = par["Timebase_interval"]
    fp.run_begin(2)
    t = fp.timestep_params(2, [1, 1] + [0] * 27, Timebase_interval=tbi, ErrTolIntAccuracy=0.025, CourantFac=0.15, MaxSizeTimestep=2 * tbi,
                           MinSizeTimestep=0.0, MinEgySpec=0.0, SofteningTable=par["Softening"], OuterBoundary=1.4e10,
                           UnitVelocity_in_cm_per_s=2.98e6, FeedBackVelocity=1.0, SMBHmass=0.5, InnerBoundary=2.0e-2, SinkBoundary=1.0e-4,
                           AccDtBlackHole=0.007)
    fp.advance_and_find_timesteps(t)
    fp.run(t, 3)
    torch.cuda.synchronize()
    dist.barrier()
    ls0 = fp.let_stats()
    t0 = time.perf_counter()
    info = fp.run(t, ksteps)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / ksteps
    ls1 = fp.let_stats()
    st1 = fp.stats()
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    ii = torch.tensor([float(info["interactions"]) / ksteps, float(info["active_total"]) / ksteps], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dist.all_reduce(ii, op=dist.ReduceOp.SUM)
    return dict(value=float(ii[0].item()) / float(tt.item()), unit=UNIT, ms_per_step=float(tt.item()) * 1e3, steps=ksteps,
                active_per_step=float(ii[1].item()), Ti_Current=int(t.Ti_Current),
                rank0=dict(let_builds=ls1["builds"] - ls0["builds"], let_violations=ls1["violations"] - ls0["violations"],
                           imported_particles=ls1["imported_particles"], ms_own_tree=ls1["ms_own_tree"], ms_select=ls1["ms_select"],
                           ms_exchange=ls1["ms_exchange"], ms_sph_exchange=ls1["ms_sph_exchange"],
                           phase_ms={k: round(st1[k], 3) for k in ("ms_keys_sort", "ms_treebuild", "ms_moments", "ms_gravity_walk", "ms_density",
                                                                   "ms_hmax", "ms_hydro")}),
                h2d_bytes_per_step=0,
                d2h_bytes_per_step=world * (128 * world + 3 * 24 + 8 + 8 + 8 * world),
                api="gl_run_steps in multi-rank mode: every rank's own Peano-Hilbert segment stays in HBM; per step the "
                    "TimeBinCount[] all-gather, drift, locally-essential exchange over NCCL + force path, kick; max over ranks (wall clock)")


def identity_check(dfp, ps, par, n, rank, world, dist):
    """One more step on N ranks and the same step on ONE GPU (rank 0) from the same state (Hsml and OldAcc gathered from
    the ranks): interaction counts, pair counts and accelerations of every rank's own particles must be those of the
    single-GPU run - the merged locally-essential tree is the global tree wherever a target can look."""
    import numpy as np
    import torch
    from gadget_leicester_b200.gadget import ForcePath
    from gadget_leicester_b200.multigpu import gather_ragged
    r = dfp.fetch()
    k = dfp.hi - dfp.lo
    hs = gather_ragged(r["hsml"][:k].contiguous(), dfp.ranges)
    oa = gather_ragged(r["oldacc"][:k].contiguous(), dfp.ranges)
    # every rank: one more step from that state (its own in_hsml / in_oldacc already hold it)
    dfp.fp.set_TreeReconstructFlag(1)
    dfp.compute_accelerations()
    mine = dfp.fetch()
    lib, ctx = dfp.fp._lib, dfp.fp.ctx
    dp = torch.zeros(k, dtype=torch.int32, device="cuda")
    hp = torch.zeros(k, dtype=torch.int32, device="cuda")
    import ctypes as C
    dfp.fp._chk(lib.gl_get_pair_counts(ctx, C.c_void_p(dp.data_ptr()), C.c_void_p(hp.data_ptr()), 1))
    torch.cuda.synchronize()
    ref = torch.zeros(n, 6, dtype=torch.float64, device="cuda")      # gravcost, denspairs, hydropairs, acc[3]
    if rank == 0:
        order = dfp.order
        pso = {kk: (v[order] if hasattr(v, "shape") and v.shape[0] == n else v) for kk, v in ps.items()}
        pso["hsml"] = hs.cpu().numpy()
        pso["oldacc"] = oa.cpu().numpy()
        one = ForcePath(dict(par, ErrTolTheta=0.0))
        one.load(pso)
        g = one.gravity_tree()
        d = one.density()
        one.force_update_hmax()
        h = one.hydro_force()
        one.close()
        ref[:, 0] = torch.from_numpy(g["GravCost"].astype(np.float64)).cuda()
        ref[:, 1] = torch.from_numpy(d["DensPairs"].astype(np.float64)).cuda()
        ref[:, 2] = torch.from_numpy(h["HydroPairs"].astype(np.float64)).cuda()
        ref[:, 3:6] = torch.from_numpy(g["GravAccel"]).cuda()
    dist.broadcast(ref, src=0)
    own = ref[dfp.lo:dfp.hi]
    ty = torch.from_numpy(ps["type"][dfp.order[dfp.lo:dfp.hi]]).cuda()
    gas = ty == 0
    sphm = gas | (ty == 5) | (ty == 2)
    bad = torch.stack([(mine["gravcost"].double() != own[:, 0]).sum(),
                       ((dp.double() != own[:, 1]) & sphm).sum(),
                       ((hp.double() != own[:, 2]) & gas).sum(),
                       (mine["acc"] != own[:, 3:6]).any(dim=1).sum()]).double()
    dist.all_reduce(bad, op=dist.ReduceOp.SUM)
    b = [int(x) for x in bad.cpu().numpy()]
    return dict(particles=int(n), gravcost_mismatches=b[0], density_pair_mismatches=b[1], hydro_pair_mismatches=b[2],
                gravaccel_not_bitwise_equal=b[3], identical=bool(b[0] == 0 and b[1] == 0 and b[2] == 0),
                how="one step on %d ranks vs the same step on one GPU (rank 0) from the same Hsml / OldAcc state" % world)


def strong_scaling(args, rank, world, local, dist):
    """BASELINE configs 3 (Plummer sphere, gravity only) and 4 (uniform gas cube, density + hydro only) at 2^27
    particles: the same set on ONE GPU (rank 0, the others wait) and on all N GPUs; efficiency = t1 / (N * tN)."""
    import numpy as np
    import torch
    from gadget_leicester_b200 import ics
    from gadget_leicester_b200.multigpu import DistributedForcePath, LetForcePath
    n = args.strong_nparts or (1 << 27)
    dev = torch.device("cuda", local)
    out = dict(particles=n, steps=3)

    def setup(cfg):
        if cfg == 3:
            par = ics.default_params(Softening=[1e-5, ics.plummer_softening(n), 5e-5, 1e-7, 0.01, 5e-5])
            ps = ics.plummer_torch(n, dev, seed=42)
            kw = dict(sph=False, gravity=True)
        else:
            par = ics.default_params(DesNumNgb=64, MaxNumNgbDeviation=0.5)
            ps = ics.cube_torch(n, dev, seed=42)
            ps["hsml"].fill_((3.0 * 64 / (4 * np.pi * n)) ** (1.0 / 3.0) * 1.3)
            kw = dict(sph=True, gravity=False)
        return par, ps, kw

    def timed(d, kw, steps=3, warm=2):
        d.compute_accelerations(**kw)                  # settle: Barnes-Hut pass / cold-start density
        for _ in range(warm):
            d.fp.set_TreeReconstructFlag(1)
            d.compute_accelerations(**kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(steps):
            torch.cuda.synchronize()
            e0.record(torch.cuda.current_stream())
            d.fp.set_TreeReconstructFlag(1)
            d.compute_accelerations(**kw)
            torch.cuda.synchronize()
            e1.record(torch.cuda.current_stream())
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / steps

    for cfg in (3, 4):
        par, ps, kw = setup(cfg)
        t1 = torch.zeros(1, dtype=torch.float64, device=dev)
        if rank == 0:
            one = DistributedForcePath(par, rank=0, world=1, device=local)
            one.decompose_device(ps)
            t1[0] = timed(one, kw)
            one.close()
            del one
            torch.cuda.empty_cache()
        dist.barrier()
        dist.broadcast(t1, src=0)
        d = LetForcePath(par, rank=rank, world=world, device=local)
        d.decompose_device(ps)
        del ps
        torch.cuda.empty_cache()
        tn = torch.tensor([timed(d, kw)], dtype=torch.float64, device=dev)
        dist.all_reduce(tn, op=dist.ReduceOp.MAX)
        ls = d.fp.let_stats()
        imp = torch.tensor([float(ls["imported_particles"]), float(ls["violations"])], dtype=torch.float64, device=dev)
        dist.all_reduce(imp, op=dist.ReduceOp.MAX)
        d.close()
        del d
        torch.cuda.empty_cache()
        out["config%d_ms_1gpu" % cfg] = float(t1.item())
        out["config%d_ms" % cfg] = float(tn.item())
        out["config%d_efficiency" % cfg] = float(t1.item() / (world * tn.item()))
        out["config%d_imported_particles_max" % cfg] = float(imp[0].item())
        out["config%d_let_violations" % cfg] = float(imp[1].item())
    out["efficiency"] = min(out["config3_efficiency"], out["config4_efficiency"])
    out["note"] = ("same 2^27-particle set on 1 GPU (rank 0) and on %d GPUs in this run; step = tree build (+ locally-essential "
                   "exchange) + gravity walk (config 3) or density + hydro (config 4); efficiency = t1 / (N tN), the smaller of the two" % world)
    return out


# ----------------------------------------------------------------------------------------------- our arm ----
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ngas", type=int, default=0, help="override gas particles per GPU (default 1.6M)")
    ap.add_argument("--config", type=int, default=0, choices=[0, 1, 2, 3, 4],
                    help="0: 1.6M disc (default, headline); 1: Evrard sphere, 1M gas; 2: 16M disc; 3: Plummer sphere 2^27, gravity "
                         "only; 4: uniform gas cube 2^27, density+hydro only")
    ap.add_argument("--nparts", type=int, default=0, help="total particles for --config 3/4 (default 2^27)")
    ap.add_argument("--rebalance", action="store_true", help="config 0: re-split the Peano-Hilbert segments by the settle step's GravCost")
    ap.add_argument("--no-rebalance", action="store_true", help="keep the equal-count Peano-Hilbert split (default: work-balanced "
                    "split from the settle step's GravCost, domain.c:1075)")
    ap.add_argument("--cpu-sample-ngas", type=int, default=CPU_SAMPLE_NGAS,
                    help="gas particles per CPU instance of the reference arm / cpu_baseline leg (the line states the sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="let", choices=["let", "replicate"],
                    help="N > 1: 'let' = locally-essential tree exchanged inside the library over NCCL (default); 'replicate' = the "
                         "round-1 path, every source all-gathered and the global tree built on every GPU")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the 2^27-particle strong-scaling lines (configs 3 and 4)")
    ap.add_argument("--strong-nparts", type=int, default=0, help="particles of the strong-scaling sets (default 2^27)")
    ap.add_argument("--no-identity", action="store_true", help="N > 1: skip the N-GPU == 1-GPU comparison")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    if args.config in (3, 4):
        return big_config(args)

    import numpy as np
    import torch
    from gadget_leicester_b200 import ics
    from gadget_leicester_b200.multigpu import DistributedForcePath
    from gadget_leicester_b200.gadget import ForcePath
    from gadget_leicester_b200 import aos

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warm = max(3, args.warmup)
    par = ics.default_params()
    if args.config == 1:
        # BASELINE config 1: Evrard collapse, 1M gas, gravitational softening of the gas = Hsml
        ngas = (args.ngas or 1000000) * world
        par = ics.default_params(Softening=[1e-4, 5e-5, 5e-5, 1e-7, 0.01, 5e-5])
        ps = ics.evrard(ngas, seed=42)
        workload = "BASELINE config 1: Evrard adiabatic gas-sphere collapse, %d gas particles" % ngas
    else:
        ngas = (args.ngas or (16000000 if args.config == 2 else N_GAS_PER_GPU)) * world
        ps = ics.disc(ngas, seed=42)
        workload = ("BASELINE config 2: scaled self-gravitating disc, %d gas + 2 sinks" if args.config == 2 else
                    "BASELINE config 0: Pd_N16e5 protoplanetary disc, %d gas + 2 sinks") % ngas
    ps["hsml"] = ics.guess_hsml(ps, par["DesNumNgb"])
    ps["hsml"][ps["type"] == 5] = par["Softening"][5]
    n = ps["pos"].shape[0]

    use_let = world > 1 and args.exchange == "let"
    if use_let:
        from gadget_leicester_b200.multigpu import LetForcePath
        dfp = LetForcePath(par, rank=rank, world=world, device=local)
    else:
        dfp = DistributedForcePath(par, rank=rank, world=world, device=local)
    dfp.decompose(ps)
    # settle: Barnes-Hut pass (OldAcc=0) then cold-start density; afterwards every step is a steady-state step
    dfp.compute_accelerations()
    cold_passes = dfp.fp.stats()["density_passes_max"]
    if world > 1 and args.rebalance and not use_let:
        # domain.c:1075 balances the measured GRAVITY work; two thirds of this step is SPH work, which equal counts
        # balance better on the disc (measured at N=2: 18.3 ms equal counts, 18.7 ms GravCost-balanced) - so off by default
        for _ in range(2):
            dfp.rebalance(timed=True)
            dfp.compute_accelerations()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def one_step():
        dfp.fp.set_TreeReconstructFlag(1)
        dfp.compute_accelerations()

    for _ in range(warm):
        one_step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    dfp.fp.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot_ms, acc = 0.0, dict(ms_keys_sort=0, ms_treebuild=0, ms_moments=0, ms_gravity_walk=0, ms_density=0, ms_hmax=0, ms_hydro=0)
    inter = 0
    passes = []
    wall = {}
    for _ in range(args.steps):
        flush.fill_(1)                         # L2 flush between timed iterations (untimed)
        torch.cuda.synchronize()
        e0.record(torch.cuda.current_stream())
        one_step()
        torch.cuda.synchronize()
        e1.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        tot_ms += e0.elapsed_time(e1)
        st = dfp.fp.stats()
        for k in acc:
            acc[k] += st[k]
        inter += st["interactions"]
        passes.append(st["density_passes_max"])
        for k, v in dfp.t.items():
            wall[k] = wall.get(k, 0.0) + v * 1e3 / args.steps
    launches = dfp.fp.stats()["kernel_launches"]
    clocks = sampler.stop()
    t = torch.tensor([tot_ms], dtype=torch.float64, device="cuda")
    it = torch.tensor([float(inter)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(it, op=dist.ReduceOp.SUM)
    ms_per_step = float(t.item()) / args.steps
    inter_per_step = float(it.item()) / args.steps
    value = inter_per_step / (ms_per_step * 1e-3)
    phase = {k: v / args.steps for k, v in acc.items()}
    rank_phase = None
    let_line = None
    if dist is not None:
        keys = ("ms_keys_sort", "ms_moments", "ms_gravity_walk", "ms_density", "ms_hydro")
        extra = [wall.get("allgather_state", 0.0), wall.get("allgather_sph", 0.0), float(dfp.hi - dfp.lo)]
        names = ("wall_allgather_state", "wall_allgather_sph", "targets")
        if use_let:
            ls = dfp.fp.let_stats()
            extra = [ls["ms_own_tree"], ls["ms_select"], ls["ms_exchange"], ls["ms_sph_exchange"], float(ls["imported_particles"]),
                     float(ls["imported_cells"]), float(ls["exported_particles"]), float(ls["violations"]), float(dfp.hi - dfp.lo)]
            names = ("let_ms_own_tree", "let_ms_select", "let_ms_exchange", "let_ms_sph_exchange", "let_imported_particles",
                     "let_imported_cells", "let_exported_particles", "let_violations", "targets")
        mine = torch.tensor([phase[k] for k in keys] + extra, dtype=torch.float64, device="cuda")
        allr = torch.empty(world * mine.numel(), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.view(world, -1).cpu().numpy()
        rank_phase = {k: [round(float(x), 2) for x in allr[:, j]] for j, k in enumerate(keys + names)}
        if use_let:
            imp = allr[:, len(keys) + 4]
            let_line = dict(exchange="locally-essential tree inside libgl_b200 (NCCL all-gather of target boxes + two all-to-alls)",
                            imported_particles_mean=float(imp.mean()), imported_particles_max=float(imp.max()),
                            replicated_would_import=float(n - (dfp.hi - dfp.lo)),
                            bytes_received_per_rank_max=float(imp.max() * 80 + allr[:, len(keys) + 5].max() * 64))

    # ---- N ranks against ONE GPU on the same particles and the same state (north_star: identical results) ----
    identity = None
    if dist is not None and use_let and not args.no_identity:
        identity = identity_check(dfp, ps, par, n, rank, world, dist)

    # ---- roofline of the dominant kernel (per-launch durations from the library's CUDA events) ----
    pk = peaks()
    own_gas = int((ps["type"][dfp.order[dfp.lo:dfp.hi]] == 0).sum())
    nsm = torch.cuda.get_device_properties(local).multi_processor_count
    fp32_peak = pk.get("fp32_tflops") or nsm * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    fp64_peak = pk.get("fp64_tflops") or nsm * 64 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12          # 64 DFMA lanes per SM on B200
    peak_note = ("measured FFMA / DFMA peak, profiles/" + pk["fma_source"]) if pk.get("fma_source") else \
        "SMs*128*2*sm_max_mhz (derived; no profiles/*_fma_peaks.json)"
    st = dfp.fp.stats()                                                 # pair / pass totals of the last step (this rank)
    t_w, t_d, t_h = phase["ms_gravity_walk"] * 1e-3, phase["ms_density"] * 1e-3, phase["ms_hydro"] * 1e-3
    tr = (lambda k: ncu_traffic(k) if world == 1 and not args.ngas and args.config == 0 else None)
    ach = FLOP_PER_INTERACTION * (inter / args.steps) / t_w / 1e12
    roofs = dict(
        gravity_walk=dict(kernel="k_gravity_walk", bound="fp32", achieved=ach, peak=fp32_peak, unit="TFLOP/s", frac=ach / fp32_peak,
                          traffic=tr("k_gravity_walk"),
                          note="25 flop/interaction (SURVEY 8(d)) x %d interactions; peak = %s" % (inter // args.steps, peak_note)))
    for nm, t, per, units, pairs, fpp, bpp in (("density", t_d, DENS_BYTES_PER_UPDATE, st["density_target_passes"], st["density_pairs"], 68.0, 36.0),
                                               ("hydro", t_h, HYDRO_BYTES_PER_UPDATE, own_gas, st["hydro_pairs"], 98.0, 60.0)):
        ach = per * units / t / 1e9
        roofs[nm] = dict(kernel="k_" + nm, bound="hbm", achieved=ach, peak=pk["hbm_gbs"], unit="GB/s", frac=ach / pk["hbm_gbs"],
                         traffic=tr("k_" + nm), algorithmic_bytes=per * units,
                         pair_work=dict(pairs=int(pairs), flop_per_pair=fpp, fp64_tflops=fpp * pairs / t / 1e12,
                                        fp64_peak_tflops=fp64_peak, frac_fp64=fpp * pairs / t / 1e12 / fp64_peak,
                                        gather_bytes_per_pair=bpp, gather_gbs=bpp * pairs / t / 1e9),
                         note="%g B/update x %d target evaluations (SURVEY 8(d) compulsory HBM bytes; %s passes of the Hsml "
                              "iteration counted per target); peak from %s; traffic = ncu dram bytes per launch of the same "
                              "workload (profiles/); pair_work = the L1/L2-served pair arithmetic that actually bounds the kernel"
                              % (per, units, "all" if nm == "density" else "no", pk["source"]))
    cand = dict(gravity_walk=t_w, density=t_d, hydro=t_h)
    dom = max(cand, key=cand.get)
    roof = dict(roofs[dom], others={k: v for k, v in roofs.items() if k != dom})

    # ---- run()'s loop with every rank's state resident in HBM (the N-rank form of resident_loop below) ----
    resident = None
    if dist is not None and use_let and not args.no_e2e:
        try:
            resident = resident_loop_ranks(dfp, par, dist, world, max(2, min(args.steps, 10)))
        except Exception as ex:  # collective: a failure here is a failure on every rank
            resident = dict(error=str(ex)[:300])

    # ---- e2e: the same step through the drop-in boundary with host P[]/SphP[] (rank-local set) ----
    e2e = None
    if not args.no_e2e and world > 1 and use_let:
        # every rank owns the host P[]/SphP[] of its Peano-Hilbert segment, like every MPI rank of the reference, and
        # calls the drop-in entry point on it; the exchange happens inside the library
        r = dfp.fetch()
        seg = dfp.order[dfp.lo:dfp.hi]
        ps_seg = {k: (v[seg] if hasattr(v, "shape") and v.shape[0] == n else v) for k, v in ps.items()}
        ps_seg["hsml"] = r["hsml"].cpu().numpy()
        ps_seg["oldacc"] = r["oldacc"].cpu().numpy()
        dfp.fp.set_params(ErrTolTheta=0.0)
        arr = aos.GadgetArrays(ps_seg, pinned=True)
        drop = aos.GadgetDropIn(dfp.fp, arr)
        for _ in range(2):
            drop.compute_accelerations()
        ksteps = max(2, min(args.steps, 5))
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        ie = 0
        for _ in range(ksteps):
            drop.compute_accelerations()
            ie += dfp.fp.stats()["interactions"]
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / ksteps
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        ii = torch.tensor([float(ie) / ksteps], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ii, op=dist.ReduceOp.SUM)
        h2d, d2h = drop.bytes_per_step()
        e2e = dict(value=float(ii.item()) / float(tt.item()), unit=UNIT, h2d_bytes_per_step=h2d * world, d2h_bytes_per_step=d2h * world,
                   ms_per_step=float(tt.item()) * 1e3, steps=ksteps,
                   api="gl_gadget_compute_accelerations on every rank's own pinned host P[]/SphP[] segment (pack + H2D, "
                       "locally-essential exchange over NCCL, force path, D2H + scatter); max over ranks")
        if resident is not None and "error" not in resident:
            try:
                resident["through_structs"] = _block_on(dfp.fp, arr, par, 10, dist=dist, world=world)
            except Exception as ex:  # collective: a failure here is a failure on every rank
                resident["through_structs"] = dict(error=str(ex)[:300])
    if not args.no_e2e and world > 1 and not use_let:
        # every rank owns the host P[]/SphP[] of its Peano-Hilbert segment, like every MPI rank of the reference
        from gadget_leicester_b200.multigpu import DistributedDropIn
        seg = dfp.order[dfp.lo:dfp.hi]
        ps_seg = {k: (v[seg] if hasattr(v, "shape") and v.shape[0] == n else v) for k, v in ps.items()}
        ps_seg["hsml"] = dfp.res["hsml"][dfp.lo:dfp.hi].cpu().numpy()
        ps_seg["oldacc"] = dfp.res["oldacc"][dfp.lo:dfp.hi].cpu().numpy()
        dfp.fp.set_params(ErrTolTheta=0.0)
        arr = aos.GadgetArrays(ps_seg, pinned=True)
        drop = DistributedDropIn(dfp, arr)
        for _ in range(2):
            dfp.fp.set_TreeReconstructFlag(1)
            drop.compute_accelerations()
        ksteps = max(2, min(args.steps, 5))
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        ie = 0
        for _ in range(ksteps):
            dfp.fp.set_TreeReconstructFlag(1)
            drop.compute_accelerations()
            ie += dfp.fp.stats()["interactions"]
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / ksteps
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        ii = torch.tensor([float(ie) / ksteps], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ii, op=dist.ReduceOp.SUM)
        h2d, d2h = drop.bytes_per_step()
        e2e = dict(value=float(ii.item()) / float(tt.item()), unit=UNIT, h2d_bytes_per_step=h2d * world, d2h_bytes_per_step=d2h * world,
                   ms_per_step=float(tt.item()) * 1e3, steps=ksteps,
                   api="DistributedDropIn.compute_accelerations: every rank's own pinned host P[]/SphP[] segment "
                       "(gl_gadget_pack_segment + H2D, all-gathers, force path, async D2H + scatter); max over ranks")
    if not args.no_e2e and world == 1:
        res_h = dfp.res["hsml"].cpu().numpy()
        res_o = dfp.res["oldacc"].cpu().numpy()
        ps2 = {k: (v[dfp.order] if hasattr(v, "shape") and v.shape[0] == n else v) for k, v in ps.items()}
        ps2["hsml"], ps2["oldacc"] = res_h, res_o
        dfp.close()
        del dfp
        torch.cuda.empty_cache()
        fp = ForcePath(dict(par, ErrTolTheta=0.0))       # steady state: relative criterion (gravtree.c:396-397)
        arr = aos.GadgetArrays(ps2, pinned=True)
        drop = aos.GadgetDropIn(fp, arr)
        for _ in range(2):
            fp.set_TreeReconstructFlag(1)
            drop.compute_accelerations()
        torch.cuda.synchronize()
        ksteps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        ie = 0
        for _ in range(ksteps):
            fp.set_TreeReconstructFlag(1)
            drop.compute_accelerations()
            ie += fp.stats()["interactions"]
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / ksteps
        h2d, d2h = drop.bytes_per_step()
        e2e = dict(value=(ie / ksteps) / dt, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=dt * 1e3,
                   steps=ksteps, api="gl_gadget_compute_accelerations on pinned host P[]/SphP[] (pack + H2D, force path, pipelined D2H + scatter)")
        fp.close()

    # ---- run()'s loop with the state resident in HBM: drift + force path + kick per step, nothing but TimeBinCount[] crosses PCIe ----
    if not args.no_e2e and world == 1:
        try:
            resident = resident_loop(ForcePath, par, ps2, n, max(2, min(args.steps, 10)))
        except Exception as ex:  # an extra record, never a reason to lose the bench line
            resident = dict(error=str(ex)[:300])
        try:
            resident["through_structs"] = resident_block(ForcePath, par, ps2, n, 10)
        except Exception as ex:
            resident["through_structs"] = dict(error=str(ex)[:300])

    # ---- strong scaling on the 2^27-particle sets (north_star: >= 70 % at 8 GPUs), part of the N > 1 line ----
    strong = None
    if dist is not None and use_let and not args.no_strong and args.config == 0:
        dfp.close()
        del dfp
        torch.cuda.empty_cache()
        strong = strong_scaling(args, rank, world, local, dist)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    cpu_base, parity = None, None
    if world == 1 and not args.no_cpu_baseline and args.config == 0:
        try:
            ci, cs, kind, dump = run_cpu_instances(1, 2, args.cpu_sample_ngas, keep_dump=True)
            cpu_base = dict(value=float(ci[0, 1] / cs[0, 1]), unit=UNIT, cores=1, kind=kind,
                            sample="1 single-rank instance, %d-gas disc + 2 sinks (same generator, bounded sample), "
                                   "one settled force step: %.2f s" % (args.cpu_sample_ngas, cs[0, 1]),
                            sph_updates_per_s=float(args.cpu_sample_ngas / cs[0, 1]))
            if kind == "reference":
                # the same run is the checker for a parity record at this size (tests/parity_bench.py): the CUDA path
                # on the sample's particles against the reference's dump
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                import parity_bench
                psd, pard, out = dump
                npd, ngd = psd["pos"].shape[0], int((psd["type"] == 0).sum())
                parity = parity_bench.parity_record(ForcePath, psd, pard, parity_bench.golden_layout(out, npd, ngd))
        except Exception as ex:  # the baseline is a reported number, never a reason to lose the bench line
            if cpu_base is None:
                cpu_base = dict(value=None, unit=UNIT, cores=1, kind="unavailable", sample=str(ex)[:200])
            else:
                parity = dict(ok=None, error=str(ex)[:300])

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=warm, ms_per_step=ms_per_step,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64 predicates + f32 force magnitude (gravity), f64 (SPH)",
                data="synthetic",
                config=dict(workload=workload + ("" if world == 1 else " (weak scaling, %d per GPU)" % (ngas // world)),
                            particles=n, parallelism="ph-segments x%d" % world + ("" if world == 1 else ", exchange=" + args.exchange),
                            l2="256 MiB flush between timed steps",
                            step="tree rebuild + gravity (relative criterion) + density + hmax + hydro"),
                interactions_per_step=inter_per_step, sph_updates_per_s=(ngas / (ms_per_step * 1e-3)),
                force_eval_ms=ms_per_step, phase_ms=phase, density_passes=dict(cold_start=cold_passes, steady=passes),
                gpu_launches=int(launches), clocks=clocks, roofline=roof,
                wall_phase_ms={k: round(v, 3) for k, v in wall.items()}, rank_phase_ms=rank_phase)
    if e2e is not None:
        line["e2e"] = e2e
    if resident is not None:
        line["resident_loop"] = resident
    if cpu_base is not None:
        line["cpu_baseline"] = cpu_base
    if parity is not None:
        line["parity_at_bench"] = parity
    if let_line is not None:
        line["exchange"] = let_line
    if identity is not None:
        line["ngpu_equals_1gpu"] = identity
    if strong is not None:
        line["strong_2p27"] = strong
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
