"""CPU suite: host-side logic - IC generators, the GLIN/GLOU containers, the AoS layout, and the N>1
partition + all-gather plumbing on the gloo backend (world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gadget_leicester_b200 import ics, aos
from gadget_leicester_b200.multigpu import (partition_range, gather_padded, gather_ragged, work_balanced_ranges, packed_views,
                                            unpack_gathered, allgather_block, F64_ROWS, I32_ROWS)
from oracle import glio


def test_generators_are_deterministic_and_gas_first():
    for gen, n in ((ics.disc, 500), (ics.evrard, 300), (ics.cube, 300), (ics.plummer, 300)):
        a, b = gen(n, seed=9), gen(n, seed=9)
        assert (a["pos"] == b["pos"]).all() and (a["vel"] == b["vel"]).all()
        ng = int((a["type"] == 0).sum())
        assert (a["type"][:ng] == 0).all() and (a["type"][ng:] != 0).all()
        assert len(set(map(tuple, a["pos"]))) == a["pos"].shape[0]  # no coincident points (forcetree.c:208-232)
    d = ics.disc(500)
    assert d["pos"].shape[0] == 502 and (d["type"][-2:] == 5).all()


def test_glin_roundtrip(tmp_path):
    ps, par = ics.disc(100, seed=2), ics.default_params()
    p = tmp_path / "x.glin"
    glio.write_glin(str(p), ps, par)
    raw = open(p, "rb").read()
    assert raw[:4] == b"GLIN"
    n = int(np.frombuffer(raw[8:16], dtype=np.int64)[0])
    assert n == 102
    parr = np.frombuffer(raw[24:24 + 8 * glio.NPAR], dtype=np.float64)
    assert parr[glio.P_DESNUMNGB] == 40 and parr[glio.P_SOFT0 + 5] == 5e-5
    pos = np.frombuffer(raw[24 + 8 * glio.NPAR:24 + 8 * glio.NPAR + 24 * n], dtype=np.float64).reshape(n, 3)
    assert (pos == ps["pos"]).all()


def test_aos_layout_matches_reference_struct_sizes():
    lay = aos.layout()
    assert lay["p_stride"] == 536 and lay["sph_stride"] == 264
    assert lay["p_GravCost"] == 520 and lay["p_Type"] == 532 and lay["s_HydroAccel"] == 96
    ps = ics.disc(64, seed=3)
    ps["hsml"][:] = 0.1
    arr = aos.GadgetArrays(ps, pinned=False)
    assert (arr.pview("p_Pos", np.float64, 3) == ps["pos"]).all()
    assert (arr.pview("p_Type", np.int16) == ps["type"]).all()
    assert (arr.sview("s_VelPred", np.float64, 3) == ps["vel"][:64]).all()
    r = arr.results()
    assert r["Hsml"][0] == 0.1 and r["Density"][0] == -1.0


def test_partition_covers_everything():
    for n in (1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            segs = [partition_range(n, world, r) for r in range(world)]
            assert segs[0][0] == 0 and segs[-1][1] == n
            for a, b in zip(segs, segs[1:]):
                assert a[1] == b[0]
            assert all(hi - lo <= m for lo, hi, m in segs)


def test_work_balanced_split_equalises_cost():
    """domain_findSplit_work_balanced (domain.c:1075): contiguous segments of (nearly) equal summed cost."""
    rng = np.random.default_rng(4)
    gravcost = rng.integers(50, 4000, size=100003).astype(np.float64)
    timebin = rng.integers(0, 3, size=gravcost.size)
    cost = (1 + gravcost) / 2.0 ** timebin            # domain.c:378-384
    for world in (1, 2, 3, 8):
        rg = work_balanced_ranges(cost, world)
        assert rg[0][0] == 0 and rg[-1][1] == cost.size and all(a[1] == b[0] for a, b in zip(rg, rg[1:]))
        seg = np.array([cost[a:b].sum() for a, b in rg])
        assert seg.max() - seg.min() <= 2 * cost.max()
    # equal weights reproduce the equal-count split of the first step (GravCost = 1, init.c:252)
    rg = work_balanced_ranges(np.ones(1000), 4)
    assert [b - a for a, b in rg] == [250, 250, 250, 250]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi, m = partition_range(n, world, rank)
    full = torch.arange(n * 3, dtype=torch.float64).reshape(n, 3)
    got = gather_padded(full[lo:hi].clone(), n, m)
    ids = gather_padded(torch.arange(lo, hi, dtype=torch.int32), n, m)
    ok = bool((got == full).all()) and bool((ids == torch.arange(n, dtype=torch.int32)).all())
    # ragged (work-balanced) segments
    cut = max(1, n // 3)
    ranges = [(0, cut), (cut, n)]
    a, b = ranges[rank]
    got2 = gather_ragged(full[a:b].clone(), ranges)
    ok = ok and bool((got2 == full).all())
    # the packed state blocks of DistributedForcePath.exchange_state: one collective per dtype, ragged segments
    mm = max(b2 - a2 for a2, b2 in ranges)
    f64 = torch.zeros(F64_ROWS * mm, dtype=torch.float64)
    i32 = torch.zeros(I32_ROWS * mm, dtype=torch.int32)
    v = packed_views(f64, i32, mm)
    g = torch.arange(a, b, dtype=torch.float64)
    v["pos"][: b - a] = full[a:b]
    v["vel"][: b - a] = -full[a:b]
    for k, name in enumerate(("mass", "hsml", "oldacc", "entropy", "dtentropy")):
        v[name][: b - a] = g * 10 + k
    for k, name in enumerate(("type", "timebin", "ti_begstep")):
        v[name][: b - a] = (g * 4 + k).to(torch.int32)
    out = unpack_gathered(allgather_block(f64, world), allgather_block(i32, world), ranges, mm)
    gg = torch.arange(n, dtype=torch.float64)
    ok = ok and bool((out["pos"] == full).all()) and bool((out["vel"] == -full).all())
    ok = ok and all(bool((out[name] == gg * 10 + k).all()) for k, name in enumerate(("mass", "hsml", "oldacc", "entropy", "dtentropy")))
    ok = ok and all(bool((out[name] == (gg * 4 + k).to(torch.int32)).all()) for k, name in enumerate(("type", "timebin", "ti_begstep")))
    # the "unconverged targets remaining" reduction of density.c:656 in NCCL/gloo form
    left = torch.tensor([rank + 1], dtype=torch.int64)
    dist.all_reduce(left)
    ok = ok and int(left.item()) == world * (world + 1) // 2
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 11, 1001])
def test_allgather_plumbing_gloo_world2(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


@pytest.mark.parametrize("sanitizer", ["", "thread"])
def test_host_worker_pool_native(tmp_path, sanitizer):
    """gl_host.h (the worker pool behind the drop-in layer's pack / scatter / drift / reorder loops) as plain C++:
    chunk coverage, FIFO hand-out, non-blocking submit, concurrent submitters (tests/native/hostpool_check.cpp) -
    once optimised, once under ThreadSanitizer (any reported race fails the run)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "hostpool_check")
    flags = ["-O1", "-g", "-fsanitize=" + sanitizer] if sanitizer else ["-O2"]
    cc = subprocess.run(["g++"] + flags + ["-std=c++17", "-pthread", "-I", os.path.join(root, "gadget_leicester_b200", "csrc"),
                         "-o", exe, os.path.join(root, "tests", "native", "hostpool_check.cpp")], capture_output=True, text=True)
    if cc.returncode != 0 and sanitizer:
        pytest.skip("no lib%ssan for this g++" % sanitizer[0])
    assert cc.returncode == 0, cc.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, GL_HOST_THREADS="6"))
    assert r.returncode == 0 and "HOSTPOOL OK" in r.stdout, r.stdout + r.stderr


def test_timestep_params_mirror_of_the_reference_globals():
    """ForcePath.timestep_params (host mirror of gl_timestep_params): All.* names in, the struct the resident loop and the
    host kick read out - no GPU involved."""
    from gadget_leicester_b200.gadget import ForcePath
    t = ForcePath.timestep_params(64, [1 if b <= 6 else 0 for b in range(29)], Timebase_interval=4.0e-5, ErrTolIntAccuracy=0.025,
                                  CourantFac=0.15, MaxSizeTimestep=8.0e-3, MinSizeTimestep=0.0, MinEgySpec=1.8,
                                  SofteningTable=[1e-5, 0.01, 5e-5, 1e-7, 0.01, 5e-5], SMBHmass=0.5)
    assert t.Ti_Current == 64 and list(t.TimeBinActive)[:8] == [1] * 7 + [0] and sum(t.TimeBinActive) == 7
    assert t.Timebase_interval == 4.0e-5 and t.MaxSizeTimestep == 8.0e-3 and t.MinEgySpec == 1.8
    assert abs(t.GammaMinus1 - 0.4) < 1e-15                                   # GAMMA = 7/5 (allvars.h:64)
    assert t.SofteningTable[5] == 5e-5 and t.SMBHmass == 0.5
    assert t.p_DeltaDustMomentum == -1 and t.p_NewDensity == -1              # SoA state: no struct offsets
