"""GPU suite: the multi-rank force path (include/gl_b200.h "multi-rank mode", csrc/gl_let.cu) against a single-context
run of the union of all ranks' particles - through the C ABI, several contexts of this process as ranks (in-process
group: one GPU is enough), plus the NCCL transport on two GPUs when the box has them.

What must hold (north_star: "results must match ... on the same particles, tree and opening criterion"):
  * P.GravCost of every rank's own particles, Barnes-Hut and relative criterion: bit-exact, and GravAccel BITWISE equal
    (the merged locally-essential tree is the global tree wherever a target can look, and the walk is the same walk);
  * density and hydro pair counts: bit-exact; Hsml, Density, HydroAccel, DtEntropy to 1e-12 (FP64 sums formed in a
    different candidate order);
  * a cut cell that a walk wants to open / a search sphere reaches is detected and the phase is repeated with a full
    export on all ranks (cold-start density of the disc: the sinks' Hsml grows a thousandfold).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

pytestmark = pytest.mark.gpu


def _run(case, n, world):
    import let_check as LC
    from gadget_leicester_b200.gadget import ForcePath
    ps, par = LC.make_case(case, n)
    ngas = int((ps["type"] == 0).sum())
    sph = ngas > 0
    one = ForcePath(par)
    one.load(ps)
    order = one.peano_hilbert_order(ngas)
    one.load(ps)

    def reload_one(out):
        one.load(dict(ps, hsml=out["d"]["Hsml"].copy(), oldacc=out["g2"]["OldAcc"].copy()))
    ref = LC.sequence(one, sph, reload_one)
    one.close()
    res, stats, _ = LC.run_world(ps, par, order, world, sph)
    return LC.compare(ref, res, sph, ps["type"]), stats, ps


@pytest.mark.parametrize("case,n,world", [("plummer", 30000, 2), ("disc", 40000, 4), ("evrard", 60000, 3), ("cube", 50000, 8)])
def test_ranks_reproduce_the_single_context_run(case, n, world):
    rep, stats, ps = _run(case, n, world)
    for k, v in rep.items():
        if k.endswith("mismatches"):
            assert v == 0, "%s: %s = %d" % (case, k, v)          # counts bit-exact, accelerations bitwise equal
        elif k.endswith("max rel") or k.endswith("abs/scale"):
            assert v <= 1e-12, "%s: %s = %g" % (case, k, v)
    a, b = rep["interactions"]
    assert a == b
    assert sum(s["own_particles"] for s in stats) == ps["pos"].shape[0]
    if case in ("evrard", "cube", "plummer"):
        # the export of the settled step is a pruned set, not a replication; only a cold-start density (Hsml far from
        # its converged value) may have had to be repeated, never a gravity walk
        for s in stats:
            assert s["full"] == 0 and s["violations"] <= (0 if case == "plummer" else 1)
            assert s["imported_particles"] < 0.9 * (ps["pos"].shape[0] - s["own_particles"])
            assert s["imported_cells"] > 0


def test_cut_cell_violation_is_detected_and_repeated():
    """Cold-start density of the disc: the two sinks start at Hsml = 5e-5 and end near 0.1 - far beyond the padding of
    the export test.  Every rank must notice, all must repeat with a full export, and the results must still be exact."""
    rep, stats, _ = _run("disc", 200000, 4)
    assert all(s["violations"] >= 1 for s in stats)
    assert len({s["violations"] for s in stats}) == 1               # the decision was collective
    assert all(v == 0 for k, v in rep.items() if k.endswith("mismatches"))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "scripts", "let_nccl_check.py")), reason="driver script missing")
def test_two_gpus_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the in-process tests above cover the same path on one)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "scripts", "let_nccl_check.py"), "--particles", "200000"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "IDENTICAL" in r.stdout


def test_c_abi_two_ranks_from_cpp(tmp_path):
    """tests/native/let_ranks.cpp: the same through the C ABI alone - what a reference-side binding does at NTask = 2."""
    exe = str(tmp_path / "let_ranks")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "native", "let_ranks.cpp"), "-ldl", "-pthread"])
    r = subprocess.run([exe, os.path.join(ROOT, "gadget_leicester_b200", "libgl_b200.so"), "40000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout and "mismatching particles 0" in r.stdout


def test_dropin_call_on_every_ranks_own_arrays():
    """gl_gadget_compute_accelerations at NTask = 3: every rank calls it on its OWN P[] / SphP[] (its Peano-Hilbert
    segment, gas first), like the reference's compute_accelerations() at any NTask.  The structs must end up with the
    values a single rank writes into the same particles (counts exact, GravAccel bitwise, SPH fields to 1e-12), twice in a
    row (the second call starts from the Hsml / OldAcc the first one left in P[])."""
    import threading
    import let_check as LC
    from gadget_leicester_b200 import aos, lib as L
    from gadget_leicester_b200.gadget import ForcePath
    ps, par = LC.make_case("disc", 60000)
    n = ps["pos"].shape[0]
    ngas = int((ps["type"] == 0).sum())
    one = ForcePath(par)
    one.load(ps)
    order = one.peano_hilbert_order(ngas)
    pso = {k: (np.ascontiguousarray(v[order]) if hasattr(v, "shape") and v.shape[0] == n else v) for k, v in ps.items()}
    arr1 = aos.GadgetArrays(pso, pinned=False)
    d1 = aos.GadgetDropIn(one, arr1)
    d1.compute_accelerations()
    one.set_params(Ti_Current=1)
    d1.compute_accelerations()
    ref = arr1.results()
    one.close()

    world = 3
    lib = L.load_library()
    group = lib.gl_comm_local_group_create(world)
    # segments: gas split evenly, the two sinks go to the last rank (every rank's P[] has its gas first)
    cuts = [ngas * r // world for r in range(world + 1)]
    segs = [np.arange(cuts[r], cuts[r + 1]) for r in range(world)]
    segs[-1] = np.concatenate([segs[-1], np.arange(ngas, n)])
    out, err = [None] * world, [None] * world

    def worker(r):
        try:
            psr = {k: (np.ascontiguousarray(v[segs[r]]) if hasattr(v, "shape") and v.shape[0] == n else v) for k, v in pso.items()}
            fp = ForcePath(par)
            fp.comm_init_local(r, group)
            arr = aos.GadgetArrays(psr, pinned=False)
            d = aos.GadgetDropIn(fp, arr)
            d.compute_accelerations()
            fp.set_params(Ti_Current=1)
            d.compute_accelerations()
            out[r] = arr.results()
            fp.close()
        except Exception as ex:     # noqa: BLE001
            err[r] = ex
    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    lib.gl_comm_local_group_destroy(group)
    for e in err:
        if e is not None:
            raise e
    for r in range(world):
        sg = segs[r]
        g = sg[sg < ngas]
        assert (out[r]["GravCost"] == ref["GravCost"][sg]).all()
        assert (out[r]["GravAccel"] == ref["GravAccel"][sg]).all()
        assert (out[r]["OldAcc"] == ref["OldAcc"][sg]).all()
        for k in ("Hsml", "NumNgb"):
            e = np.abs(out[r][k] - ref[k][sg]) / np.maximum(np.abs(ref[k][sg]), 1e-300)
            assert e.max() <= 1e-12, k
        for k in ("Density", "Pressure", "HydroAccel", "MaxSignalVel", "DhsmlDensityFactor"):
            a, b = out[r][k][: len(g)], ref[k][g]
            num = np.abs(a - b) if a.ndim == 1 else np.linalg.norm(a - b, axis=-1)
            den = np.abs(b) if a.ndim == 1 else np.linalg.norm(b, axis=-1)
            assert (num / np.maximum(den, 1e-300)).max() <= 1e-11, k


def test_collective_partition_cuts_the_global_peano_hilbert_order_by_cost():
    """SURVEY 8(e) partition (gl_domain_split_collective): four ranks holding an ARBITRARY quarter of a disc each (every
    fourth particle) agree on three splitter keys; every particle is sent to the rank whose stretch of the global
    Peano-Hilbert order holds its key, and the stretches carry equal cost (costs spread over 2^6 here) to within the
    sampling error - against the exact cut computed from all keys at once."""
    import threading
    from gadget_leicester_b200 import ics, lib as L
    from gadget_leicester_b200.gadget import ForcePath
    world = 4
    par = ics.default_params()
    ps = ics.disc(119999, seed=31)                       # + 2 sinks: 120 001 particles, not a multiple of the rank count
    n = ps["pos"].shape[0]
    g = np.random.default_rng(3)
    cost = 2.0 ** (-g.integers(0, 7, n).astype(np.float64))
    lib = L.load_library()
    group = lib.gl_comm_local_group_create(world)
    res, err = [None] * world, [None] * world

    def worker(r):
        try:
            mine = np.arange(r, n, world)
            psr = {k: (np.ascontiguousarray(v[mine]) if hasattr(v, "shape") and v.shape[0] == n else v) for k, v in ps.items()}
            fp = ForcePath(par)
            fp.comm_init_local(r, group, world)
            fp.load(psr)
            dest, togo, spl = fp.domain_split_collective(cost[mine])
            keys = fp.peano_and_morton_keys()[0]
            res[r] = (mine, dest, togo, spl, keys)
            fp.close()
        except Exception as ex:   # noqa: BLE001
            err[r] = ex
    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    lib.gl_comm_local_group_destroy(group)
    for e in err:
        if e is not None:
            raise e
    spl = res[0][3]
    assert spl.size == world - 1 and (np.diff(spl.astype(np.float64)) >= 0).all()
    dest = np.zeros(n, np.int32)
    keys = np.zeros(n, np.uint64)
    for r in range(world):
        mine, d, togo, s, k = res[r]
        assert np.array_equal(s, spl)                                   # the same splitters everywhere
        dest[mine], keys[mine] = d, k
        assert togo[r] == 0 and all(togo[q] == (d == q).sum() for q in range(world) if q != r)
    assert np.array_equal(dest, np.searchsorted(spl, keys, side="right"))       # a key decides the owner, monotonically
    share = np.array([cost[dest == r].sum() for r in range(world)]) / cost.sum()
    assert np.abs(share - 1.0 / world).max() < 2.0 * world / 2048 / world + 1e-3, share
    # against the exact cut of the global order
    order = np.argsort(keys, kind="stable")
    cum = np.cumsum(cost[order])
    exact = np.zeros(n, np.int32)
    exact[order] = np.minimum((cum - cost[order] * 0.5) / cum[-1] * world, world - 1).astype(np.int32)
    assert (exact != dest).mean() < 0.01
