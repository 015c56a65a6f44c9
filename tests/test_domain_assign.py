"""CPU suite: the multi-rank half of domain_decompose() (domain.c:470-487) - gl_domain_assign and
gl_domain_count_to_go (host code of libgl_b200.so) against the reference's OWN functions at NTask = 2 ... 8.

domain_findSplit_work_balanced (:1075), domain_assign_load_or_work_balanced (:1153), domain_check_memory_bound (:595) and
domain_findSplit_load_balanced (:1225) run identically on every rank on tables summed over the ranks, so no second rank is
needed to run them: oracle/_ref/ref_domain_assign includes domain.c where it lies (the tables are file-static) and calls
them with NTask set.  Runs against that binary where it exists and against a committed dump of its answers anywhere."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import util
from gadget_leicester_b200 import ics
from gadget_leicester_b200 import lib as L
from test_domain import top_tree, costfactor

REF = os.path.join(util.ROOT, "oracle", "_ref", "ref_domain_assign")
FIX = os.path.join(util.ROOT, "tests", "golden", "aux", "domain_assign_cases.npz")
MD = 16         # MULTIPLEDOMAINS of the shipped Makefile


def ours(ntask, work, count, countsph, max_load, max_load_sph):
    lib = L.load_library()
    lib.gl_domain_assign.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_longlong, C.c_longlong] + [C.c_void_p] * 6
    n = work.size
    task = np.full(n, -1, np.int32)
    sl, el = np.zeros(MD * ntask, np.int32), np.zeros(MD * ntask, np.int32)
    ll, ls, lw = np.zeros(ntask, np.int32), np.zeros(ntask, np.int32), np.zeros(ntask)
    rc = lib.gl_domain_assign(ntask, MD, n, *[a.ctypes.data for a in (work, count, countsph)], max_load, max_load_sph,
                              *[a.ctypes.data for a in (task, sl, el, ll, ls, lw)])
    return dict(status=rc, task=task, start=sl, end=el, load=ll, loadsph=ls, work=lw)


def reference(ntask, work, count, countsph, max_load, max_load_sph, tmp):
    fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    with open(fin, "wb") as f:
        f.write(np.array([ntask, work.size, max_load, max_load_sph], np.int32).tobytes())
        f.write(work.tobytes() + count.tobytes() + countsph.tobytes())
    subprocess.run([REF, fin, fout], check=True, stdout=subprocess.DEVNULL)
    b = open(fout, "rb").read()
    n, p = work.size, 4
    out = dict(status=int(np.frombuffer(b[:4], np.int32)[0]))
    for k, cnt, dt in (("task", n, np.int32), ("start", MD * ntask, np.int32), ("end", MD * ntask, np.int32), ("load", ntask, np.int32),
                       ("loadsph", ntask, np.int32), ("work", ntask, np.float64)):
        sz = cnt * np.dtype(dt).itemsize
        out[k] = np.frombuffer(b[p:p + sz], dt).copy()
        p += sz
    return out


def leaf_tables(ntask, n=60000, seed=5):
    """leaf work / count / gas count of a disc's top-level tree as NTask ranks would build it (limits scale with NTask)"""
    ps = ics.disc(n, seed=seed)
    g = np.random.default_rng(seed)
    ps["timebin"] = g.integers(0, 7, ps["pos"].shape[0]).astype(np.int32)
    lib = L.load_library()
    from gadget_leicester_b200.gadget import ForcePath                     # noqa: F401  (keys come from the host function below)
    pos = ps["pos"]
    lo, hi = pos.min(axis=0), pos.max(axis=0)
    length = (hi - lo).max() * 1.001
    fac = (1 << 21) / length
    ijk = ((pos - lo) * fac).astype(np.int64)
    keys = np.array([lib.gl_peano_hilbert_key(int(a), int(b), int(c), 21) for a, b, c in ijk], dtype=np.uint64)
    cost = costfactor(ps["timebin"])
    rc, nodes, nleaves = top_tree(keys, cost, 20000, ntask=ntask)
    assert rc == 0 and nleaves >= MD * ntask
    lib.gl_domain_split.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int] + [C.c_void_p] * 4
    raw = (L.GLTopNode * nodes.size).from_buffer_copy(nodes.tobytes())
    sl, el = np.zeros(MD * ntask, np.int32), np.zeros(MD * ntask, np.int32)
    work, count = np.zeros(nleaves, np.float32), np.zeros(nleaves, np.int32)
    assert lib.gl_domain_split(raw, nleaves, keys.ctypes.data, cost.ctypes.data, keys.size, MD * ntask, sl.ctypes.data, el.ctypes.data,
                               work.ctypes.data, count.ctypes.data) == 0
    # gas count per leaf
    countsph = np.zeros(nleaves, np.int32)
    gas = ps["type"] == 0
    leaf_of = np.zeros(keys.size, np.int64)
    for i in np.nonzero(gas)[0][::1]:
        no = 0
        while nodes["Daughter"][no] >= 0:
            no = nodes["Daughter"][no] + int((int(keys[i]) - int(nodes["StartKey"][no])) // (int(nodes["Size"][no]) >> 3))
        leaf_of[i] = nodes["Leaf"][no]
    np.add.at(countsph, leaf_of[gas], 1)
    return dict(keys=keys, types=ps["type"].astype(np.int32), nodes=nodes, work=work, count=count, countsph=countsph)


def cases():
    out = {}
    for ntask in (2, 3, 4, 8):
        t = leaf_tables(ntask, n=20000 if ntask < 8 else 40000, seed=ntask)
        n = int(t["count"].sum())
        out["disc_ntask%d" % ntask] = (ntask, t["work"], t["count"], t["countsph"], 2 * n, 2 * n)
        # a ceiling the work-balanced deal exceeds: second variant; and one nothing fits under
        out["disc_ntask%d_tight" % ntask] = (ntask, t["work"], t["count"], t["countsph"], int(1.02 * n / ntask) + 1, 2 * n)
        out["disc_ntask%d_impossible" % ntask] = (ntask, t["work"], t["count"], t["countsph"], n // ntask - 1, 2 * n)
    g = np.random.default_rng(1)
    for ntask in (2, 5):
        m = MD * ntask + 37
        work = (g.random(m) ** 4 * 100).astype(np.float32)
        count = g.integers(1, 500, m).astype(np.int32)
        out["random_ntask%d" % ntask] = (ntask, work, count, (count // 2).astype(np.int32), int(count.sum()), int(count.sum()))
        # a ceiling between what the two variants need: the work-balanced deal is rejected, the load-balanced one fits (NTask 5)
        out["random_ntask%d_ceiling" % ntask] = (ntask, work, count, (count // 2).astype(np.int32), int(1.16 * count.sum() / ntask), int(count.sum()))
        out["ties_ntask%d" % ntask] = (ntask, np.ones(m, np.float32), np.full(m, 10, np.int32), np.full(m, 5, np.int32), 10 * m, 10 * m)
    return out


def same(a, b):
    assert a["status"] == b["status"]
    for k in ("task", "start", "end", "load", "loadsph"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["work"], b["work"])


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ref_domain_assign not built (needs /root/reference)")
def test_assignment_equals_the_reference_functions(tmp_path):
    seen = set()
    for name, (ntask, work, count, countsph, ml, mls) in cases().items():
        r = reference(ntask, work, count, countsph, ml, mls, str(tmp_path))
        same(ours(ntask, work, count, countsph, ml, mls), r)
        seen.add(r["status"])
        if r["status"] < 2:
            assert set(r["task"].tolist()) == set(range(ntask)) and r["load"].sum() == count.sum()
    assert seen == {0, 1, 2}             # work-balanced, load-balanced after the memory ceiling, nothing fits


def test_assignment_equals_committed_reference_answers():
    z = np.load(FIX)
    names = sorted({k.split(".")[0] for k in z.files})
    assert len(names) >= 10
    for name in names:
        q = {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith(name + ".")}
        got = ours(int(q["ntask"]), q["in_work"], q["in_count"], q["in_countsph"], int(q["max_load"]), int(q["max_load_sph"]))
        same(got, dict(status=int(q["status"]), task=q["task"], start=q["start"], end=q["end"], load=q["load"], loadsph=q["loadsph"], work=q["work"]))


def test_count_to_go_follows_the_assignment():
    ntask = 4
    t = leaf_tables(ntask, n=20000, seed=4)
    n = int(t["count"].sum())
    a = ours(ntask, t["work"], t["count"], t["countsph"], 2 * n, 2 * n)
    lib = L.load_library()
    lib.gl_domain_count_to_go.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 3
    raw = (L.GLTopNode * t["nodes"].size).from_buffer_copy(t["nodes"].tobytes())
    # the particles of "rank 1" = every fourth particle (an arbitrary distribution before the exchange)
    mine = np.arange(1, t["keys"].size, ntask)
    keys, types = np.ascontiguousarray(t["keys"][mine]), np.ascontiguousarray(t["types"][mine])
    togo, togosph, dest = np.zeros(ntask, np.int32), np.zeros(ntask, np.int32), np.zeros(mine.size, np.int32)
    assert lib.gl_domain_count_to_go(raw, keys.ctypes.data, types.ctypes.data, mine.size, a["task"].ctypes.data, 1, ntask,
                                     togo.ctypes.data, togosph.ctypes.data, dest.ctypes.data) == 0
    assert togo[1] == 0 and togo.sum() == (dest != 1).sum()
    for ta in range(ntask):
        if ta != 1:
            assert togo[ta] == (dest == ta).sum() and togosph[ta] == ((dest == ta) & (types == 0)).sum()
    # summed over the "ranks", every rank receives exactly the load the assignment promised
    total = np.zeros(ntask, np.int64)
    for r in range(ntask):
        m = np.arange(r, t["keys"].size, ntask)
        k = np.ascontiguousarray(t["keys"][m])
        d = np.zeros(m.size, np.int32)
        tg = np.zeros(ntask, np.int32)
        assert lib.gl_domain_count_to_go(raw, k.ctypes.data, None, m.size, a["task"].ctypes.data, r, ntask, tg.ctypes.data, None, d.ctypes.data) == 0
        total += np.bincount(d, minlength=ntask)
    assert np.array_equal(total, a["load"])


if __name__ == "__main__":          # python tests/test_domain_assign.py  ->  rewrites the committed dump (needs the reference)
    import tempfile
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for name, (ntask, work, count, countsph, ml, mls) in cases().items():
            r = reference(ntask, work, count, countsph, ml, mls, td)
            out[name + ".ntask"], out[name + ".max_load"], out[name + ".max_load_sph"] = np.int32(ntask), np.int64(ml), np.int64(mls)
            out[name + ".in_work"], out[name + ".in_count"], out[name + ".in_countsph"] = work, count, countsph
            for k, v in r.items():
                out[name + "." + k] = np.asarray(v)
    np.savez_compressed(FIX, **out)
    print("wrote", FIX, os.path.getsize(FIX) // 1024, "KiB")
