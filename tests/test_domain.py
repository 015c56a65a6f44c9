"""CPU suite: gl_domain_top_tree (host code of libgl_b200.so) against the top-level tree the reference's own
domain_Decomposition() builds (domain.c:1738 domain_determineTopTree; dumped by oracle/_ref/ref_harness as top_*):
node numbering, StartKey, Size, Daughter and the leaf numbers, for particle sets with equal and with very unequal cost
factors.  Runs against the compiled reference where it exists and against a committed dump of it anywhere."""
import ctypes as C
import os

import numpy as np
import pytest

import util
import golden_util as G
from gadget_leicester_b200 import ics
from gadget_leicester_b200 import lib as L

FIX = os.path.join(util.ROOT, "tests", "golden", "aux", "toptree_cases.npz")
TIMEBASE = 1 << 29


def costfactor(timebin, gravcost=1.0):
    """domain_particle_costfactor, domain.c:378-384"""
    tb = np.asarray(timebin, dtype=np.int64)
    return np.where(tb > 0, (1.0 + gravcost) / (1 << np.maximum(tb, 0)), (1.0 + gravcost) / TIMEBASE)


def top_tree(keys, cost, max_nodes, ntask=1, multipledomains=16, topnodefactor=2.5):
    lib = L.load_library()
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    nodes = (L.GLTopNode * max_nodes)()
    nn, nl = C.c_int(0), C.c_int(0)
    rc = lib.gl_domain_top_tree(L._ptr(keys), L._ptr(cost), keys.size, ntask, multipledomains, topnodefactor, nodes, max_nodes,
                                C.byref(nn), C.byref(nl))
    a = np.frombuffer(nodes, dtype=np.dtype([("Size", "u8"), ("StartKey", "u8"), ("Count", "i8"), ("Cost", "f8"), ("Daughter", "i4"),
                                             ("Leaf", "i4"), ("Parent", "i4"), ("PIndex", "i4")]))[: nn.value].copy()
    return rc, a, nl.value


def cases():
    out = {}
    for name in ("cube_timebins_2k", "disc_3k_sinks", "mixed_336", "plummer_2k"):
        ps, par, _ = G.load_golden(name)
        out[name] = (ps, par)
    ps = ics.disc(40000, seed=7)
    g = np.random.default_rng(8)
    ps["timebin"] = g.integers(0, 9, ps["pos"].shape[0]).astype(np.int32)       # costs spread over 2^28
    par = ics.default_params()
    par["PartAllocFactor"] = 10.0
    out["disc_40k_timebins"] = (ps, par)
    return out


def reference_dump(ps, par):
    """keys in INPUT order (what domain_determineTopTree sees) and the reference's TopNodes[]"""
    o = util.run_oracle(util.REF_HARNESS, ps, par, "K")
    by_id = dict(zip(o["0_K_id"].tolist(), o["0_K_phkey"].astype(np.uint64).tolist()))
    keys = np.array([by_id[int(i)] for i in ps["id"]], dtype=np.uint64)
    return dict(keys=keys, cost=costfactor(ps["timebin"]), n=o["top_n"], StartKey=o["top_StartKey"].astype(np.uint64),
                Size=o["top_Size"].astype(np.uint64), Daughter=o["top_Daughter"], Leaf=o["top_Leaf"], max_nodes=int(o["top_domain"][4]))


def check(d):
    rc, a, nleaves = top_tree(d["keys"], d["cost"], int(d["max_nodes"]))
    assert rc == 0
    assert (a.size, nleaves) == (int(d["n"][0]), int(d["n"][1]))
    assert np.array_equal(a["StartKey"], d["StartKey"]) and np.array_equal(a["Size"], d["Size"])
    assert np.array_equal(a["Daughter"], d["Daughter"])
    leaf = a["Daughter"] < 0
    assert np.array_equal(a["Leaf"][leaf], d["Leaf"][leaf])
    # the tree is a partition of the particles
    assert a["Count"][0] == d["keys"].size and a["Count"][leaf].sum() == d["keys"].size
    return a


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref/ref_harness not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["cube_timebins_2k", "disc_3k_sinks", "mixed_336", "plummer_2k", "disc_40k_timebins"])
def test_top_tree_equals_reference(name):
    ps, par = cases()[name]
    a = check(reference_dump(ps, par))
    if name == "disc_40k_timebins":
        n = ps["pos"].shape[0]
        refined = a["Daughter"] >= 0
        assert a.size > 100 and (refined & (a["Count"] <= n / 40.0)).any()      # some cells were split for their COST alone


@pytest.mark.parametrize("name", ["cube_timebins_2k", "disc_40k_timebins"])
def test_top_tree_equals_committed_reference_dump(name):
    z = np.load(FIX)
    check({k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + ".")})


REF_HOST_DROPIN = os.path.join(util.ROOT, "oracle", "_ref", "ref_host_dropin")


@pytest.mark.skipif(not (util.have_ref() and os.path.exists(REF_HOST_DROPIN)), reason="oracle/_ref/ref_host_dropin not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["cube_timebins_2k", "disc_3k_sinks", "mixed_336", "plummer_2k", "disc_40k_timebins"])
def test_reference_host_with_the_glue_domain_decomposition(name):
    """The reference harness linked with the one-rank domain_Decomposition() of integration/gl_gadget_glue.c (-DGL_DOMAIN;
    its pieces run in the library) next to the all-reference harness: the cube, TopNodes[], the particle order and -
    through the CPU tree built on that top-level tree - every output of a tree build, a density pass, a gravity pass and
    the node table come out bit-identical."""
    ps, par = cases()[name]
    ops = "K T S D R G N" if (ps["type"] == 0).any() else "K T R G N"        # (no gas: no density pass)
    want = util.run_oracle(util.REF_HARNESS, ps, par, ops)
    got = util.run_oracle(REF_HOST_DROPIN, ps, par, ops)
    assert set(got) == set(want) and {"top_StartKey", "top_Leaf", "top_StartList", "top_EndList", "top_DomainTask"} <= set(want)
    compared = 0
    for k in want:
        if "time" in k or "CPU_Step" in k:
            continue
        assert got[k].shape == want[k].shape and np.array_equal(got[k], want[k]), k
        compared += 1
    assert compared > 15


def test_top_tree_reports_lack_of_room():
    z = np.load(FIX)
    pre = "disc_40k_timebins."
    rc, _, _ = top_tree(z[pre + "keys"], z[pre + "cost"], 40)
    assert rc == 1                   # the reference raises All.TopNodeAllocFactor and repeats (domain.c:246-269)


if __name__ == "__main__":          # python tests/test_domain.py  ->  rewrites the committed dump (needs the reference)
    out = {}
    for name in ("cube_timebins_2k", "disc_40k_timebins"):
        d = reference_dump(*cases()[name])
        for k, v in d.items():
            out[name + "." + k] = np.asarray(v)
    np.savez_compressed(FIX, **out)
    print("wrote", FIX, os.path.getsize(FIX) // 1024, "KiB")
