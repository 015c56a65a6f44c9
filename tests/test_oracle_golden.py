"""CPU suite: the C restatement (oracle/gl_oracle.c) must reproduce the golden vectors that the
compiled, unmodified reference produced (tests/golden/make_golden.py).  This is what pins the oracle."""
import os
import subprocess

import numpy as np
import pytest

import golden_util as G
import util

ROOT = util.ROOT


@pytest.fixture(scope="module")
def port_bin():
    src = os.path.join(ROOT, "oracle", "gl_oracle.c")
    if (not os.path.exists(util.PORT_BIN)) or os.path.getmtime(util.PORT_BIN) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", util.PORT_BIN, src, "-lm"])
    return util.PORT_BIN


@pytest.mark.parametrize("name", G.golden_names())
def test_port_reproduces_reference_golden(port_bin, name):
    ps, par, ref = G.load_golden(name)
    n = ps["pos"].shape[0]
    ngas = int((ps["type"] == 0).sum())
    out = util.run_oracle(port_bin, ps, par, G.OPS)
    checked = 0
    for key, want in ref.items():
        stage, field = key.split("__")
        pre = G.STAGES[stage]
        if field in ("nodes", "nodeflags", "numnodes"):
            continue  # the reference adds its (non-interacting) top-tree cells; compared as a set below
        got = out[pre + field]
        if want.size == 0:
            assert got.size == 0
            continue
        if got.shape[0] in (n, ngas) and (pre + "id") in out and field != "domain":
            got = util.by_id(out, pre, field, got.shape[0])
        if want.dtype.kind in "iu":
            assert (got == want).all(), "%s/%s: %d integer mismatches" % (name, key, int((got != want).sum()))
        else:
            # FP64, same operation order, no FMA: identical up to libm pow() rounding
            err = util.rel_err_vec(got, want)
            assert err.max() <= 1e-13, "%s/%s: rel err %.3e" % (name, key, err.max())
        checked += 1
    assert checked >= 10
    # tree: every multi-particle cell of the reference exists in the port with identical moments
    rn, rf = ref["nodes__nodes"], ref["nodes__nodeflags"]
    rn = rn[(rf & 128) != 0]
    pn, pf = out["10_N_nodes"], out["10_N_nodeflags"]
    pn = pn[(pf & 128) != 0]
    key = lambda t: {tuple(r[0:4]): r for r in t}
    a, b = key(rn), key(pn)
    assert set(a) == set(b)
    for k in a:
        assert (a[k][4:10] == b[k][4:10]).all()


def test_interaction_totals_match_golden(port_bin):
    ps, par, ref = G.load_golden("plummer_2k")
    out = util.run_oracle(port_bin, ps, par, G.OPS)
    assert out["5_G_interactions"][0] == ref["grav1__interactions"][0]
    assert out["6_G_interactions"][0] == ref["grav2__interactions"][0]
    assert out["6_G_theta_after"][0] == 0.0  # gravtree.c:396-397
