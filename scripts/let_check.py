#!/usr/bin/env python
"""Multi-rank path on ONE GPU: `world` contexts of this process (one host thread each, in-process group) against a
single-context run of the whole set.  Prints what was exported/imported and how the results compare.

  python scripts/let_check.py [--n 200000] [--world 4] [--case disc|evrard|plummer|cube]
"""
import argparse
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gadget_leicester_b200 import ics, lib as L          # noqa: E402
from gadget_leicester_b200.gadget import ForcePath       # noqa: E402


def make_case(case, n):
    par = ics.default_params()
    if case == "disc":
        ps = ics.disc(n, seed=42)
    elif case == "evrard":
        ps = ics.evrard(n, seed=42)
        par = ics.default_params(Softening=[1e-4, 5e-5, 5e-5, 1e-7, 0.01, 5e-5])
    elif case == "plummer":
        ps = ics.plummer(n, seed=42)
        par = ics.default_params(Softening=[1e-5, ics.plummer_softening(n), 5e-5, 1e-7, 0.01, 5e-5])
    else:
        ps = ics.cube(n, seed=42)
    if (ps["type"] == 0).any():
        ps["hsml"] = ics.guess_hsml(ps, par["DesNumNgb"])
        ps["hsml"][ps["type"] == 5] = par["Softening"][5]
    return ps, par


def sequence(fp, sph, reload=None):
    out = {}
    out["g1"] = fp.gravity_tree()            # Barnes-Hut pass (OldAcc = 0)
    out["i1"] = fp.stats()["interactions"]
    out["g2"] = fp.gravity_tree()            # relative criterion
    out["i2"] = fp.stats()["interactions"]
    out["ms_walk"] = fp.stats()["ms_gravity_walk"]
    if sph:
        out["d"] = fp.density()
        out["passes"] = fp.stats()["density_passes_max"]
        out["slowest"] = (fp.stats()["density_slowest_index"], fp.stats()["density_slowest_passes"])
        fp.force_update_hmax()
        out["h"] = fp.hydro_force()
        # a settled second step: the caller loads its particles again (new Hsml / OldAcc), as every step of a run does
        if reload is not None:
            reload(out)
        fp.set_TreeReconstructFlag(1)
        out["g3"] = fp.gravity_tree()
        out["d2"] = fp.density()
        out["slowest2"] = (fp.stats()["density_slowest_index"], fp.stats()["density_slowest_passes"])
        fp.force_update_hmax()
        out["h2"] = fp.hydro_force()
    return out


def run_world(ps, par, order, world, sph, padding=None):
    n = ps["pos"].shape[0]
    lib = L.load_library()
    group = lib.gl_comm_local_group_create(world)
    m = -(-n // world)
    res, err, stats = [None] * world, [None] * world, [None] * world

    def worker(r):
        try:
            seg = order[r * m: min(n, (r + 1) * m)]
            psr = {k: (np.ascontiguousarray(v[seg]) if hasattr(v, "shape") and v.shape[0] == n else v) for k, v in ps.items()}
            fp = ForcePath(par)
            fp.comm_init_local(r, group)
            if padding:
                fp.comm_set_options(True, padding)
            fp.load(psr)

            def reload(out):
                p2 = dict(psr, hsml=out["d"]["Hsml"].copy(), oldacc=out["g2"]["OldAcc"].copy())
                fp.load(p2)
            res[r] = sequence(fp, sph, reload)
            stats[r] = fp.let_stats()
            res[r]["seg"] = seg
            fp.close()
        except Exception as ex:   # noqa: BLE001
            err[r] = ex
    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    t0 = time.time()
    for t in th:
        t.start()
    for t in th:
        t.join()
    lib.gl_comm_local_group_destroy(group)
    for e in err:
        if e is not None:
            raise e
    return res, stats, time.time() - t0


def compare(ref, res, sph, ps_type):
    """ref: single-context outputs in input order; res: per-rank outputs (own order) + 'seg' (input indices)."""
    rep = {}

    def put(name, a, b, exact=False):
        a, b = np.asarray(a), np.asarray(b)
        if exact:
            rep[name] = int((a != b).sum())
        else:
            den = np.abs(b) if a.ndim == 1 else np.linalg.norm(b, axis=-1)
            num = np.abs(a - b) if a.ndim == 1 else np.linalg.norm(a - b, axis=-1)
            den = np.where(den > 0, den, 1.0)
            rep[name] = float((num / den).max()) if num.size else 0.0
    # concatenate ranks first
    seg = np.concatenate([r["seg"] for r in res])
    cat = lambda tag, k: np.concatenate([r[tag][k] for r in res])          # noqa: E731
    for tag in ("g1", "g2") + (("g3",) if sph else ()):
        put(tag + ".GravCost mismatches", cat(tag, "GravCost"), ref[tag]["GravCost"][seg], True)
        put(tag + ".GravAccel bitwise mismatches", cat(tag, "GravAccel"), ref[tag]["GravAccel"][seg], True)
        put(tag + ".GravAccel max rel", cat(tag, "GravAccel"), ref[tag]["GravAccel"][seg])
    rep["interactions"] = (sum(r["i2"] for r in res), ref["i2"])
    if sph:
        for tag, htag in (("d", "h"), ("d2", "h2")):
            put(tag + ".DensPairs mismatches", cat(tag, "DensPairs"), ref[tag]["DensPairs"][seg], True)
            put(tag + ".Hsml max rel", cat(tag, "Hsml"), ref[tag]["Hsml"][seg])
            gas = ps_type[seg] == 0         # SphP fields exist for gas only (hydra.c:258, density.c:440): other entries are never written
            put(tag + ".Density max rel", cat(tag, "Density")[gas], ref[tag]["Density"][seg][gas])
            put(htag + ".HydroPairs mismatches", cat(htag, "HydroPairs")[gas], ref[htag]["HydroPairs"][seg][gas], True)
            put(htag + ".HydroAccel max rel", cat(htag, "HydroAccel"), ref[htag]["HydroAccel"][seg])
            put(htag + ".DtEntropy max abs/scale", cat(htag, "DtEntropy") / (np.abs(ref[htag]["DtEntropy"]).max() + 1e-300),
                ref[htag]["DtEntropy"][seg] / (np.abs(ref[htag]["DtEntropy"]).max() + 1e-300))
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200000)
    ap.add_argument("--world", type=int, default=4)
    ap.add_argument("--case", default="disc")
    ap.add_argument("--padding", type=float, default=0.0)
    ap.add_argument("--debug", action="store_true")
    a = ap.parse_args()
    ps, par = make_case(a.case, a.n)
    n = ps["pos"].shape[0]
    ngas = int((ps["type"] == 0).sum())
    sph = ngas > 0
    one = ForcePath(par)
    one.load(ps)
    order = one.peano_hilbert_order(ngas)
    one.load(ps)

    def reload_one(out):
        one.load(dict(ps, hsml=out["d"]["Hsml"].copy(), oldacc=out["g2"]["OldAcc"].copy()))
    ref = sequence(one, sph, reload_one)
    one.close()
    res, stats, wall = run_world(ps, par, order, a.world, sph, a.padding or None)
    rep = compare(ref, res, sph, ps["type"])
    print("case %s, %d particles, %d ranks (in-process group), wall %.2f s" % (a.case, n, a.world, wall))
    for r, st in enumerate(stats):
        print("  rank %d: own %d | exported %d cells %d particles | imported %d cells %d particles (%.1f %% of the rest) | "
              "violations %d full %d | own tree %.2f select %.2f exchange %.2f sph %.2f ms" % (
                  r, st["own_particles"], st["exported_cells"], st["exported_particles"], st["imported_cells"], st["imported_particles"],
                  100.0 * st["imported_particles"] / max(1, n - st["own_particles"]), st["violations"], st["full"],
                  st["ms_own_tree"], st["ms_select"], st["ms_exchange"], st["ms_sph_exchange"]))
    if sph and a.debug:
        seg = np.concatenate([r["seg"] for r in res])
        for dt, ht in (("d", "h"), ("d2", "h2")):
            mine = np.concatenate([r[ht]["HydroPairs"] for r in res])
            hs = np.concatenate([r[dt]["Hsml"] for r in res])
            badi = np.nonzero((mine != ref[ht]["HydroPairs"][seg]) & (ps["type"][seg] == 0))[0]
            for b in badi[:10]:
                g = seg[b]
                print("   %s mismatch: particle %d type %d pairs %d vs %d  Hsml %.17g vs %.17g" % (
                    ht, g, ps["type"][g], mine[b], ref[ht]["HydroPairs"][g], hs[b], ref[dt]["Hsml"][g]))
            if len(badi):
                # who is right?  brute force over all gas particles with the single-context Hsml (hydra.c:1266)
                hall = ref[dt]["Hsml"]
                gasm = (ps["type"] == 0) & (ps["mass"] != 0)
                for b in badi[:5]:
                    g = seg[b]
                    d2 = ((ps["pos"][gasm] - ps["pos"][g]) ** 2).sum(axis=1)
                    cnt = int((((d2 < hall[g] ** 2) | (d2 < hall[gasm] ** 2)) & (d2 > 0)).sum())
                    print("      brute force for particle %d: %d pairs" % (g, cnt))
            dh = np.abs(hs - ref[dt]["Hsml"][seg]) / ref[dt]["Hsml"][seg]
            print("   %s: particles whose Hsml differs: %d (max rel %.3g)" % (dt, int((dh > 0).sum()), dh.max()))
    if sph:
        for tag in ("slowest", "slowest2"):
            i, k = ref[tag]
            print("  single context, %s: P[%d] (type %s) took %d passes" % (tag, i, ps["type"][i] if i >= 0 else "-", k))
    bad = 0
    for k, v in rep.items():
        print("  %-36s %s" % (k, v))
        if k.endswith("mismatches") and "bitwise" not in k and v != 0:
            bad += 1
    print("RESULT:", "OK" if bad == 0 else "MISMATCH")
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
