"""Parity of the CUDA path against the compiled reference AT BENCHMARK SIZE, as a record (numbers, not asserts).

Used by (1) tests/test_gpu_parity_bench.py, which asserts on the record at 200 k particles, (2) bench.py's
cpu_baseline leg, which already runs oracle/_ref/ref_harness on a 200 k sample of config 0 and now keeps its dump
(`parity_at_bench` in the bench line), and (3) scripts/parity_full_n.py for the full 1.6 M disc.

What is compared (BASELINE.json north_star; reference lines in brackets):
  * P.GravCost of the Barnes-Hut pass and of the relative-criterion pass, the latter with the reference's OldAcc as
    input: bit-exact  [forcetree.c:2214, gravtree.c:27]
  * density / hydro pair counts at the reference's converged Hsml: bit-exact  [density.c:856, hydra.c:1266]
  * GravAccel, Density, HydroAccel, DtEntropy: relative error per particle (vector norm), median and 99th percentile,
    tolerance 1e-5  [gravtree.c:403, density.c:875, hydra.c:1616-1621]
TEST INFRASTRUCTURE: the oracle is the checker here, never the thing measured.
"""
import numpy as np

import util

OPS_SETTLE = "T S D R G G D U H C"        # cold start, BH pass, relative pass, settled density, hmax, hydro, pair counts
STAGES = dict(hsml0="1_S_", dens1="2_D_", grav1="4_G_", grav2="5_G_", dens2="6_D_", hydro="8_H_", pairs="9_C_")


def golden_layout(out, n, ngas):
    """ref_harness dump (Peano-Hilbert order + ids) -> {stage__field: array in input order}"""
    R = {}
    for stage, pre in STAGES.items():
        for k, v in out.items():
            if not k.startswith(pre) or k.endswith("_id") or k.endswith("_time"):
                continue
            field = k[len(pre):]
            if (pre + "id") in out and v.shape[0] in (n, ngas) and v.size:
                v = util.by_id(out, pre, field, v.shape[0])
            R[stage + "__" + field] = v
    return R


def _pct(e):
    e = np.asarray(e)
    return dict(p50=float(np.median(e)), p99=float(np.percentile(e, 99)), max=float(e.max()))


def parity_record(ForcePath, ps, par, R, device=0):
    """Run the CUDA path on (ps, par) and compare with the reference results R (golden layout)."""
    n = ps["pos"].shape[0]
    gas = ps["type"] == 0
    ngas = int(gas.sum())
    sph = gas | (ps["type"] == 5) | (ps["type"] == 2)
    ps = dict(ps)
    ps["hsml"] = R["hsml0__Hsml"].copy()
    rec = dict(particles=int(n), gas=ngas, checker="oracle/_ref/ref_harness (the unmodified reference, gcc -O3 -ffp-contract=off)")
    fp = ForcePath(par, device=device)
    fp.load(ps)
    fp.force_treebuild()
    d1 = fp.density()                                   # cold start: the whole Hsml iteration
    e = util.rel_err_vec(d1["Hsml"], R["dens1__Hsml"])
    rec["hsml_after_iteration_max_rel"] = float(e.max())
    rec["hsml_after_iteration_differs_at"] = int((d1["Hsml"] != R["dens1__Hsml"]).sum())
    fp.set_Hsml(R["dens1__Hsml"])                       # from here on: fixed h = the reference's (bit-exact inputs)
    fp.set_TreeReconstructFlag(1)
    g1 = fp.gravity_tree()                              # Barnes-Hut criterion
    rec["gravcost_bh_mismatches"] = int((g1["GravCost"] != R["grav1__GravCost"]).sum())
    rec["interactions_bh"] = int(fp.stats()["interactions"])
    rec["gravaccel_bh"] = _pct(util.rel_err_vec(g1["GravAccel"], R["grav1__GravAccel"]))
    fp.set_OldAcc(R["grav1__OldAcc"])
    g2 = fp.gravity_tree()                              # relative criterion
    rec["gravcost_rel_mismatches"] = int((g2["GravCost"] != R["grav2__GravCost"]).sum())
    rec["interactions_rel"] = int(fp.stats()["interactions"])
    rec["interactions_rel_reference"] = int(R["grav2__GravCost"].astype(np.int64).sum())
    rec["gravaccel_rel"] = _pct(util.rel_err_vec(g2["GravAccel"], R["grav2__GravAccel"]))
    if ngas:
        d2 = fp.density()                               # settled pass at fixed h
        rec["density_passes_settled"] = int(fp.stats()["density_passes_max"])
        rec["density_pair_mismatches"] = int((d2["DensPairs"][sph] != R["pairs__DensPairs"][sph]).sum())
        rec["density"] = _pct(util.rel_err_vec(d2["Density"][:ngas], R["dens2__Density"]))
        rec["pressure"] = _pct(util.rel_err_vec(d2["Pressure"][:ngas], R["dens2__Pressure"]))
        fp.force_update_hmax()
        h = fp.hydro_force()
        rec["hydro_pair_mismatches"] = int((h["HydroPairs"][:ngas] != R["pairs__HydroPairs"][:ngas]).sum())
        rec["hydroaccel"] = _pct(util.rel_err_vec(h["HydroAccel"][:ngas], R["hydro__HydroAccel"]))
        ref = R["hydro__DtEntropy"]
        rec["dtentropy"] = _pct(np.abs(h["DtEntropy"][:ngas] - ref) / (np.abs(ref) + 1e-12 * np.abs(ref).max() + 1e-300))
        rec["maxsignalvel"] = _pct(util.rel_err_vec(h["MaxSignalVel"][:ngas], R["hydro__MaxSignalVel"]))
    fp.close()
    rec["counts_bit_exact"] = bool(rec["gravcost_bh_mismatches"] == 0 and rec["gravcost_rel_mismatches"] == 0 and
                                   rec.get("density_pair_mismatches", 0) == 0 and rec.get("hydro_pair_mismatches", 0) == 0)
    worst = max(rec["gravaccel_bh"]["p99"], rec["gravaccel_rel"]["p99"], rec.get("density", dict(p99=0))["p99"],
                rec.get("hydroaccel", dict(p99=0))["p99"], rec.get("dtentropy", dict(p99=0))["p99"])
    rec["fields_p99_worst"] = float(worst)
    rec["tolerance"] = 1e-5
    rec["ok"] = bool(rec["counts_bit_exact"] and worst <= 1e-5)
    return rec
