"""GPU suite, slow: parity against the COMPILED REFERENCE at the size bench.py's cpu_baseline leg uses - BASELINE config 0's
generator at 200 000 gas + 2 sinks (the full 1.6 M record is scripts/parity_full_n.py -> profiles/r02_parity_full_n.json).

Counts bit-exact (P.GravCost for both opening criteria, density and hydro pair counts at the reference's Hsml); GravAccel,
Density, HydroAccel, DtEntropy within the north_star tolerance 1e-5 (median and 99th percentile are in the record)."""
import pytest

import parity_bench
import util
from gadget_leicester_b200 import ics

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref/ref_harness did not travel with this snapshot")
def test_config0_sample_against_the_compiled_reference():
    from gadget_leicester_b200.gadget import ForcePath
    par = ics.default_params()
    ps = ics.disc(200000, seed=42)
    n, ngas = ps["pos"].shape[0], int((ps["type"] == 0).sum())
    out = util.run_oracle(util.REF_HARNESS, ps, par, parity_bench.OPS_SETTLE)
    rec = parity_bench.parity_record(ForcePath, ps, par, parity_bench.golden_layout(out, n, ngas))
    assert rec["gravcost_bh_mismatches"] == 0 and rec["gravcost_rel_mismatches"] == 0, rec
    assert rec["interactions_rel"] == rec["interactions_rel_reference"]
    assert rec["density_pair_mismatches"] == 0 and rec["hydro_pair_mismatches"] == 0, rec
    for k in ("gravaccel_bh", "gravaccel_rel"):
        assert rec[k]["p99"] <= 1e-5 and rec[k]["p50"] <= 1e-6, (k, rec[k])
    for k in ("density", "pressure", "hydroaccel", "dtentropy"):
        assert rec[k]["p99"] <= 1e-5, (k, rec[k])          # FP64 sums: 1e-15 in practice
    assert rec["hsml_after_iteration_max_rel"] <= 1e-9
    assert rec["ok"]
