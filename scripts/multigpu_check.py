"""torchrun --nproc-per-node N scripts/multigpu_check.py [ngas]: N-GPU sharded force evaluation vs a single-GPU one.

Every rank evaluates its Peano-Hilbert segment through DistributedForcePath (NCCL all-gathers); rank 0 also
evaluates ALL particles alone and the owned slices must agree: gravity bit for bit (same tree, same per-target
arithmetic), SPH to FP64 round-off (warp grouping changes the summation order)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gadget_leicester_b200 import ics  # noqa: E402
from gadget_leicester_b200.multigpu import DistributedForcePath  # noqa: E402
from gadget_leicester_b200.gadget import ForcePath  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ngas = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
par = ics.default_params()
ps = ics.disc(ngas, seed=9)
ps["hsml"] = ics.guess_hsml(ps, par["DesNumNgb"])
ps["hsml"][ps["type"] == 5] = par["Softening"][5]
n = ps["pos"].shape[0]
d = DistributedForcePath(par, rank=rank, world=world, device=local)
order = d.decompose(ps)
d.compute_accelerations()      # BH pass + cold-start density + hydro
res0 = {k: v[d.lo:d.hi].clone().cpu().numpy() for k, v in d.res.items()}
d.fp.set_TreeReconstructFlag(1)
d.compute_accelerations()      # relative criterion, steady density
torch.cuda.synchronize()
res = {k: v[d.lo:d.hi].cpu().numpy() for k, v in d.res.items()}
ok = True
if True:
    # single-GPU reference of the same two evaluations, on this rank's GPU, all targets
    ps_ph = {k: (v[order] if hasattr(v, "shape") and v.shape[0] == n else v) for k, v in ps.items()}
    fp = ForcePath(par, device=local)
    fp.load(ps_ph)
    for _ in range(2):           # the same two evaluations: (BH, cold density, hydro) then (relative, steady, hydro)
        fp.set_TreeReconstructFlag(1)
        fp.gravity_tree_nocopy()
        fp.density_nocopy()
        fp.force_update_hmax()
        fp.hydro_force_nocopy()
    acc, old, cost = np.zeros((n, 3)), np.zeros(n), np.zeros(n, np.float32)
    import gadget_leicester_b200.lib as L
    fp._chk(fp._lib.gl_get_gravity(fp.ctx, L._ptr(acc), L._ptr(old), L._ptr(cost), 0))
    hs, de = np.zeros(n), np.zeros(n)
    fp._chk(fp._lib.gl_get_density(fp.ctx, L._ptr(hs), None, L._ptr(de), None, None, None, None, None, None, None, 0))
    ha = np.zeros((n, 3))
    fp._chk(fp._lib.gl_get_hydro(fp.ctx, L._ptr(ha), None, None, 0))
    sl = slice(d.lo, d.hi)
    gas = ps_ph["type"][sl] == 0
    checks = {
        "GravCost identical": bool((res["gravcost"] == cost[sl]).all()),
        "GravAccel identical": bool((res["acc"] == acc[sl]).all()),
        "Hsml rel<=1e-12": bool(np.abs(res["hsml"] / hs[sl] - 1)[ps_ph["type"][sl] != 1].max() <= 1e-12),
        "Density rel<=1e-11": bool(np.abs(res["density"][gas] / de[sl][gas] - 1).max() <= 1e-11),
        "HydroAccel rel<=1e-8": bool((np.linalg.norm(res["hacc"][gas] - ha[sl][gas], axis=1) /
                                      np.linalg.norm(ha[sl][gas], axis=1)).max() <= 1e-8),
    }
    ok = all(checks.values())
    print("rank %d/%d segment [%d,%d): %s" % (rank, world, d.lo, d.hi, checks), flush=True)
    fp.close()
# the same first evaluation through the rank-local drop-in boundary: host P[]/SphP[] of this rank's segment
from gadget_leicester_b200 import aos  # noqa: E402
from gadget_leicester_b200.multigpu import DistributedDropIn  # noqa: E402
d2 = DistributedForcePath(par, rank=rank, world=world, device=local)
order2 = d2.decompose(ps)
seg = order2[d2.lo:d2.hi]
ps_seg = {k: (v[seg] if hasattr(v, "shape") and v.shape[0] == n else v) for k, v in ps.items()}
arr = aos.GadgetArrays(ps_seg, pinned=True)
DistributedDropIn(d2, arr).compute_accelerations()
r = arr.results()
ng = arr.N_gas
checks2 = {
    "dropin GravCost identical": bool((r["GravCost"] == res0["gravcost"]).all()),
    "dropin GravAccel identical": bool((r["GravAccel"] == res0["acc"]).all()),
    "dropin OldAcc identical": bool((r["OldAcc"] == res0["oldacc"]).all()),
}
if ng:
    sphsel = np.isin(ps_seg["type"], (0, 2, 5))
    checks2["dropin Hsml rel<=1e-12"] = bool(np.abs(r["Hsml"][sphsel] / res0["hsml"][sphsel] - 1).max() <= 1e-12)
    checks2["dropin Density rel<=1e-11"] = bool(np.abs(r["Density"] / res0["density"][:ng] - 1).max() <= 1e-11)
    checks2["dropin HydroAccel rel<=1e-8"] = bool((np.linalg.norm(r["HydroAccel"] - res0["hacc"][:ng], axis=1) /
                                                   np.linalg.norm(res0["hacc"][:ng], axis=1)).max() <= 1e-8)
    checks2["dropin DtEntropy/MaxSignalVel"] = bool(np.allclose(r["DtEntropy"], res0["dtentr"][:ng], rtol=1e-8, atol=0) and
                                                    np.allclose(r["MaxSignalVel"], res0["maxsig"][:ng], rtol=1e-12, atol=0))
ok = ok and all(checks2.values())
print("rank %d/%d drop-in segment: %s" % (rank, world, checks2), flush=True)
d2.close()
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU_CHECK", "PASS" if int(t.item()) == 1 else "FAIL")
d.close()
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
