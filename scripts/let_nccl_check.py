#!/usr/bin/env python
"""Two (or more) GPUs over NCCL: every rank loads its Peano-Hilbert segment of a disc, the library exchanges the
locally-essential tree (gl_comm_init), and the results are compared with one GPU running the whole set (rank 0).
Launch with torchrun; prints IDENTICAL or MISMATCH."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gadget_leicester_b200 import ics                              # noqa: E402
from gadget_leicester_b200.gadget import ForcePath                # noqa: E402
from gadget_leicester_b200.multigpu import LetForcePath           # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", dest="n", type=int, default=200000)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    par = ics.default_params()
    ps = ics.disc(a.n, seed=42)
    ps["hsml"] = ics.guess_hsml(ps, par["DesNumNgb"])
    ps["hsml"][ps["type"] == 5] = par["Softening"][5]
    n = ps["pos"].shape[0]
    d = LetForcePath(par, rank=rank, world=world, device=local)
    d.decompose(ps)
    d.compute_accelerations()            # Barnes-Hut pass + relative pass, cold-start density, hydro
    d.fp.set_TreeReconstructFlag(1)
    d.compute_accelerations()            # a settled step
    mine = d.fetch()
    ref = torch.zeros(n, 4, dtype=torch.float64, device="cuda")
    if rank == 0:
        pso = {k: (v[d.order] if hasattr(v, "shape") and v.shape[0] == n else v) for k, v in ps.items()}
        one = ForcePath(par)
        one.load(pso)
        one.compute_accelerations()
        one.set_TreeReconstructFlag(1)
        one.compute_accelerations()
        import ctypes as C
        acc = np.zeros((n, 3)); old = np.zeros(n); cost = np.zeros(n, np.float32)
        one._chk(one._lib.gl_get_gravity(one.ctx, acc.ctypes.data_as(C.c_void_p), old.ctypes.data_as(C.c_void_p), cost.ctypes.data_as(C.c_void_p), 0))
        one.close()
        ref[:, 0] = torch.from_numpy(cost.astype(np.float64)).cuda()
        ref[:, 1:4] = torch.from_numpy(acc).cuda()
    dist.broadcast(ref, src=0)
    own = ref[d.lo:d.hi]
    bad = torch.stack([(mine["gravcost"].double() != own[:, 0]).sum(), (mine["acc"] != own[:, 1:4]).any(dim=1).sum()]).double()
    dist.all_reduce(bad)
    # the collective partition over the same communicator: every rank must arrive at the same splitter keys, and the
    # stretches must carry equal numbers of particles (equal costs)
    dest, togo, spl = d.fp.domain_split_collective()
    sp = torch.from_numpy(spl.astype(np.int64)).cuda()
    allsp = [torch.zeros_like(sp) for _ in range(world)]
    dist.all_gather(allsp, sp)
    cnt = torch.from_numpy(np.bincount(dest, minlength=world).astype(np.float64)).cuda()
    dist.all_reduce(cnt)
    same = all(bool((q == sp).all()) for q in allsp)
    share = (cnt / cnt.sum()).cpu().numpy()
    part_ok = same and float(np.abs(share - 1.0 / world).max()) < 0.005
    bad = torch.cat([bad, torch.tensor([0.0 if part_ok else 1.0], dtype=torch.float64, device="cuda")])
    if rank == 0:
        print("collective partition: splitters agree %s, shares %s" % (same, np.round(share, 4)))
        st = d.fp.let_stats()
        print("rank 0 imported %d particles + %d cells of %d foreign particles; violations %d" % (
            st["imported_particles"], st["imported_cells"], n - st["own_particles"], st["violations"]))
        print("IDENTICAL" if float(bad.sum()) == 0 else "MISMATCH %s" % bad.cpu().numpy())
    d.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if float(bad.sum()) == 0 else 1)


if __name__ == "__main__":
    main()
