"""How much of the global tree does one GPU's Peano-Hilbert segment really need?  (Planning aid for the locally-essential-
tree exchange of DESIGN.md section 8 - NOT part of the product and not a parity test.)

The particle set is ordered along the Peano-Hilbert curve and cut into K segments like multigpu.py does.  For every
segment the global oct-tree is walked once with a GROUP form of the reference's relative opening criterion
(forcetree.c:2085-2105): a cell is shipped as ONE record if every target in the segment's bounding box would accept it
(mass * len^2 <= d_min^4 * a_min with d_min the distance from the cell's centre of mass to the box, a_min the smallest
ErrTolForceAcc * OldAcc in the segment, and the box clear of the 0.6-len guard cube); otherwise its daughters are looked
at; single particles are shipped as particles.  Cells made only of the segment's own particles cost nothing.  The SPH
import is the set of foreign gas particles within max(h_i) of the box or with the box inside their own h_j.
OldAcc comes from the oracle binary (TEST INFRASTRUCTURE, used here as a tool).  Prints one JSON line.

    python scripts/let_census.py [--ngas 200000] [--ranks 8]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util                                          # noqa: E402
from gadget_leicester_b200 import ics, lib as L      # noqa: E402


def keys(pos, corner, fac, which):
    lib = L.load_library()
    f = getattr(lib, which)
    f.restype = C.c_ulonglong
    g = ((pos - corner) * fac).astype(np.int64)
    return np.array([f(int(x), int(y), int(z), 21) for x, y, z in g], dtype=np.uint64)


def box_dist2(p, lo, hi):
    """squared distance from point p to the nearest of the boxes lo[k], hi[k]"""
    d = np.maximum(np.maximum(lo - p, p - hi), 0.0)
    return float((d * d).sum(axis=-1).min())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ngas", type=int, default=200000)
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--sub", type=int, default=64, help="boxes per segment")
    args = ap.parse_args()
    ps = ics.disc(args.ngas, seed=42)
    par = ics.default_params()
    par["PartAllocFactor"] = 2.5
    binary = util.REF_HARNESS if util.have_ref() else util.PORT_BIN
    o = util.run_oracle(binary, ps, par, "T S D R G G")
    n = ps["pos"].shape[0]
    idmap = {int(i): k for k, i in enumerate(ps["id"])}
    perm = np.array([idmap[int(i)] for i in o["5_G_id"]])
    oldacc = np.empty(n)
    oldacc[perm] = o["5_G_OldAcc"]
    hsml = np.empty(n)
    hperm = np.array([idmap[int(i)] for i in o["2_D_id"]])
    hsml[:] = 0
    hsml[hperm[: o["2_D_Hsml"].shape[0]]] = o["2_D_Hsml"]
    pos, mass, typ = ps["pos"], ps["mass"], ps["type"]
    lo_all, hi_all = pos.min(0), pos.max(0)
    length = 1.001 * (hi_all - lo_all).max()
    corner = 0.5 * (lo_all + hi_all) - 0.5 * length
    fac = (1 << 21) / length
    ph = keys(pos, corner, fac, "gl_peano_hilbert_key")
    mo = keys(pos, corner, fac, "gl_morton_key")
    # owner of every particle: equal-count segments of the Peano-Hilbert order
    K = args.ranks
    order_ph = np.argsort(ph, kind="stable")
    owner = np.empty(n, np.int32)
    for r in range(K):
        owner[order_ph[r * n // K:(r + 1) * n // K]] = r
    # Morton-sorted arrays for the tree
    om = np.argsort(mo, kind="stable")
    mk, mp, mm, mown = mo[om], pos[om], mass[om], owner[om]
    cm = np.vstack([np.zeros(3), np.cumsum(mp * mm[:, None], axis=0)])
    cmass = np.concatenate([[0.0], np.cumsum(mm)])
    cown = np.vstack([np.zeros(K, np.int64), np.cumsum(np.eye(K, dtype=np.int64)[mown], axis=0)])
    aerr = par["ErrTolForceAcc"]
    res = []
    for r in range(K):
        mine = owner == r
        # the segment as SUB consecutive pieces of the curve, each with its own box (one box around a whole segment of a
        # thin disc is mostly empty space)
        seg = order_ph[r * n // K:(r + 1) * n // K]
        pieces = np.array_split(seg, args.sub)
        blo = np.array([pos[q].min(0) for q in pieces])
        bhi = np.array([pos[q].max(0) for q in pieces])
        amin = aerr * oldacc[mine].min()
        nodes = parts = 0
        stack = [(0, n, 0, corner + 0.5 * length, length)]
        while stack:
            a, b, level, ctr, ln = stack.pop()
            cnt = b - a
            if cnt == 0:
                continue
            own = cown[b, r] - cown[a, r]
            if own == cnt:
                continue                                  # the segment's own particles: nothing to import
            if cnt == 1:
                parts += 1
                continue
            m = cmass[b] - cmass[a]
            s = (cm[b] - cm[a]) / m
            d2 = box_dist2(s, blo, bhi)
            guard_clear = bool(np.all(np.any(np.maximum(blo - ctr, ctr - bhi) > 0.6 * ln, axis=1)))
            if m * ln * ln <= d2 * d2 * amin and guard_clear:
                nodes += 1
                continue
            if level == 21:
                parts += cnt - own
                continue
            shift = np.uint64(3 * (20 - level))
            base = (mk[a] >> (shift + np.uint64(3))) << np.uint64(3)
            sub = a
            for j in range(8):
                nxt = a + int(np.searchsorted(mk[a:b], (base + np.uint64(j + 1)) << shift, side="left")) if j < 7 else b
                off = np.array([(j & 1) - 0.5, ((j >> 1) & 1) - 0.5, ((j >> 2) & 1) - 0.5]) * (ln / 2)
                stack.append((sub, nxt, level + 1, ctr + off, ln / 2))
                sub = nxt
        # SPH import, exactly: foreign gas j with some own gas i inside max(h_i, h_j) (hydra.c:1266)
        from scipy.spatial import cKDTree
        gas_other = np.flatnonzero((~mine) & (typ == 0))
        gas_own = np.flatnonzero(mine & (typ == 0))
        need = np.zeros(gas_other.size, bool)
        if gas_own.size and gas_other.size:
            d, _ = cKDTree(pos[gas_own]).query(pos[gas_other])
            need |= d < hsml[gas_other]
            hits = cKDTree(pos[gas_other]).query_ball_point(pos[gas_own], r=hsml[gas_own])
            for h in hits:
                need[h] = True
        sph = int(need.sum())
        res.append(dict(rank=r, own=int(mine.sum()), let_nodes=nodes, let_particles=parts, sph_import=sph))
    tot_let = np.mean([x["let_nodes"] * 64 + x["let_particles"] * 44 for x in res])
    tot_sph = np.mean([x["sph_import"] * 100 for x in res])
    out = dict(particles=n, ranks=K, boxes_per_segment=args.sub, per_rank=res, mean_let_bytes=float(tot_let), mean_sph_import_bytes=float(tot_sph),
               replicate_gravity_bytes=n * 44, replicate_full_bytes=n * 100,
               let_fraction_of_replication=float((tot_let + tot_sph) / (n * 100)))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
