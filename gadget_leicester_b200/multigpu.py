"""One-process-per-GPU sharding of the force path.

The reference partitions TARGETS by contiguous Peano-Hilbert key segments (domain.c:1075-1113) and, per
force phase, ships every target that touches a foreign top-tree leaf to the owner and the partial sums
back (gravtree.c:246/:307, density.c:258/:305, hydra.c:412/:463).  Inside one NVSwitch box the pattern is
inverted: SOURCES are replicated with one NCCL all-gather per phase, every GPU builds the identical
global tree, and walks only its own Peano-Hilbert segment of targets.  No result reduction exists
because every target is computed entirely by its owner, and because tree and per-target arithmetic are
the same on every GPU the results are bit-identical for any GPU count.

Exchange steps per force evaluation (the only collectives on the data path):
  1. all-gather of the owned particle state (pos, vel, mass, type, Hsml, OldAcc, time bins, entropy);
  2. after density: all-gather of Hsml and the SphP fields hydro reads from neighbours
     (Density, Pressure, DhsmlDensityFactor, DivVel, CurlVel) - replaces hydra.c:412 and the hmax
     exchange of forcetree.c:1722-1747.

torch / torch.distributed are used for device memory, streams and the NCCL plumbing only.
"""
import ctypes as C

import numpy as np
import torch

from . import lib as L
from .gadget import ForcePath

try:
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None


def partition_range(n, world, rank):
    """Contiguous equal-count segments of the Peano-Hilbert ordered particle list.  All segments have
    m = ceil(n/world) particles except the last ones, which may be shorter (or empty)."""
    m = -(-n // world)
    lo = min(n, rank * m)
    hi = min(n, lo + m)
    return lo, hi, m


def work_balanced_ranges(cost, world):
    """domain_findSplit_work_balanced (domain.c:1075-1113) for one contiguous segment per GPU: cut the
    Peano-Hilbert ordered list where the running cost crosses k/world of the total.  `cost` is the reference's
    per-particle weight (1 + GravCost) / 2^TimeBin (domain.c:378-384), in Peano-Hilbert order (numpy array).
    Returns a list of (lo, hi)."""
    c = np.cumsum(np.asarray(cost, dtype=np.float64))
    n = c.size
    tot = c[-1] if n else 0.0
    cuts = [0]
    for k in range(1, world):
        cuts.append(int(np.searchsorted(c, tot * k / world * (1 + 1e-12), side="right")) if n else 0)
    cuts.append(n)
    for k in range(1, len(cuts)):
        cuts[k] = max(cuts[k], cuts[k - 1])
    return [(cuts[k], cuts[k + 1]) for k in range(world)]


def gather_ragged(own, ranges, group=None):
    """all-gather segments of different lengths: pad every segment to the longest one, gather, then drop the
    padding.  ranges: list of (lo, hi) for all ranks."""
    world = len(ranges)
    if world == 1:
        return own
    mmax = max(hi - lo for lo, hi in ranges)
    n = ranges[-1][1]
    send = own
    if own.shape[0] != mmax:
        send = torch.zeros((mmax,) + tuple(own.shape[1:]), dtype=own.dtype, device=own.device)
        send[: own.shape[0]] = own
    full = torch.empty((world * mmax,) + tuple(own.shape[1:]), dtype=own.dtype, device=own.device)
    dist.all_gather_into_tensor(full, send.contiguous(), group=group)
    if all(hi - lo == mmax for lo, hi in ranges[:-1]):
        return full[:n]            # only the last segment is short: the padding sits at the very end
    return torch.cat([full[r * mmax: r * mmax + (hi - lo)] for r, (lo, hi) in enumerate(ranges)], dim=0)


def gather_padded(own, n, m, group=None):
    """all-gather segments of m rows (the last one zero-padded) and return the first n rows.
    own: [<=m, ...] tensor of this rank's segment.  Works with NCCL (cuda) and gloo (cpu)."""
    world = dist.get_world_size(group) if (dist is not None and dist.is_initialized()) else 1
    if world == 1:
        return own[:n]
    shape = (m,) + tuple(own.shape[1:])
    send = own
    if own.shape[0] != m:
        send = torch.zeros(shape, dtype=own.dtype, device=own.device)
        send[: own.shape[0]] = own
    full = torch.empty((world * m,) + tuple(own.shape[1:]), dtype=own.dtype, device=own.device)
    dist.all_gather_into_tensor(full, send.contiguous(), group=group)
    return full[:n]


STATE_FIELDS = ("pos", "vel", "mass", "type", "hsml", "oldacc", "timebin", "ti_begstep", "entropy", "dtentropy")
SPH_FIELDS = ("hsml", "density", "pressure", "dhsml", "divvel", "curlvel")


class DistributedForcePath:
    """Owns this rank's Peano-Hilbert segment on the GPU and evaluates forces for it."""

    def __init__(self, params, rank=0, world=1, device=0, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.device = torch.device("cuda", device)
        self.fp = ForcePath(params, device=device)
        self.params = params
        self.n = 0
        self.own = {}
        self.full = {}

    # -- set-up: the analogue of domain_Decomposition + peano_hilbert_order (domain.c:100, peano.c:25) --------
    def decompose(self, ps, cost=None):
        """ps: the FULL particle set as host numpy arrays (every rank passes the same set).  Orders it along the
        Peano-Hilbert curve (gas block first, like the reference) and keeps this rank's segment on the GPU.
        cost: optional per-particle work weight in input order, (1 + GravCost) / 2^TimeBin as domain.c:378-384;
        None = equal counts (what the reference gets on its first step, GravCost = 1 for everybody)."""
        n = ps["pos"].shape[0]
        ngas = int((ps["type"] == 0).sum())
        self.fp.load(ps)
        order = self.fp.peano_hilbert_order(ngas)
        self.n, self.ngas = n, ngas
        if cost is None:
            self.ranges = [partition_range(n, self.world, r)[:2] for r in range(self.world)]
        else:
            self.ranges = work_balanced_ranges(np.asarray(cost)[order], self.world)
        lo, hi = self.ranges[self.rank]
        m = max(b - a for a, b in self.ranges)
        self.lo, self.hi, self.m = lo, hi, m
        seg = order[lo:hi]
        dev = self.device
        tb = ps.get("ti_begstep", np.zeros(n, np.int32))
        self.own = dict(
            pos=torch.from_numpy(ps["pos"][seg]).to(dev), vel=torch.from_numpy(ps["vel"][seg]).to(dev),
            mass=torch.from_numpy(ps["mass"][seg]).to(dev), type=torch.from_numpy(ps["type"][seg].astype(np.int32)).to(dev),
            hsml=torch.from_numpy(ps["hsml"][seg]).to(dev), oldacc=torch.from_numpy(ps["oldacc"][seg]).to(dev),
            timebin=torch.from_numpy(ps["timebin"][seg].astype(np.int32)).to(dev),
            ti_begstep=torch.from_numpy(tb[seg].astype(np.int32)).to(dev),
            entropy=torch.from_numpy(ps["entropy"][seg]).to(dev), dtentropy=torch.from_numpy(ps["dtentropy"][seg]).to(dev))
        self.targets = torch.arange(lo, hi, dtype=torch.int32, device=dev)
        self.order = order
        # result buffers (full length, only [lo,hi) is meaningful on this rank)
        f64 = lambda *s: torch.zeros(*s, dtype=torch.float64, device=dev)
        self.res = dict(acc=f64(n, 3), oldacc=f64(n), gravcost=torch.zeros(n, dtype=torch.float32, device=dev),
                        hsml=f64(n), numngb=f64(n), density=f64(n), dhsml=f64(n), divvel=f64(n), curlvel=f64(n),
                        pressure=f64(n), hacc=f64(n, 3), dtentr=f64(n), maxsig=f64(n))
        return order

    def decompose_device(self, ps):
        """Same as decompose() for a particle set that already lives on this GPU as torch tensors (every rank
        generated the same set): nothing of size N ever touches the host."""
        n = ps["pos"].shape[0]
        ngas = int((ps["type"] == 0).sum().item())
        assert bool((ps["type"][:ngas] == 0).all().item()), "gas must come first"
        dev = self.device
        self.n, self.ngas = n, ngas
        torch.cuda.current_stream(dev).synchronize()
        self.fp.load_device(n, [ps[k].data_ptr() for k in STATE_FIELDS])
        order = torch.empty(n, dtype=torch.int32, device=dev)
        self.fp.peano_hilbert_order_device(ngas, order.data_ptr())
        self.ranges = [partition_range(n, self.world, r)[:2] for r in range(self.world)]
        lo, hi, m = partition_range(n, self.world, self.rank)
        self.lo, self.hi, self.m = lo, hi, m
        seg = order[lo:hi].long()
        self.own = {k: ps[k].index_select(0, seg).contiguous() for k in STATE_FIELDS}
        self.targets = torch.arange(lo, hi, dtype=torch.int32, device=dev)
        self.order = order
        f64 = lambda *s: torch.zeros(*s, dtype=torch.float64, device=dev)
        self.res = dict(acc=f64(n, 3), oldacc=f64(n), gravcost=torch.zeros(n, dtype=torch.float32, device=dev))
        if ngas > 0:
            self.res.update(hsml=f64(n), numngb=f64(n), density=f64(n), dhsml=f64(n), divvel=f64(n), curlvel=f64(n),
                            pressure=f64(n), hacc=f64(n, 3), dtentr=f64(n), maxsig=f64(n))
        return order

    def _sync_lib(self):
        # the library runs on its own stream and synchronises before returning; make torch's stream wait for
        # nothing further (results are complete), and make sure torch work is done before the library reads
        torch.cuda.current_stream(self.device).synchronize()

    # -- one force evaluation ------------------------------------------------------------------------------
    def exchange_state(self):
        """Exchange step 1: replicate the sources (NCCL all-gather over NVLink)."""
        for k in STATE_FIELDS:
            self.full[k] = gather_ragged(self.own[k], self.ranges, self.group).contiguous()
        self._sync_lib()
        f = self.full
        ptrs = [f[k].data_ptr() for k in STATE_FIELDS]
        self.fp.load_device(self.n, ptrs)
        lib, ctx = self.fp._lib, self.fp.ctx
        self.fp._chk(lib.gl_set_targets(ctx, C.c_void_p(self.targets.data_ptr()), self.targets.numel(), 1))

    def gravity(self):
        self.fp.gravity_tree_nocopy()
        r = self.res
        lib, ctx = self.fp._lib, self.fp.ctx
        self.fp._chk(lib.gl_get_gravity(ctx, C.c_void_p(r["acc"].data_ptr()), C.c_void_p(r["oldacc"].data_ptr()),
                                        C.c_void_p(r["gravcost"].data_ptr()), 1))
        self.own["oldacc"] = r["oldacc"][self.lo:self.hi].clone()

    def density_and_exchange(self):
        """density() for the owned targets, then exchange step 2."""
        self.fp.density_nocopy()
        r = self.res
        lib, ctx = self.fp._lib, self.fp.ctx
        p = lambda t: C.c_void_p(t.data_ptr())
        self.fp._chk(lib.gl_get_density(ctx, p(r["hsml"]), p(r["numngb"]), p(r["density"]), p(r["dhsml"]), p(r["divvel"]),
                                        p(r["curlvel"]), p(r["pressure"]), None, None, None, 1))
        self.own["hsml"] = r["hsml"][self.lo:self.hi].clone()
        if self.world > 1:
            g = {k: gather_ragged(r[k][self.lo:self.hi], self.ranges, self.group).contiguous() for k in SPH_FIELDS}
            self._sync_lib()
            self.fp._chk(lib.gl_set_hsml(ctx, p(g["hsml"]), 1))
            self.fp._chk(lib.gl_set_sph_state(ctx, p(g["density"]), p(g["pressure"]), p(g["dhsml"]), p(g["divvel"]),
                                              p(g["curlvel"]), 1))
            self._keep = g

    def hydro(self):
        self.fp.force_update_hmax()
        self.fp.hydro_force_nocopy()
        r = self.res
        lib, ctx = self.fp._lib, self.fp.ctx
        self.fp._chk(lib.gl_get_hydro(ctx, C.c_void_p(r["hacc"].data_ptr()), C.c_void_p(r["dtentr"].data_ptr()),
                                      C.c_void_p(r["maxsig"].data_ptr()), 1))

    def compute_accelerations(self, sph=True, gravity=True):
        """accel.c:27 force part for this rank's segment, with the two exchange steps."""
        self.exchange_state()
        if gravity:
            self.gravity()
        else:
            self.fp.force_treebuild()
        if sph and self.ngas > 0:
            self.density_and_exchange()
            self.hydro()

    def close(self):
        self.fp.close()
