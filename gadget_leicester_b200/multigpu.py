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
# layout of one rank's all-gather block (include/gl_b200.h gl_load_particles_gathered): offsets in units of m
# (what the tree build and the gravity walk read comes first, so a gravity-only exchange sends a prefix)
F64_LAYOUT = (("pos", 0, 3), ("mass", 3, 1), ("hsml", 4, 1), ("oldacc", 5, 1), ("vel", 6, 3), ("entropy", 9, 1),
              ("dtentropy", 10, 1))
I32_LAYOUT = (("type", 0, 1), ("timebin", 1, 1), ("ti_begstep", 2, 1))
F64_ROWS, I32_ROWS = 11, 3
F64_ROWS_GRAV, I32_ROWS_GRAV = 6, 1


def packed_views(f64, i32, m):
    """Named views into one rank's send block: pos/vel as [m,3], the rest as [m]."""
    v = {}
    for name, off, w in F64_LAYOUT:
        t = f64[off * m:(off + w) * m]
        v[name] = t.view(m, 3) if w == 3 else t
    for name, off, w in I32_LAYOUT:
        v[name] = i32[off * m:(off + w) * m]
    return v


def unpack_gathered(full_f64, full_i32, ranges, m):
    """Pure-torch statement of what gl_load_particles_gathered does with the all-gather receive buffers
    (used by the CPU tests to pin the layout): returns the global arrays as a dict."""
    out = {}
    for name, off, w in F64_LAYOUT:
        parts = []
        for r, (lo, hi) in enumerate(ranges):
            blk = full_f64[r * F64_ROWS * m:(r + 1) * F64_ROWS * m][off * m:(off + w) * m]
            parts.append((blk.view(m, 3) if w == 3 else blk)[: hi - lo])
        out[name] = torch.cat(parts, dim=0)
    for name, off, w in I32_LAYOUT:
        out[name] = torch.cat([full_i32[r * I32_ROWS * m:(r + 1) * I32_ROWS * m][off * m:(off + w) * m][: hi - lo]
                               for r, (lo, hi) in enumerate(ranges)], dim=0)
    return out


def allgather_block(own, world, group=None):
    """One collective for a whole block: own [k] -> [world*k]."""
    if world == 1:
        return own
    full = torch.empty(world * own.numel(), dtype=own.dtype, device=own.device)
    dist.all_gather_into_tensor(full, own, group=group)
    return full


class DistributedForcePath:
    """Owns this rank's Peano-Hilbert segment on the GPU and evaluates forces for it.

    Per force evaluation the data path has exactly three collectives (NCCL all-gathers over NVLink): the packed
    float64 state block, the packed int32 state block, and - after density - the packed SphP block."""

    def __init__(self, params, rank=0, world=1, device=0, group=None, force_path=None):
        """device: CUDA device index.  force_path: an object with ForcePath's interface to drive instead of the CUDA
        library - only the CPU (gloo) tests of the exchange choreography pass one, together with device="cpu"."""
        self.rank, self.world, self.group = rank, world, group
        self.device = torch.device(device) if isinstance(device, (str, torch.device)) else torch.device("cuda", device)
        self.fp = force_path if force_path is not None else ForcePath(params, device=device)
        self.params = params
        self.n = 0
        self.own = {}
        self.split_exchange = True   # send what only density / hydro read behind the tree build and the walk
        self._late = None
        self.t = {}          # wall-clock seconds per sub-phase of the last compute_accelerations (diagnostics)

    # -- set-up: the analogue of domain_Decomposition + peano_hilbert_order (domain.c:100, peano.c:25) --------
    def _adopt(self, n, ngas, order, ranges, take):
        """take(name, lo, hi) -> this rank's rows of a state field as a device tensor, in Peano-Hilbert order."""
        dev = self.device
        self.n, self.ngas, self.order, self.ranges = n, ngas, order, ranges
        lo, hi = ranges[self.rank]
        m = max(b - a for a, b in ranges)
        self.lo, self.hi, self.m = lo, hi, m
        self.seg_start = (C.c_longlong * (self.world + 1))(*([a for a, _ in ranges] + [n]))
        self.own_f64 = torch.zeros(F64_ROWS * m, dtype=torch.float64, device=dev)
        self.own_i32 = torch.zeros(I32_ROWS * m, dtype=torch.int32, device=dev)
        self.own = packed_views(self.own_f64, self.own_i32, m)
        for k in STATE_FIELDS:
            self.own[k][: hi - lo] = take(k, lo, hi)
        self.sph_send = torch.zeros(len(SPH_FIELDS) * m, dtype=torch.float64, device=dev)
        self.targets = torch.arange(lo, hi, dtype=torch.int32, device=dev)
        f64 = lambda *s: torch.zeros(*s, dtype=torch.float64, device=dev)
        # result buffers (full length, only [lo,hi) is meaningful on this rank)
        self.res = dict(acc=f64(n, 3), oldacc=f64(n), gravcost=torch.zeros(n, dtype=torch.float32, device=dev))
        if ngas > 0:
            self.res.update(hsml=f64(n), numngb=f64(n), density=f64(n), dhsml=f64(n), divvel=f64(n), curlvel=f64(n),
                            pressure=f64(n), hacc=f64(n, 3), dtentr=f64(n), maxsig=f64(n))
        return order

    def decompose(self, ps, cost=None):
        """ps: the FULL particle set as host numpy arrays (every rank passes the same set).  Orders it along the
        Peano-Hilbert curve (gas block first, like the reference) and keeps this rank's segment on the GPU.
        cost: optional per-particle work weight in input order, (1 + GravCost) / 2^TimeBin as domain.c:378-384;
        None = equal counts (what the reference gets on its first step, GravCost = 1 for everybody)."""
        n = ps["pos"].shape[0]
        ngas = int((ps["type"] == 0).sum())
        self.fp.load(ps)
        order = self.fp.peano_hilbert_order(ngas)
        if cost is None:
            ranges = [partition_range(n, self.world, r)[:2] for r in range(self.world)]
        else:
            ranges = work_balanced_ranges(np.asarray(cost)[order], self.world)
        tb = ps.get("ti_begstep", np.zeros(n, np.int32))
        src = dict(ps, ti_begstep=tb)
        i32 = ("type", "timebin", "ti_begstep")

        def take(k, lo, hi):
            a = np.asarray(src[k])[order[lo:hi]]
            return torch.from_numpy(np.ascontiguousarray(a.astype(np.int32) if k in i32 else a.astype(np.float64))).to(self.device)
        return self._adopt(n, ngas, order, ranges, take)

    def decompose_device(self, ps):
        """Same as decompose() for a particle set that already lives on this GPU as torch tensors (every rank
        generated the same set): nothing of size N ever touches the host."""
        n = ps["pos"].shape[0]
        ngas = int((ps["type"] == 0).sum().item())
        assert bool((ps["type"][:ngas] == 0).all().item()), "gas must come first"
        torch.cuda.current_stream(self.device).synchronize()
        self.fp.load_device(n, [ps[k].data_ptr() for k in STATE_FIELDS])
        order = torch.empty(n, dtype=torch.int32, device=self.device)
        self.fp.peano_hilbert_order_device(ngas, order.data_ptr())
        ranges = [partition_range(n, self.world, r)[:2] for r in range(self.world)]
        return self._adopt(n, ngas, order, ranges, lambda k, lo, hi: ps[k].index_select(0, order[lo:hi].long()))

    def _sync_lib(self):
        # the library runs on its own stream and synchronises before returning; make sure torch / NCCL work is
        # done before the library reads it
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()

    def _tick(self, key, t0):
        import time
        t1 = time.perf_counter()
        self.t[key] = self.t.get(key, 0.0) + (t1 - t0)
        return t1

    # -- one force evaluation ------------------------------------------------------------------------------
    def exchange_state(self, full=True):
        """Exchange step 1: replicate the sources (two NCCL all-gathers over NVLink: the float64 and the int32
        block), unpacked into the library's input arrays by one kernel.  full=False (gravity only): just the
        prefix of the blocks the tree build and the walk read (44 instead of 100 bytes per particle)."""
        import time
        t0 = time.perf_counter()
        m = self.m
        self._late = None
        split = full and self.world > 1 and self.split_exchange and self.device.type == "cuda"
        nf, ni = (F64_ROWS, I32_ROWS) if (full and not split) else (F64_ROWS_GRAV, I32_ROWS_GRAV)
        self.full_f64 = allgather_block(self.own_f64[: nf * m], self.world, self.group)
        self.full_i32 = allgather_block(self.own_i32[: ni * m], self.world, self.group)
        if split:
            # the part only density / hydro read travels while the tree is built and walked
            ev = torch.cuda.Event()
            ev.record()
            rest_f = torch.empty(self.world * (F64_ROWS - nf) * m, dtype=torch.float64, device=self.device)
            rest_i = torch.empty(self.world * (I32_ROWS - ni) * m, dtype=torch.int32, device=self.device)
            w1 = dist.all_gather_into_tensor(rest_f, self.own_f64[nf * m:], group=self.group, async_op=True)
            w2 = dist.all_gather_into_tensor(rest_i, self.own_i32[ni * m:], group=self.group, async_op=True)
            self._late = (rest_f, rest_i, w1, w2)
            ev.synchronize()              # the prefix has arrived; the rest is still in flight
        else:
            self._sync_lib()
        t0 = self._tick("allgather_state", t0)
        lib, ctx = self.fp._lib, self.fp.ctx
        self.fp._chk(lib.gl_load_particles_gathered(ctx, self.n, self.world, m, self.seg_start,
                                                    C.c_void_p(self.full_f64.data_ptr()), C.c_void_p(self.full_i32.data_ptr()),
                                                    1 if (full and not split) else 0))
        self.fp._chk(lib.gl_set_result_range(ctx, self.lo, self.hi - self.lo))   # this rank only ever reads its own results
        self.fp.N = self.n
        # the segment is consecutive along the Peano-Hilbert curve already: the library keeps the order
        import os
        setter = lib.gl_set_targets if os.environ.get("GL_NO_ORDERED") else lib.gl_set_targets_ordered
        self.fp._chk(setter(ctx, C.c_void_p(self.targets.data_ptr()), self.targets.numel(), 1))
        self._tick("load_state", t0)

    def gravity(self, fetch=True):
        import time
        t0 = time.perf_counter()
        self.fp.gravity_tree_nocopy()
        t0 = self._tick("gravity_tree", t0)
        if not fetch:
            return
        r = self.res
        lib, ctx = self.fp._lib, self.fp.ctx
        self.fp._chk(lib.gl_get_gravity(ctx, C.c_void_p(r["acc"].data_ptr()), C.c_void_p(r["oldacc"].data_ptr()),
                                        C.c_void_p(r["gravcost"].data_ptr()), 1))
        self.own["oldacc"][: self.hi - self.lo] = r["oldacc"][self.lo:self.hi]
        self._sync_lib()
        self._tick("get_gravity", t0)

    def finish_exchange(self):
        """Second half of a split state exchange: hand the SPH-side inputs to the library (before density)."""
        if self._late is None:
            return
        import time
        t0 = time.perf_counter()
        rest_f, rest_i, w1, w2 = self._late
        w1.wait()
        w2.wait()
        self._sync_lib()
        lib, ctx = self.fp._lib, self.fp.ctx
        self.fp._chk(lib.gl_set_sph_inputs_gathered(ctx, self.world, self.m, self.seg_start, C.c_void_p(rest_f.data_ptr()),
                                                    C.c_void_p(rest_i.data_ptr())))
        self._late = None
        self._tick("late_state", t0)

    def density_and_exchange(self):
        """density() for the owned targets, then exchange step 2 (one all-gather of the packed SphP block)."""
        import time
        self.finish_exchange()
        t0 = time.perf_counter()
        self.fp.density_nocopy()
        t0 = self._tick("density", t0)
        r = self.res
        lib, ctx = self.fp._lib, self.fp.ctx
        p = lambda t: C.c_void_p(t.data_ptr())
        self.fp._chk(lib.gl_get_density(ctx, p(r["hsml"]), p(r["numngb"]), p(r["density"]), p(r["dhsml"]), p(r["divvel"]),
                                        p(r["curlvel"]), p(r["pressure"]), None, None, None, 1))
        k = self.hi - self.lo
        self.own["hsml"][:k] = r["hsml"][self.lo:self.hi]
        t0 = self._tick("get_density", t0)
        if self.world > 1:
            m = self.m
            for j, f in enumerate(SPH_FIELDS):
                self.sph_send[j * m: j * m + k] = r[f][self.lo:self.hi]
            self.full_sph = allgather_block(self.sph_send, self.world, self.group)
            self._sync_lib()
            t0 = self._tick("allgather_sph", t0)
            self.fp._chk(lib.gl_set_sph_state_gathered(ctx, self.world, m, self.seg_start, p(self.full_sph)))
            self._tick("load_sph", t0)
        else:
            self._sync_lib()

    def hydro(self):
        import time
        t0 = time.perf_counter()
        self.fp.force_update_hmax()
        self.fp.hydro_force_nocopy()
        t0 = self._tick("hydro", t0)
        r = self.res
        lib, ctx = self.fp._lib, self.fp.ctx
        self.fp._chk(lib.gl_get_hydro(ctx, C.c_void_p(r["hacc"].data_ptr()), C.c_void_p(r["dtentr"].data_ptr()),
                                      C.c_void_p(r["maxsig"].data_ptr()), 1))
        self._tick("get_hydro", t0)

    def compute_accelerations(self, sph=True, gravity=True):
        """accel.c:27 force part for this rank's segment, with the two exchange steps."""
        self.t = {}
        self.exchange_state(full=bool(sph and self.ngas > 0))
        if gravity:
            self.gravity()
        else:
            self.fp.force_treebuild()
        if sph and self.ngas > 0:
            self.density_and_exchange()
            self.hydro()
        self.finish_exchange()

    def rebalance(self, timed=False):
        """domain_findSplit_work_balanced (domain.c:1075-1113) with the measured cost of the last gravity call:
        weight (1 + GravCost) / 2^TimeBin (domain.c:378-384), one contiguous Peano-Hilbert segment per GPU.
        Three all-gathers (cost, and the current float64 / int32 state blocks); every rank then cuts its new
        segment out of the replicated state locally - the analogue of domain_exchange (domain.c:596)."""
        dev, m, k = self.device, self.m, self.hi - self.lo
        send = torch.zeros(m, dtype=torch.float64, device=dev)
        send[:k] = 1.0 + self.res["gravcost"][self.lo:self.hi].double()
        send[:k] /= torch.pow(2.0, self.own["timebin"][:k].clamp(min=0).double())
        full = allgather_block(send, self.world, self.group).view(self.world, m)
        if timed:
            # cost -> time: scale every rank's weights by (its measured kernel time) / (its summed cost), the role the
            # measured CPU time plays in the reference's cost adjustment; a GPU whose segment marches through more
            # tree entries per interaction then gets fewer particles
            stt = self.fp.stats()
            mine = torch.tensor([stt["ms_gravity_walk"] + stt["ms_density"] + stt["ms_hydro"], float(send[:k].sum().item())],
                                dtype=torch.float64, device=dev)
            allt = allgather_block(mine, self.world, self.group).view(self.world, 2)
            scale = allt[:, 0] / allt[:, 1].clamp(min=1e-300)
            full = full * (scale / scale.mean()).view(self.world, 1)
        cost = torch.cat([full[r, : hi - lo] for r, (lo, hi) in enumerate(self.ranges)])
        new_ranges = work_balanced_ranges(cost.cpu().numpy(), self.world)      # (the function the gloo/CPU tests cover)
        del full, cost
        glob = unpack_gathered(allgather_block(self.own_f64, self.world, self.group),
                               allgather_block(self.own_i32, self.world, self.group), self.ranges, m)
        self.full_f64 = self.full_i32 = None
        self._adopt(self.n, self.ngas, self.order, new_ranges, lambda name, lo, hi: glob[name][lo:hi])
        return new_ranges

    def close(self):
        self.fp.close()


class LetForcePath(DistributedForcePath):
    """The multi-rank force path with the exchange INSIDE the library (include/gl_b200.h "multi-rank mode",
    csrc/gl_let.cu): every rank loads its own Peano-Hilbert segment, gl_comm_init attaches an NCCL communicator, and one
    gl_compute_accelerations call per step does the rest - own tree, locally-essential exchange over NVLink, merged
    tree, gravity, density, SphP exchange, hydro.  No Python between the phases.

    torch.distributed is used once, to broadcast the 128-byte NCCL id (the reference would MPI_Bcast it)."""

    def __init__(self, params, rank=0, world=1, device=0, group=None):
        super().__init__(params, rank=rank, world=world, device=device, group=group)
        self._comm = False

    def _attach(self):
        if self._comm or self.world == 1:
            return
        lib = self.fp._lib
        idt = torch.zeros(128, dtype=torch.uint8, device=self.device)
        if self.rank == 0:
            buf = C.create_string_buffer(128)
            if lib.gl_comm_unique_id(C.cast(buf, C.c_void_p)) != 0:
                raise L.GLError("gl_comm_unique_id failed")
            idt.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        dist.broadcast(idt, src=0, group=self.group)
        self.fp.comm_init(self.rank, self.world, bytes(idt.cpu().numpy().tobytes()))
        self._comm = True

    def _adopt(self, n, ngas, order, ranges, take):
        order = super()._adopt(n, ngas, order, ranges, take)
        self._attach()
        k = self.hi - self.lo
        f64 = lambda *s: torch.zeros(*s, dtype=torch.float64, device=self.device)
        # results of the own particles only, indexed like the own segment
        self.own_res = dict(acc=f64(k, 3), oldacc=f64(k), gravcost=torch.zeros(k, dtype=torch.float32, device=self.device))
        if ngas > 0:
            self.own_res.update(hsml=f64(k), numngb=f64(k), density=f64(k), dhsml=f64(k), divvel=f64(k), curlvel=f64(k),
                                pressure=f64(k), hacc=f64(k, 3), dtentr=f64(k), maxsig=f64(k))
        self._load_own()
        return order

    def _load_own(self):
        """gl_load_particles of the own segment (device pointers into the packed block the base class keeps)."""
        k = self.hi - self.lo
        self._sync_lib()
        self.fp.load_device(k, [self.own[f].data_ptr() for f in STATE_FIELDS])

    def compute_accelerations(self, sph=True, gravity=True):
        import time
        self.t = {}
        t0 = time.perf_counter()
        lib, ctx = self.fp._lib, self.fp.ctx
        if self.world > 1 and not gravity:
            self.fp.comm_set_options(False, 1.3)
        if gravity and sph and self.ngas > 0:
            self.fp._chk(lib.gl_compute_accelerations(ctx, 0))
            self.fp._relative_criterion_engaged()
        else:
            if gravity:
                self.fp.gravity_tree_nocopy()
            else:
                self.fp.force_treebuild()
            if sph and self.ngas > 0:
                self.fp.density_nocopy()
                self.fp.force_update_hmax()
                self.fp.hydro_force_nocopy()
        self._tick("compute_accelerations", t0)

    def fetch(self):
        """Copy the own results into self.own_res (device) and keep Hsml / OldAcc of the own block current."""
        lib, ctx = self.fp._lib, self.fp.ctx
        r = self.own_res
        p = lambda t: C.c_void_p(t.data_ptr())
        self.fp._chk(lib.gl_get_gravity(ctx, p(r["acc"]), p(r["oldacc"]), p(r["gravcost"]), 1))
        if self.ngas > 0 and "hsml" in r:
            self.fp._chk(lib.gl_get_density(ctx, p(r["hsml"]), p(r["numngb"]), p(r["density"]), p(r["dhsml"]), p(r["divvel"]),
                                            p(r["curlvel"]), p(r["pressure"]), None, None, None, 1))
            self.fp._chk(lib.gl_get_hydro(ctx, p(r["hacc"]), p(r["dtentr"]), p(r["maxsig"]), 1))
        self._sync_lib()
        return r

    def rebalance(self, timed=False):
        raise NotImplementedError("LetForcePath keeps the decomposition it was given (bench.py re-splits through DistributedForcePath)")


class DistributedDropIn:
    """compute_accelerations() (accel.c:27, force part) on THIS RANK'S OWN host P[] / SphP[] - its Peano-Hilbert
    segment, the way every MPI rank of the reference holds its local particles - at any GPU count:
    pack + H2D of the segment, the all-gathers, the force path, and the rank's slice of every result block copied
    back and scattered while the next phase computes (gl_gadget_pack_segment / gl_gadget_push_segment_async)."""

    def __init__(self, dfp, arrays):
        from . import aos
        self.d, self.a = dfp, arrays
        self._lay = aos.layout_struct()
        fp = dfp.fp
        fp._chk(fp._lib.gl_gadget_bind(fp.ctx, C.byref(self._lay)))
        assert arrays.NumPart == dfp.hi - dfp.lo

    def bytes_per_step(self):
        n, ng = self.a.NumPart, self.a.N_gas
        h2d = self.d.m * (11 * 8 + 3 * 4)
        other = 5 * 8 if n > ng else 0
        d2h = n * ((3 + 1) * 8 + 4) + n * (7 * 8 + other) + n * 5 * 8
        return h2d, d2h

    def compute_accelerations(self):
        d, a = self.d, self.a
        fp = d.fp
        lib, ctx = fp._lib, fp.ctx
        P, S = L._ptr(a.P), L._ptr(a.SphP)
        fp._chk(lib.gl_gadget_pack_segment(ctx, P, S, a.NumPart, a.N_gas, d.m, d.lo, C.c_void_p(d.own_f64.data_ptr()),
                                           C.c_void_p(d.own_i32.data_ptr())))
        d.t = {}
        d.exchange_state()
        d.gravity(fetch=False)
        fp._chk(lib.gl_gadget_push_segment_async(ctx, P, S, 1))
        if d.ngas > 0:
            d.density_and_exchange()
            fp._chk(lib.gl_gadget_push_segment_async(ctx, P, S, 2))
            d.fp.force_update_hmax()
            d.fp.hydro_force_nocopy()
            fp._chk(lib.gl_gadget_push_segment_async(ctx, P, S, 4))
        fp._chk(lib.gl_gadget_push_wait(ctx))
