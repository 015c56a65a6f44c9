"""CPU suite: the N>1 exchange choreography of DistributedForcePath (decompose -> exchange_state -> gravity ->
density_and_exchange -> hydro -> rebalance) on the gloo backend, world_size 2.

The CUDA library is replaced by a recording stand-in that reads the very buffers the real entry points would get
(same pointers, same block layout as include/gl_b200.h documents for gl_load_particles_gathered /
gl_set_sph_state_gathered) - so what is checked is everything on the Python side of the C ABI: which rows every rank
sends, where they land in the replicated arrays, that every rank asks only for its own Peano-Hilbert segment, and that a
work-balanced re-split (domain.c:1075-1113) moves the particles - including the fields updated by the force
phases (OldAcc, Hsml) - to the right owner.  No compute is done here; the numerical parity of the same sequence is
scripts/multigpu_check.py and the -m gpu tests."""
import ctypes as C
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gadget_leicester_b200.multigpu import (DistributedForcePath, work_balanced_ranges, partition_range, F64_LAYOUT,
                                            I32_LAYOUT, F64_ROWS, I32_ROWS, F64_ROWS_GRAV, I32_ROWS_GRAV, SPH_FIELDS,
                                            STATE_FIELDS)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _view(ptr, count, dtype):
    addr = ptr.value if isinstance(ptr, C.c_void_p) else int(ptr)
    ct = {np.float64: C.c_double, np.int32: C.c_int32, np.float32: C.c_float}[dtype]
    return np.ctypeslib.as_array((ct * count).from_address(addr))


def _unpack(buf, layout, rows, world, m, seg):
    """global arrays out of an all-gather receive buffer: rank r's block starts at r*rows*m, field at off*m"""
    out = {}
    for name, off, w in layout:
        if off + w > rows:
            continue
        parts = []
        for r in range(world):
            k = seg[r + 1] - seg[r]
            blk = buf[r * rows * m + off * m: r * rows * m + (off + w) * m]
            parts.append(blk.reshape(m, 3)[:k] if w == 3 else blk[:k])
        out[name] = np.concatenate(parts, axis=0).copy()
    return out


class _RecordingLib:
    """Stand-in for libgl_b200.so: every entry point DistributedForcePath calls, reading / writing the caller's
    buffers like the real one, with results that are simple functions of the global Peano-Hilbert index."""

    def __init__(self, n, ngas):
        self.n, self.ngas = n, ngas
        self.calls = []
        self.state = {}
        self.sph = None
        self.lo = self.cnt = None
        self.targets = None

    @staticmethod
    def grav_cost(i):
        return (i % 7).astype(np.float32) * 3.0 + (i > i.max() // 2) * 20.0      # lopsided on purpose

    def gl_load_particles_gathered(self, ctx, n, world, m, seg, f64, i32, full):
        rf, ri = (F64_ROWS, I32_ROWS) if full else (F64_ROWS_GRAV, I32_ROWS_GRAV)
        seg = list(seg)
        assert seg[0] == 0 and seg[-1] == n == self.n
        self.state = _unpack(_view(f64, world * rf * m, np.float64), F64_LAYOUT, rf, world, m, seg)
        self.state.update(_unpack(_view(i32, world * ri * m, np.int32), I32_LAYOUT, ri, world, m, seg))
        self.calls.append(("load", int(full)))
        return 0

    def gl_set_result_range(self, ctx, lo, cnt):
        self.lo, self.cnt = int(lo), int(cnt)
        return 0

    def gl_set_targets_ordered(self, ctx, ptr, k, on_device):
        self.targets = _view(ptr, k, np.int32).copy()
        self.calls.append(("targets", int(k)))
        return 0

    gl_set_targets = gl_set_targets_ordered

    def _own(self):
        return np.arange(self.lo, self.lo + self.cnt)

    def gl_get_gravity(self, ctx, acc, oldacc, gravcost, on_device):
        i = self._own()
        _view(acc, self.n * 3, np.float64).reshape(self.n, 3)[i] = self.state["pos"][i] * 2.0
        _view(oldacc, self.n, np.float64)[i] = i + 0.5
        _view(gravcost, self.n, np.float32)[i] = self.grav_cost(np.arange(self.n))[i]
        self.calls.append(("get_gravity",))
        return 0

    def gl_get_density(self, ctx, hsml, numngb, density, dhsml, divvel, curlvel, pressure, a, b, c, on_device):
        own = self._own()
        i = own[own < self.ngas]
        _view(hsml, self.n, np.float64)[own] = self.state["hsml"][own]       # Hsml comes back for every particle ...
        for k, p in enumerate((hsml, numngb, density, dhsml, divvel, curlvel, pressure)):
            _view(p, self.n, np.float64)[i] = 1000.0 * (k + 1) + i           # ... the iteration only changes the gas
        self.calls.append(("get_density",))
        return 0

    def gl_set_sph_state_gathered(self, ctx, world, m, seg, ptr):
        rows = len(SPH_FIELDS)
        lay = tuple((f, j, 1) for j, f in enumerate(SPH_FIELDS))
        self.sph = _unpack(_view(ptr, world * rows * m, np.float64), lay, rows, world, m, list(seg))
        self.calls.append(("sph_state",))
        return 0

    def gl_get_hydro(self, ctx, hacc, dtentr, maxsig, on_device):
        i = self._own()
        i = i[i < self.ngas]
        _view(dtentr, self.n, np.float64)[i] = -1.0 - i
        self.calls.append(("get_hydro",))
        return 0


class _StubForcePath:
    """ForcePath's interface as far as DistributedForcePath drives it."""

    def __init__(self, n, ngas, order):
        self._lib = _RecordingLib(n, ngas)
        self.ctx = None
        self._order = order
        self.N = 0
        self.phases = []

    def _chk(self, rc):
        assert rc == 0

    def load(self, ps):
        self.loaded_n = ps["pos"].shape[0]

    def peano_hilbert_order(self, ngas):
        return self._order

    def gravity_tree_nocopy(self):
        self.phases.append("gravity")

    def force_treebuild(self):
        self.phases.append("treebuild")

    def density_nocopy(self):
        self.phases.append("density")

    def force_update_hmax(self):
        self.phases.append("hmax")

    def hydro_force_nocopy(self):
        self.phases.append("hydro")

    def close(self):
        pass


def _particle_set(n, ngas, seed=5):
    g = np.random.default_rng(seed)
    ps = dict(pos=g.random((n, 3)), vel=g.standard_normal((n, 3)), mass=g.random(n) + 0.1,
              type=np.where(np.arange(n) < ngas, 0, 1).astype(np.int32), hsml=g.random(n), oldacc=g.random(n),
              timebin=g.integers(0, 3, n).astype(np.int32), ti_begstep=g.integers(0, 64, n).astype(np.int32),
              entropy=g.random(n), dtentropy=g.standard_normal(n))
    order = np.concatenate([g.permutation(ngas), ngas + g.permutation(n - ngas)]).astype(np.int32)   # gas block first
    return ps, order


def _sph_expected(n, ngas, hsml_in):
    """what the recording library's density results look like once every rank's rows have been exchanged"""
    i = np.arange(n, dtype=np.float64)
    gas = np.arange(n) < ngas
    col = dict(hsml=1, numngb=2, density=3, dhsml=4, divvel=5, curlvel=6, pressure=7)
    return {f: np.where(gas, 1000.0 * col[f] + i, hsml_in if f == "hsml" else 0.0) for f in SPH_FIELDS}


def _check_state(lib, ps, order, expect_hsml=None, expect_oldacc=None, full=True):
    names = STATE_FIELDS if full else ("pos", "mass", "hsml", "oldacc", "type")
    for k in names:
        want = np.asarray(ps[k])[order]
        if k == "hsml" and expect_hsml is not None:
            want = expect_hsml
        if k == "oldacc" and expect_oldacc is not None:
            want = expect_oldacc
        assert np.array_equal(lib.state[k], want), k
    if not full:
        assert "vel" not in lib.state and "timebin" not in lib.state       # the gravity-only exchange sends a prefix


def _worker(rank, world, port, n, ngas, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        ps, order = _particle_set(n, ngas)
        fp = _StubForcePath(n, ngas, order)
        lib = fp._lib
        d = DistributedForcePath(None, rank=rank, world=world, device="cpu", force_path=fp)
        got_order = d.decompose(ps)
        assert np.array_equal(got_order, order) and fp.loaded_n == n
        lo, hi, _ = partition_range(n, world, rank)
        assert (d.lo, d.hi) == (lo, hi)

        # -- one full force evaluation on equal-count segments
        d.compute_accelerations()
        _check_state(lib, ps, order)
        assert (lib.lo, lib.cnt) == (lo, hi - lo)
        assert np.array_equal(lib.targets, np.arange(lo, hi, dtype=np.int32))
        assert fp.phases == ["gravity", "density", "hmax", "hydro"]
        assert [c[0] for c in lib.calls] == ["load", "targets", "get_gravity", "get_density", "sph_state", "get_hydro"]
        want = _sph_expected(n, ngas, np.asarray(ps["hsml"])[order])
        for f in SPH_FIELDS:                      # every rank now holds BOTH ranks' density results
            assert np.array_equal(lib.sph[f], want[f]), f
        # results of the own segment only
        own = np.arange(lo, hi)
        assert np.array_equal(d.res["oldacc"].numpy()[own], own + 0.5)
        other = np.setdiff1d(np.arange(n), own)
        assert not d.res["oldacc"].numpy()[other].any()

        # -- work-balanced re-split from the measured cost; the state moves with it
        pidx = np.arange(n)
        cost = (1.0 + lib.grav_cost(pidx).astype(np.float64)) / 2.0 ** np.asarray(ps["timebin"])[order]
        want_ranges = work_balanced_ranges(cost, world)
        new_ranges = d.rebalance()
        assert new_ranges == want_ranges and want_ranges != [partition_range(n, world, r)[:2] for r in range(world)]
        nlo, nhi = new_ranges[rank]
        assert (d.lo, d.hi) == (nlo, nhi) and d.m == max(b - a for a, b in new_ranges)
        gas = pidx < ngas
        hs_now = np.where(gas, 1000.0 + pidx, np.asarray(ps["hsml"])[order])     # density wrote Hsml of the gas
        oa_now = pidx + 0.5                                                      # gravity wrote OldAcc of everybody
        k = nhi - nlo
        assert np.array_equal(d.own["hsml"][:k].numpy(), hs_now[nlo:nhi])
        assert np.array_equal(d.own["oldacc"][:k].numpy(), oa_now[nlo:nhi])
        assert np.array_equal(d.own["pos"][:k].numpy(), ps["pos"][order][nlo:nhi])
        assert np.array_equal(d.own["ti_begstep"][:k].numpy(), ps["ti_begstep"][order][nlo:nhi])
        assert np.array_equal(d.targets.numpy(), np.arange(nlo, nhi, dtype=np.int32))

        # -- next evaluation on the ragged segments
        lib.calls.clear()
        fp.phases.clear()
        d.compute_accelerations()
        _check_state(lib, ps, order, expect_hsml=hs_now, expect_oldacc=oa_now)
        assert (lib.lo, lib.cnt) == (nlo, nhi - nlo)
        for f in SPH_FIELDS:
            assert np.array_equal(lib.sph[f], want[f]), f

        # -- gravity-only evaluation: prefix blocks, no SPH phases
        lib.calls.clear()
        fp.phases.clear()
        d.compute_accelerations(sph=False)
        _check_state(lib, ps, order, expect_hsml=hs_now, expect_oldacc=oa_now, full=False)
        assert lib.calls[0] == ("load", 0) and fp.phases == ["gravity"]
        q.put((rank, None))
    except Exception:       # noqa: BLE001 - reported to the parent
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_distributed_force_path_choreography_gloo_world2():
    world, n, ngas = 2, 1001, 700
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, ngas, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(err is None for _, err in res), "\n".join(str(e) for _, e in res if e)
