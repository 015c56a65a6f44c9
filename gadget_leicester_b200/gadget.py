"""Host-side mirror of the reference's operator interface for the force path.

The reference exposes global C symbols acting on global arrays (proto.h:281,295,331,332;
forcetree.h:48,67,86).  `ForcePath` keeps the same entry-point names, argument meaning and
side effects, but on an explicit object: `P` is a dict of numpy arrays laid out like P[]/SphP[]
(see ics.py), and every method calls straight into libgl_b200.so (the CUDA path).  Errors from
the library raise GLError - the analogue of endrun(code) (endrun.c:23).
"""
import ctypes as C

import numpy as np

from . import lib as L


class ForcePath:
    def __init__(self, params, device=0, max_part=0):
        self._lib = L.load_library()
        self._ctx = C.c_void_p()
        if self._lib.gl_create(C.byref(self._ctx), int(device), int(max_part)) != 0:
            raise L.GLError("gl_create failed: no usable CUDA device (libgl_b200 has no CPU fallback)")
        self.params = dict(params)
        self._gp = L.make_params(self.params)
        self._chk(self._lib.gl_set_params(self._ctx, C.byref(self._gp)))
        self.N = 0
        self.P = None

    # -- plumbing ------------------------------------------------------------------------
    def _chk(self, rc):
        if rc != 0:
            raise L.GLError(self._lib.gl_last_error(self._ctx).decode())

    def close(self):
        if self._ctx:
            self._lib.gl_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def ctx(self):
        return self._ctx

    def set_ErrTolTheta(self, theta):
        """All.ErrTolTheta (gravtree.c:396-397 sets it to 0 after the first call)."""
        self.params["ErrTolTheta"] = float(theta)
        self._gp.ErrTolTheta = theta
        self._chk(self._lib.gl_set_params(self._ctx, C.byref(self._gp)))

    def set_params(self, **kw):
        self.params.update(kw)
        self._gp = L.make_params(self.params)
        self._chk(self._lib.gl_set_params(self._ctx, C.byref(self._gp)))

    # -- particle state ------------------------------------------------------------------
    def load(self, P):
        """Upload P[]/SphP[] (host numpy arrays, gas first).  The analogue of the state the
        reference's entry points find in its globals when they are called."""
        self.P = P
        n = P["pos"].shape[0]
        self.N = n
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        self._keep = [f64(P["pos"]), f64(P["vel"]), f64(P["mass"]), i32(P["type"]), f64(P["hsml"]), f64(P["oldacc"]),
                      i32(P["timebin"]), i32(P.get("ti_begstep", np.zeros(n, np.int32))), f64(P["entropy"]),
                      f64(P["dtentropy"])]
        self._chk(self._lib.gl_load_particles(self._ctx, n, *[L._ptr(a) for a in self._keep], 0))

    def load_device(self, n, ptrs):
        """Same as load() but from device pointers (ints), in gl_load_particles order; 0 = absent."""
        self.N = int(n)
        self._chk(self._lib.gl_load_particles(self._ctx, int(n), *[C.c_void_p(p) if p else None for p in ptrs], 1))

    def set_OldAcc(self, oldacc):
        """Overwrite P[].OldAcc (input of the relative opening criterion, forcetree.c:1850)."""
        a = np.ascontiguousarray(oldacc, dtype=np.float64)
        assert a.size == self.N
        self._chk(self._lib.gl_set_oldacc(self._ctx, L._ptr(a), 0))

    def set_Hsml(self, hsml):
        """Overwrite P[].Hsml."""
        a = np.ascontiguousarray(hsml, dtype=np.float64)
        assert a.size == self.N
        self._chk(self._lib.gl_set_hsml(self._ctx, L._ptr(a), 0))

    def set_targets(self, index=None):
        """Active list (FirstActiveParticle/NextActiveParticle) or a GPU's Peano-Hilbert segment."""
        if index is None:
            self._chk(self._lib.gl_set_targets(self._ctx, None, -1, 0))
        else:
            idx = np.ascontiguousarray(index, dtype=np.int32)
            self._chk(self._lib.gl_set_targets(self._ctx, L._ptr(idx), idx.size, 0))

    # -- domain / keys -------------------------------------------------------------------
    def domain_findExtent(self):
        out = np.zeros(8)
        self._chk(self._lib.gl_domain_find_extent(self._ctx, out.ctypes.data_as(L.c_double_p)))
        return dict(DomainLen=out[0], DomainFac=out[1], DomainCorner=out[2:5].copy(), DomainCenter=out[5:8].copy())

    def domain_setExtent(self, xmin, xmax):
        a = np.ascontiguousarray(xmin, dtype=np.float64)
        b = np.ascontiguousarray(xmax, dtype=np.float64)
        self._chk(self._lib.gl_domain_set_extent(self._ctx, a.ctypes.data_as(L.c_double_p),
                                                 b.ctypes.data_as(L.c_double_p)))

    def peano_and_morton_keys(self):
        ph = np.zeros(self.N, dtype=np.uint64)
        mo = np.zeros(self.N, dtype=np.uint64)
        self._chk(self._lib.gl_peano_keys(self._ctx, L._ptr(ph), L._ptr(mo), 0))
        return ph, mo

    def peano_hilbert_order(self, n_gas):
        order = np.zeros(self.N, dtype=np.int32)
        self._chk(self._lib.gl_peano_hilbert_order(self._ctx, int(n_gas), L._ptr(order), 0))
        return order

    def peano_hilbert_order_device(self, n_gas, out_ptr):
        """peano_hilbert_order into a device int32 buffer (pointer as int)."""
        self._chk(self._lib.gl_peano_hilbert_order(self._ctx, int(n_gas), C.c_void_p(out_ptr), 1))

    # -- tree ----------------------------------------------------------------------------
    def force_treebuild(self):
        nn = C.c_longlong(0)
        self._chk(self._lib.gl_force_treebuild(self._ctx, C.byref(nn)))
        return nn.value

    ngb_treebuild = force_treebuild  # ngb.c:315: the neighbour tree IS the gravity tree

    def set_TreeReconstructFlag(self, flag=1):
        self._lib.gl_set_tree_reconstruct_flag(self._ctx, int(flag))

    def nodes(self):
        st = self.stats()
        tab = np.zeros((st["numnodes"], 12))
        if st["numnodes"]:
            self._chk(self._lib.gl_get_nodes(self._ctx, tab.ctypes.data_as(L.c_double_p), st["numnodes"]))
        return tab

    # -- the force entry points ------------------------------------------------------------
    def _relative_criterion_engaged(self):
        """gravtree.c:396-397: after a force computation with TypeOfOpeningCriterion == 1 the reference leaves
        All.ErrTolTheta = 0 for good.  The library does that in its own copy of the parameters; mirror it here so that
        a later set_params() (Ti_Current per step, ...) does not switch back to the Barnes-Hut criterion."""
        if int(self.params.get("TypeOfOpeningCriterion", 0)) == 1:
            self.params["ErrTolTheta"] = 0.0
            self._gp.ErrTolTheta = 0.0

    def gravity_tree(self):
        self._chk(self._lib.gl_gravity_tree(self._ctx))
        self._relative_criterion_engaged()
        acc = np.zeros((self.N, 3))
        old = np.zeros(self.N)
        cost = np.zeros(self.N, dtype=np.float32)
        self._chk(self._lib.gl_get_gravity(self._ctx, L._ptr(acc), L._ptr(old), L._ptr(cost), 0))
        return dict(GravAccel=acc, OldAcc=old, GravCost=cost)

    def density(self):
        self._chk(self._lib.gl_density(self._ctx))
        n = self.N
        names = ["Hsml", "NumNgb", "Density", "DhsmlDensityFactor", "DivVel", "CurlVel", "Pressure", "SinkDensity",
                 "SinkEntropy"]
        out = {k: np.zeros(n) for k in names}
        out["SinkGasVel"] = np.zeros((n, 3))
        self._chk(self._lib.gl_get_density(self._ctx, *[L._ptr(out[k]) for k in names], L._ptr(out["SinkGasVel"]), 0))
        dp = np.zeros(n, dtype=np.int32)
        self._chk(self._lib.gl_get_pair_counts(self._ctx, L._ptr(dp), None, 0))
        out["DensPairs"] = dp
        return out

    def force_update_hmax(self):
        self._chk(self._lib.gl_force_update_hmax(self._ctx))

    def hydro_force(self):
        self._chk(self._lib.gl_hydro_force(self._ctx))
        n = self.N
        out = dict(HydroAccel=np.zeros((n, 3)), DtEntropy=np.zeros(n), MaxSignalVel=np.zeros(n))
        self._chk(self._lib.gl_get_hydro(self._ctx, L._ptr(out["HydroAccel"]), L._ptr(out["DtEntropy"]),
                                         L._ptr(out["MaxSignalVel"]), 0))
        hp = np.zeros(n, dtype=np.int32)
        self._chk(self._lib.gl_get_pair_counts(self._ctx, None, L._ptr(hp), 0))
        out["HydroPairs"] = hp
        return out

    def compute_accelerations(self, mode=0):
        """accel.c:27 (force part): gravity (twice at Ti_Current==0), density, hmax, hydro."""
        self._chk(self._lib.gl_compute_accelerations(self._ctx, int(mode)))
        self._relative_criterion_engaged()

    def compute_potential(self):
        """potential.c:22: P[].p.Potential for all particles with the current tree / OldAcc / criterion."""
        self._chk(self._lib.gl_compute_potential(self._ctx))
        pot = np.zeros(self.N)
        self._chk(self._lib.gl_get_potential(self._ctx, L._ptr(pot), 0))
        return pot

    # -- the sub-grid neighbour loops behind the force path (accel.c:189-196), on the tree of the step -------------------
    def set_subgrid_params(self, **kw):
        """All.SMBHmass, InnerBoundary, SinkBoundary, ViscousTimescale, BlackHoleFeedbackFactor, the unit system,
        MeanWeight, MinEgySpec.  Units default to the shipped run's (Pd_N16e5_*.txt:86-88); derived units as begrun.c."""
        q = dict(SMBHmass=0.5, InnerBoundary=2.0e-2, SinkBoundary=1.0e-4, SofteningBndry=self.params["Softening"][5],
                 ViscousTimescale=3.15e9, BlackHoleFeedbackFactor=0.0, UnitLength_in_cm=1.496e15, UnitMass_in_g=1.989e33,
                 UnitVelocity_in_cm_per_s=297837.66, MeanWeight=2.45, MinEgySpec=0.0)
        q.update(kw)
        q.setdefault("UnitTime_in_s", q["UnitLength_in_cm"] / q["UnitVelocity_in_cm_per_s"])
        q.setdefault("UnitDensity_in_cgs", q["UnitMass_in_g"] / q["UnitLength_in_cm"] ** 3)
        q.setdefault("UnitEnergy_in_cgs", q["UnitMass_in_g"] * q["UnitLength_in_cm"] ** 2 / q["UnitTime_in_s"] ** 2)
        sp = L.GLSubgridParams()
        for k, v in q.items():
            if k == "xbh":
                sp.xbh[:] = v
            else:
                setattr(sp, k, float(v))
        self._chk(self._lib.gl_set_subgrid_params(self._ctx, C.byref(sp)))
        self.subgrid = q

    @staticmethod
    def _active(active):
        if active is None:
            return None, 0, None
        a = np.ascontiguousarray(active, dtype=np.int32)
        return L._ptr(a), a.shape[0], a

    def blackhole_accretion(self, S, active=None):
        """blackhole.c:70.  S: dict of P[]-indexed arrays Vel, ID, SwallowID, BH_Mass, BH_Mdot, AccDisc_Mass, Dust_Mass,
        Total_Mass, Injected_BH_Energy (float64 / int32, updated in place) and optionally SinkDensity.  Returns the new
        P.Mass and the swallowed counts (gas, sinks, dust)."""
        n = self.N
        st = L.GLSinkState()
        keep = []
        for field, key, dt in (("vel", "Vel", np.float64), ("id", "ID", np.int32), ("swallow_id", "SwallowID", np.int32),
                               ("bh_mass", "BH_Mass", np.float64), ("bh_mdot", "BH_Mdot", np.float64),
                               ("accdisc_mass", "AccDisc_Mass", np.float64), ("dust_mass", "Dust_Mass", np.float64),
                               ("total_mass", "Total_Mass", np.float64), ("injected_bh_energy", "Injected_BH_Energy", np.float64),
                               ("sink_density", "SinkDensity", np.float64)):
            a = S.get(key)
            if a is None:
                continue
            if a.dtype != dt or not a.flags["C_CONTIGUOUS"] or a.shape[0] != n:
                raise ValueError("blackhole_accretion: %s must be a contiguous %s array of %d rows" % (key, dt.__name__, n))
            keep.append(a)
            setattr(st, field, a.ctypes.data)
        mass = np.zeros(n)
        st.mass = mass.ctypes.data
        ap, na, ak = self._active(active)
        cnt = (C.c_longlong * 3)()
        self._chk(self._lib.gl_blackhole_accretion(self._ctx, ap, na, C.byref(st), 0, cnt))
        return mass, tuple(int(x) for x in cnt)

    def dust_density(self, out, active=None):
        """dust.c:60: P.d7.DUST_particle_density of the active dust grains, written into `out` (float64[n])."""
        ap, na, ak = self._active(active)
        self._chk(self._lib.gl_dust_density(self._ctx, ap, na, L._ptr(out), 0))
        return out

    def dust_drag(self, S, active=None):
        """dust.c:263.  S: dict of P[]-indexed float64 arrays Vel, DustParticleDensity, DustRadius, DustVcoll,
        DeltaDustMomentum, NewDragAcc, DeltaDragEnergy, DragHeating (updated in place), optionally GravAccel, GasDensity,
        GasEntropy, GasVel, DustParticleVelocity, Entropy."""
        n = self.N
        st = L.GLDustState()
        keep = []
        for field, key in (("vel", "Vel"), ("gravaccel", "GravAccel"), ("gas_density", "GasDensity"), ("gas_entropy", "GasEntropy"),
                           ("gas_vel", "GasVel"), ("dust_particle_density", "DustParticleDensity"),
                           ("dust_particle_velocity", "DustParticleVelocity"), ("dust_radius", "DustRadius"), ("dust_vcoll", "DustVcoll"),
                           ("delta_dust_momentum", "DeltaDustMomentum"), ("new_drag_acc", "NewDragAcc"),
                           ("delta_drag_energy", "DeltaDragEnergy"), ("entropy", "Entropy"), ("drag_heating", "DragHeating")):
            a = S.get(key)
            if a is None:
                continue
            if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"] or a.shape[0] != n:
                raise ValueError("dust_drag: %s must be a contiguous float64 array of %d rows" % (key, n))
            keep.append(a)
            setattr(st, field, a.ctypes.data)
        ap, na, ak = self._active(active)
        self._chk(self._lib.gl_dust_drag(self._ctx, ap, na, C.byref(st), 0))

    # -- run()'s loop on the device-resident state (run.c:63-110; include/gl_b200.h "run()'s loop") -----------------------
    def set_sph_state(self, Density, Pressure, DhsmlDensityFactor, DivVel, CurlVel):
        """SphP[] state a density() leaves behind, P[]-indexed float64[n] (rows of non-gas particles are ignored)."""
        keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (Density, Pressure, DhsmlDensityFactor, DivVel, CurlVel)]
        assert all(a.size == self.N for a in keep)
        self._chk(self._lib.gl_set_sph_state(self._ctx, *[L._ptr(a) for a in keep], 0))

    def set_VelPred(self, velpred):
        """SphP[].VelPred of the leading gas block (n_gas x 3)."""
        a = np.ascontiguousarray(velpred, dtype=np.float64)
        self._chk(self._lib.gl_set_velpred(self._ctx, L._ptr(a), a.shape[0], 0))

    def set_accelerations(self, GravAccel=None, HydroAccel=None, MaxSignalVel=None):
        """P[].g.GravAccel, SphP[].a.HydroAccel (n x 3), SphP[].MaxSignalVel (n), all P[]-indexed."""
        f = lambda a, w: None if a is None else np.ascontiguousarray(a, dtype=np.float64).reshape(self.N, w) if w > 1 else \
            np.ascontiguousarray(a, dtype=np.float64).reshape(self.N)
        keep = [f(GravAccel, 3), f(HydroAccel, 3), f(MaxSignalVel, 1)]
        self._chk(self._lib.gl_set_accelerations(self._ctx, *[L._ptr(a) for a in keep], 0))

    @staticmethod
    def timestep_params(Ti_Current, TimeBinActive, **kw):
        """gl_timestep_params from All.* values (names as in allvars.h); TimeBinActive: iterable of 0/1 per bin."""
        t = L.GLTimestepParams()
        t.Ti_Current = int(Ti_Current)
        for b, v in enumerate(TimeBinActive):
            t.TimeBinActive[b] = int(v)
        t.GammaMinus1 = 7. / 5. - 1
        t.p_DeltaDustMomentum = t.p_NewDensity = -1
        for k, v in kw.items():
            if k == "SofteningTable":
                for i in range(6):
                    t.SofteningTable[i] = float(v[i])
            else:
                setattr(t, k, float(v))
        return t

    def run_begin(self, Ti_Current):
        self._chk(self._lib.gl_run_begin(self._ctx, int(Ti_Current)))

    def advance_and_find_timesteps(self, tpar, commit_dtentropy=True):
        """timestep.c:29 on the resident state, for the particles on the bins tpar.TimeBinActive marks."""
        self._chk(self._lib.gl_run_kick(self._ctx, C.byref(tpar), 1 if commit_dtentropy else 0))

    def run(self, tpar, nsteps):
        """nsteps iterations of run.c:63-110 on the device; tpar.Ti_Current / TimeBinActive advance in place."""
        info = L.GLRunInfo()
        self._chk(self._lib.gl_run_steps(self._ctx, C.byref(tpar), int(nsteps), C.byref(info)))
        self.params["Ti_Current"] = int(tpar.Ti_Current)
        self._gp.Ti_Current = int(tpar.Ti_Current)
        self._relative_criterion_engaged()
        return {k: getattr(info, k) for k, _ in L.GLRunInfo._fields_}

    def run_set_ids(self, ids):
        a = np.ascontiguousarray(ids, dtype=np.int32)
        assert a.size == self.N
        self._chk(self._lib.gl_run_set_ids(self._ctx, L._ptr(a), 0))

    def run_ids(self):
        a = np.zeros(self.N, dtype=np.int32)
        self._chk(self._lib.gl_get_run_ids(self._ctx, L._ptr(a), 0))
        return a

    def domain_exchange(self, by_cost=True):
        """domain_exchange (domain.c:665) of the resident state; collective.  Returns (particles held now, particles sent)."""
        nn, ns = C.c_longlong(0), C.c_longlong(0)
        self._chk(self._lib.gl_domain_exchange(self._ctx, 1 if by_cost else 0, C.byref(nn), C.byref(ns)))
        self.N = int(nn.value)
        return int(nn.value), int(ns.value)

    def run_state(self):
        """P[]-indexed state of the resident loop (VelPred: a row per particle, meaningful for gas)."""
        n = self.N
        out = dict(Pos=np.zeros((n, 3)), Vel=np.zeros((n, 3)), VelPred=np.zeros((n, 3)), Hsml=np.zeros(n), Entropy=np.zeros(n),
                   DtEntropy=np.zeros(n), TimeBin=np.zeros(n, np.int32), Ti_begstep=np.zeros(n, np.int32))
        self._chk(self._lib.gl_get_run_state(self._ctx, *[L._ptr(out[k]) for k in ("Pos", "Vel", "VelPred", "Hsml", "Entropy", "DtEntropy",
                                                                                     "TimeBin", "Ti_begstep")], 0))
        return out

    def results(self):
        """P[]-indexed results of the last force computations, without running anything."""
        n = self.N
        g = dict(GravAccel=np.zeros((n, 3)), OldAcc=np.zeros(n), GravCost=np.zeros(n, np.float32))
        self._chk(self._lib.gl_get_gravity(self._ctx, L._ptr(g["GravAccel"]), L._ptr(g["OldAcc"]), L._ptr(g["GravCost"]), 0))
        names = ["Hsml", "NumNgb", "Density", "DhsmlDensityFactor", "DivVel", "CurlVel", "Pressure"]
        d = {k: np.zeros(n) for k in names}
        self._chk(self._lib.gl_get_density(self._ctx, *[L._ptr(d[k]) for k in names], None, None, None, 0))
        h = dict(HydroAccel=np.zeros((n, 3)), DtEntropy=np.zeros(n), MaxSignalVel=np.zeros(n))
        self._chk(self._lib.gl_get_hydro(self._ctx, L._ptr(h["HydroAccel"]), L._ptr(h["DtEntropy"]), L._ptr(h["MaxSignalVel"]), 0))
        g.update(d)
        g.update({"HydroAccel": h["HydroAccel"], "HydroDtEntropy": h["DtEntropy"], "MaxSignalVel": h["MaxSignalVel"]})
        return g

    # device-resident variants used by bench.py's `value` leg (no result download)
    def gravity_tree_nocopy(self):
        self._chk(self._lib.gl_gravity_tree(self._ctx))
        self._relative_criterion_engaged()

    def density_nocopy(self):
        self._chk(self._lib.gl_density(self._ctx))

    def hydro_force_nocopy(self):
        self._chk(self._lib.gl_hydro_force(self._ctx))

    # -- multi-rank mode (include/gl_b200.h "multi-rank mode") ---------------------------------
    def comm_init(self, rank, world, unique_id):
        """NCCL communicator inside the library; unique_id = the 128 bytes of gl_comm_unique_id(), broadcast by the caller."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._chk(self._lib.gl_comm_init(self._ctx, int(rank), int(world), C.cast(buf, C.c_void_p)))
        self._world = int(world)

    def comm_init_local(self, rank, group, world=None):
        """Attach to an in-process group (gl_comm_local_group_create): several contexts of one process as ranks."""
        self._chk(self._lib.gl_comm_init_local(self._ctx, int(rank), group))
        if world is not None:
            self._world = int(world)

    def comm_set_options(self, want_gravity=True, search_padding=1.25):
        self._chk(self._lib.gl_comm_set_options(self._ctx, 1 if want_gravity else 0, float(search_padding)))

    def comm_finalize(self):
        self._chk(self._lib.gl_comm_finalize(self._ctx))

    def domain_split_collective(self, cost=None):
        """SURVEY 8(e) partition: where each loaded particle belongs when the global Peano-Hilbert order is cut into
        `world` stretches of equal cost.  Returns (dest[n], to_go[world], splitters[world-1]); collective."""
        n = self.N
        dest = np.zeros(n, dtype=np.int32)
        world = self._world if hasattr(self, "_world") else 1
        togo = (C.c_longlong * max(world, 1))()
        spl = np.zeros(max(world - 1, 1), dtype=np.uint64)
        cp = None
        if cost is not None:
            cost = np.ascontiguousarray(cost, dtype=np.float64)
            assert cost.size == n
            cp = L._ptr(cost)
        self._chk(self._lib.gl_domain_split_collective(self._ctx, cp, 0, L._ptr(dest), 0, togo, L._ptr(spl)))
        return dest, np.array(list(togo), dtype=np.int64), spl[: max(world - 1, 0)]

    def let_stats(self):
        st = L.GLLetStats()
        self._chk(self._lib.gl_get_let_stats(self._ctx, C.byref(st)))
        return {k: getattr(st, k) for k, _ in L.GLLetStats._fields_}

    def stats(self):
        s = L.GLStats()
        self._chk(self._lib.gl_get_stats(self._ctx, C.byref(s)))
        return {k: getattr(s, k) for k, _ in L.GLStats._fields_}

    def reset_launch_count(self):
        self._lib.gl_reset_launch_count(self._ctx)
