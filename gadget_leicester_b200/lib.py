"""ctypes binding of libgl_b200.so (the C ABI declared in include/gl_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or no GPU is visible the
constructor raises.  Nothing under oracle/ is imported here.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GL_B200_LIB") or os.path.join(_HERE, "libgl_b200.so")   # GL_B200_LIB: A/B a second in-tree build

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_float_p = C.POINTER(C.c_float)
c_u64_p = C.POINTER(C.c_ulonglong)


class GLParams(C.Structure):
    _fields_ = [
        ("ErrTolTheta", C.c_double),
        ("ErrTolForceAcc", C.c_double),
        ("TypeOfOpeningCriterion", C.c_int),
        ("DesNumNgb", C.c_int),
        ("MaxNumNgbDeviation", C.c_double),
        ("BlackHoleNgbFactor", C.c_double),
        ("ArtBulkViscConst", C.c_double),
        ("G", C.c_double),
        ("ForceSoftening", C.c_double * 6),
        ("MinGasHsml", C.c_double),
        ("Timebase_interval", C.c_double),
        ("Ti_Current", C.c_int),
        ("ComovingIntegrationOn", C.c_int),
        ("OmegaLambdaHubble2", C.c_double),
        ("SofteningTable", C.c_double * 6),
    ]


class GLStats(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "ms_keys_sort", "ms_treebuild", "ms_moments", "ms_gravity_walk", "ms_density", "ms_hmax", "ms_hydro",
        "ms_h2d", "ms_d2h", "ms_potential")] + [(k, C.c_longlong) for k in (
            "interactions", "density_targets", "density_passes_max", "hydro_targets", "numnodes", "numentries",
            "density_fat_targets", "kernel_launches", "density_pairs", "density_target_passes", "hydro_pairs",
            "density_slowest_index", "density_slowest_passes")]


class GLLetStats(C.Structure):
    _fields_ = [(k, C.c_longlong) for k in ("own_particles", "exported_cells", "exported_particles", "imported_cells",
                                            "imported_particles", "builds", "violations")] + [("full", C.c_int)] + \
               [(k, C.c_double) for k in ("ms_own_tree", "ms_select", "ms_exchange", "ms_sph_exchange")]


class GLGadgetLayout(C.Structure):
    _fields_ = [("p_stride", C.c_longlong), ("sph_stride", C.c_longlong)] + [(k, C.c_int) for k in (
        "p_Pos", "p_Vel", "p_Mass", "p_GravAccel", "p_OldAcc", "p_Hsml", "p_NumNgb", "p_GravCost", "p_Ti_begstep",
        "p_Type", "p_TimeBin", "p_SinkDensity", "p_SinkEntropy", "p_SinkGasVel", "p_DustDensity", "p_DustEntropy",
        "p_DustGasVel", "p_SwallowID", "s_Entropy", "s_Pressure", "s_VelPred", "s_MaxSignalVel", "s_Density",
        "s_DtEntropy", "s_HydroAccel", "s_DhsmlDensityFactor", "s_DivVel", "s_CurlVel", "p_Ti_current", "s_DragAccel")]


class GLIOHeader(C.Structure):
    """struct io_header (allvars.h:1733), 256 bytes."""
    _fields_ = [("npart", C.c_int * 6), ("mass", C.c_double * 6), ("time", C.c_double), ("redshift", C.c_double),
                ("flag_sfr", C.c_int), ("flag_feedback", C.c_int), ("npartTotal", C.c_uint * 6), ("flag_cooling", C.c_int),
                ("num_files", C.c_int), ("BoxSize", C.c_double), ("Omega0", C.c_double), ("OmegaLambda", C.c_double),
                ("HubbleParam", C.c_double), ("flag_stellarage", C.c_int), ("flag_metals", C.c_int),
                ("npartTotalHighWord", C.c_uint * 6), ("flag_entropy_instead_u", C.c_int), ("flag_doubleprecision", C.c_int),
                ("fill", C.c_char * 56)]


class GLIOLayout(C.Structure):
    _fields_ = [(k, C.c_int) for k in (
        "p_ID", "p_DustRadius", "p_StellarAge", "p_BH_Mass", "p_BH_Mdot", "p_AccDisc_Mass", "p_BH_Mseed", "p_NewDensity",
        "p_DeltaPhotonMomentum", "p_OldPhotonMomentum", "p_BirthPhotonEnergy", "s_Ne", "s_accreted_mass", "s_accreted_momentum")]


class GLIOParams(C.Structure):
    _fields_ = [("format", C.c_int), ("build", C.c_int), ("CoolingOn", C.c_int), ("StarformationOn", C.c_int),
                ("RestartFlag", C.c_int), ("Ti_Current", C.c_int), ("Timebase_interval", C.c_double), ("Time", C.c_double),
                ("MassTable", C.c_double * 6), ("MinEgySpec", C.c_double), ("GammaMinus1", C.c_double),
                ("InitGasTemp", C.c_double), ("InitGasU", C.c_double), ("BoxSize", C.c_double), ("Omega0", C.c_double),
                ("OmegaLambda", C.c_double), ("HubbleParam", C.c_double)]


class GLTimestepParams(C.Structure):
    _fields_ = [("Ti_Current", C.c_int), ("TimeBinActive", C.c_int * 32), ("Timebase_interval", C.c_double),
                ("ErrTolIntAccuracy", C.c_double), ("CourantFac", C.c_double), ("MaxSizeTimestep", C.c_double),
                ("MinSizeTimestep", C.c_double), ("MinEgySpec", C.c_double), ("GammaMinus1", C.c_double),
                ("SofteningTable", C.c_double * 6), ("OuterBoundary", C.c_double), ("UnitVelocity_in_cm_per_s", C.c_double),
                ("FeedBackVelocity", C.c_double), ("SMBHmass", C.c_double), ("InnerBoundary", C.c_double),
                ("SinkBoundary", C.c_double), ("AccDtBlackHole", C.c_double), ("p_DeltaDustMomentum", C.c_int),
                ("p_NewDensity", C.c_int)]


class GLRunInfo(C.Structure):
    _fields_ = [("steps_done", C.c_int), ("Ti_Current", C.c_int), ("active_total", C.c_longlong), ("active_last", C.c_longlong),
                ("interactions", C.c_longlong), ("ms_total", C.c_double)]


class GLTopNode(C.Structure):
    """struct local_topnode_data (domain.c:48)."""
    _fields_ = [("Size", C.c_ulonglong), ("StartKey", C.c_ulonglong), ("Count", C.c_longlong), ("Cost", C.c_double),
                ("Daughter", C.c_int), ("Leaf", C.c_int), ("Parent", C.c_int), ("PIndex", C.c_int)]


class GLTreeLayout(C.Structure):
    _fields_ = [("node_stride", C.c_longlong), ("ext_stride", C.c_longlong)] + [(k, C.c_int) for k in (
        "n_len", "n_center", "n_maxsoft", "n_s", "n_mass", "n_bitflags", "n_sibling", "n_nextnode", "n_father", "n_Ti_current",
        "e_vs", "e_vmax", "e_divVmax", "e_hmax", "e_Ti_lastkicked", "e_Flag")] + [("bitflags_value", C.c_uint)]


class GLSubgridParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "SMBHmass", "InnerBoundary", "SinkBoundary", "SofteningBndry", "ViscousTimescale", "BlackHoleFeedbackFactor",
        "UnitTime_in_s", "UnitMass_in_g", "UnitLength_in_cm", "UnitVelocity_in_cm_per_s", "UnitEnergy_in_cgs", "UnitDensity_in_cgs",
        "MeanWeight", "MinEgySpec")] + [("xbh", C.c_double * 3)]


class GLSinkState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in (
        "vel", "id", "swallow_id", "mass", "bh_mass", "bh_mdot", "accdisc_mass", "dust_mass", "total_mass", "sink_density",
        "injected_bh_energy")]


class GLDustState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in (
        "vel", "gravaccel", "gas_density", "gas_entropy", "gas_vel", "dust_particle_density", "dust_particle_velocity", "dust_radius",
        "dust_vcoll", "delta_dust_momentum", "new_drag_acc", "delta_drag_energy", "entropy", "drag_heating")]


# every symbol include/gl_b200.h declares (tests check that the .so exports all of them)
EXPORTS = [
    "gl_create", "gl_destroy", "gl_last_error", "gl_set_params", "gl_stream", "gl_load_particles", "gl_set_oldacc",
    "gl_set_hsml", "gl_set_sph_state", "gl_domain_find_extent", "gl_domain_set_extent", "gl_peano_keys", "gl_peano_hilbert_order",
    "gl_peano_hilbert_key", "gl_morton_key", "gl_force_treebuild", "gl_get_nodes", "gl_set_targets",
    "gl_set_targets_ordered", "gl_load_particles_gathered", "gl_set_sph_state_gathered", "gl_set_result_range", "gl_set_sph_inputs_gathered",
    "gl_gravity_tree", "gl_density", "gl_force_update_hmax", "gl_hydro_force", "gl_compute_accelerations", "gl_compute_potential", "gl_get_potential",
    "gl_set_tree_reconstruct_flag", "gl_get_gravity", "gl_get_density", "gl_get_hydro", "gl_get_pair_counts",
    "gl_get_stats", "gl_reset_launch_count", "gl_gadget_bind", "gl_gadget_pull", "gl_gadget_push", "gl_gadget_compute_accelerations",
    "gl_gadget_pack_segment", "gl_gadget_push_segment_async", "gl_gadget_push_wait", "gl_export_tree", "gl_gadget_drift", "gl_gadget_reorder",
    "gl_snapshot_write", "gl_snapshot_read", "gl_snapshot_read_header", "gl_gadget_advance_and_find_timesteps", "gl_gadget_drift_list", "gl_gadget_peano_hilbert_order", "gl_domain_top_tree", "gl_gadget_domain_keys", "gl_gadget_costfactors", "gl_domain_split",
    "gl_comm_unique_id", "gl_comm_init", "gl_comm_local_group_create", "gl_comm_local_group_destroy", "gl_comm_init_local",
    "gl_comm_finalize", "gl_comm_set_options", "gl_get_let_stats",
    "gl_set_subgrid_params", "gl_blackhole_accretion", "gl_dust_density", "gl_dust_drag",
    "gl_gadget_blackhole_accretion", "gl_gadget_dust_density", "gl_gadget_dust_drag", "gl_hdf5_read_f64",
    "gl_domain_assign", "gl_domain_count_to_go", "gl_domain_split_collective",
    "gl_set_velpred", "gl_set_accelerations", "gl_run_begin", "gl_run_kick", "gl_run_steps", "gl_get_run_state", "gl_gadget_run_steps", "gl_next_sync_point", "gl_run_set_ids", "gl_get_run_ids", "gl_domain_exchange",
]

_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libgl_b200.so is missing (%s): build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C gadget_leicester_b200/csrc` - there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.gl_create.argtypes = [C.POINTER(vp), C.c_int, C.c_longlong]
    lib.gl_destroy.argtypes = [vp]
    lib.gl_destroy.restype = None
    lib.gl_last_error.argtypes = [vp]
    lib.gl_last_error.restype = C.c_char_p
    lib.gl_stream.argtypes = [vp]
    lib.gl_stream.restype = vp
    lib.gl_set_params.argtypes = [vp, C.POINTER(GLParams)]
    lib.gl_load_particles.argtypes = [vp, C.c_longlong] + [vp] * 10 + [C.c_int]
    lib.gl_set_oldacc.argtypes = [vp, vp, C.c_int]
    lib.gl_set_hsml.argtypes = [vp, vp, C.c_int]
    lib.gl_set_sph_state.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int]
    lib.gl_domain_find_extent.argtypes = [vp, c_double_p]
    lib.gl_domain_set_extent.argtypes = [vp, c_double_p, c_double_p]
    lib.gl_peano_keys.argtypes = [vp, vp, vp, C.c_int]
    lib.gl_peano_hilbert_order.argtypes = [vp, C.c_longlong, vp, C.c_int]
    lib.gl_peano_hilbert_key.argtypes = [C.c_int] * 4
    lib.gl_peano_hilbert_key.restype = C.c_ulonglong
    lib.gl_morton_key.argtypes = [C.c_int] * 4
    lib.gl_morton_key.restype = C.c_ulonglong
    lib.gl_force_treebuild.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.gl_get_nodes.argtypes = [vp, c_double_p, C.c_longlong]
    lib.gl_set_targets.argtypes = [vp, vp, C.c_longlong, C.c_int]
    lib.gl_set_targets_ordered.argtypes = [vp, vp, C.c_longlong, C.c_int]
    lib.gl_load_particles_gathered.argtypes = [vp, C.c_longlong, C.c_int, C.c_longlong, C.POINTER(C.c_longlong), vp, vp, C.c_int]
    lib.gl_set_result_range.argtypes = [vp, C.c_longlong, C.c_longlong]
    lib.gl_set_sph_inputs_gathered.argtypes = [vp, C.c_int, C.c_longlong, C.POINTER(C.c_longlong), vp, vp]
    lib.gl_set_sph_state_gathered.argtypes = [vp, C.c_int, C.c_longlong, C.POINTER(C.c_longlong), vp]
    lib.gl_get_potential.argtypes = [vp, vp, C.c_int]
    for f in ("gl_gravity_tree", "gl_density", "gl_force_update_hmax", "gl_hydro_force", "gl_compute_potential"):
        getattr(lib, f).argtypes = [vp]
    lib.gl_compute_accelerations.argtypes = [vp, C.c_int]
    lib.gl_set_tree_reconstruct_flag.argtypes = [vp, C.c_int]
    lib.gl_set_tree_reconstruct_flag.restype = None
    lib.gl_get_gravity.argtypes = [vp, vp, vp, vp, C.c_int]
    lib.gl_get_density.argtypes = [vp] + [vp] * 10 + [C.c_int]
    lib.gl_get_hydro.argtypes = [vp, vp, vp, vp, C.c_int]
    lib.gl_get_pair_counts.argtypes = [vp, vp, vp, C.c_int]
    lib.gl_get_stats.argtypes = [vp, C.POINTER(GLStats)]
    lib.gl_get_let_stats.argtypes = [vp, C.POINTER(GLLetStats)]
    lib.gl_comm_unique_id.argtypes = [vp]
    lib.gl_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.gl_comm_local_group_create.argtypes = [C.c_int]
    lib.gl_comm_local_group_create.restype = vp
    lib.gl_comm_local_group_destroy.argtypes = [vp]
    lib.gl_comm_local_group_destroy.restype = None
    lib.gl_comm_init_local.argtypes = [vp, C.c_int, vp]
    lib.gl_comm_finalize.argtypes = [vp]
    lib.gl_comm_set_options.argtypes = [vp, C.c_int, C.c_double]
    lib.gl_reset_launch_count.argtypes = [vp]
    lib.gl_domain_split_collective.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.POINTER(C.c_longlong), vp]
    lib.gl_set_subgrid_params.argtypes = [vp, C.POINTER(GLSubgridParams)]
    lib.gl_blackhole_accretion.argtypes = [vp, vp, C.c_longlong, C.POINTER(GLSinkState), C.c_int, C.POINTER(C.c_longlong)]
    lib.gl_dust_density.argtypes = [vp, vp, C.c_longlong, vp, C.c_int]
    lib.gl_dust_drag.argtypes = [vp, vp, C.c_longlong, C.POINTER(GLDustState), C.c_int]
    lib.gl_reset_launch_count.restype = None
    lib.gl_gadget_bind.argtypes = [vp, C.POINTER(GLGadgetLayout)]
    lib.gl_gadget_pull.argtypes = [vp, vp, vp, C.c_longlong, C.c_longlong]
    lib.gl_gadget_push.argtypes = [vp, vp, vp, C.c_int]
    lib.gl_gadget_compute_accelerations.argtypes = [vp, vp, vp, C.c_longlong, C.c_longlong, C.c_int]
    lib.gl_gadget_pack_segment.argtypes = [vp, vp, vp, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, vp, vp]
    lib.gl_gadget_push_segment_async.argtypes = [vp, vp, vp, C.c_int]
    lib.gl_gadget_push_wait.argtypes = [vp]
    lib.gl_gadget_reorder.argtypes = [C.POINTER(GLGadgetLayout), vp, vp, C.c_longlong, C.c_longlong, vp]
    lib.gl_gadget_drift.argtypes = [C.POINTER(GLGadgetLayout), vp, vp, C.c_longlong, C.c_longlong, C.c_int, C.c_double, C.c_double,
                                    C.POINTER(C.c_longlong)]
    lib.gl_snapshot_write.argtypes = [C.c_char_p, C.POINTER(GLGadgetLayout), C.POINTER(GLIOLayout), C.POINTER(GLIOParams), vp, vp,
                                      C.c_longlong, vp, vp, C.c_char_p, C.c_int]
    lib.gl_domain_top_tree.argtypes = [vp, vp, C.c_longlong, C.c_int, C.c_int, C.c_double, C.POINTER(GLTopNode), C.c_int,
                                       C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.gl_gadget_drift_list.argtypes = [C.POINTER(GLGadgetLayout), vp, vp, vp, C.c_longlong, C.c_longlong, C.c_int, C.c_double,
                                         C.c_double, C.POINTER(C.c_longlong)]
    lib.gl_gadget_domain_keys.argtypes = [C.POINTER(GLGadgetLayout), vp, C.c_longlong, vp, vp, vp, vp, vp]
    lib.gl_gadget_costfactors.argtypes = [C.POINTER(GLGadgetLayout), vp, C.c_longlong, vp]
    lib.gl_domain_split.argtypes = [C.POINTER(GLTopNode), C.c_int, vp, vp, C.c_longlong, C.c_int, vp, vp, vp, vp]
    lib.gl_gadget_peano_hilbert_order.argtypes = [C.POINTER(GLGadgetLayout), vp, vp, C.c_longlong, C.c_longlong, vp, vp]
    lib.gl_gadget_advance_and_find_timesteps.argtypes = [C.POINTER(GLGadgetLayout), C.POINTER(GLTimestepParams), vp, vp, vp,
                                                         C.c_longlong, vp, C.c_char_p, C.c_int]
    lib.gl_snapshot_read_header.argtypes = [C.c_char_p, C.c_int, C.POINTER(GLIOHeader), C.c_char_p, C.c_int]
    lib.gl_snapshot_read.argtypes = [C.c_char_p, C.POINTER(GLGadgetLayout), C.POINTER(GLIOLayout), C.POINTER(GLIOParams), vp, vp,
                                     C.c_longlong, C.c_longlong, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
                                     C.POINTER(GLIOHeader), C.c_char_p, C.c_int]
    lib.gl_export_tree.argtypes = [vp, C.POINTER(GLTreeLayout), vp, vp, vp, vp, C.c_longlong, C.c_longlong, C.c_int,
                                   C.POINTER(C.c_longlong)]
    lib.gl_set_velpred.argtypes = [vp, vp, C.c_longlong, C.c_int]
    lib.gl_set_accelerations.argtypes = [vp, vp, vp, vp, C.c_int]
    lib.gl_run_begin.argtypes = [vp, C.c_int]
    lib.gl_run_kick.argtypes = [vp, C.POINTER(GLTimestepParams), C.c_int]
    lib.gl_run_steps.argtypes = [vp, C.POINTER(GLTimestepParams), C.c_int, C.POINTER(GLRunInfo)]
    lib.gl_get_run_state.argtypes = [vp] + [vp] * 8 + [C.c_int]
    lib.gl_gadget_run_steps.argtypes = [vp, vp, vp, C.c_longlong, C.c_longlong, C.POINTER(GLTimestepParams), C.c_int, C.POINTER(GLRunInfo)]
    lib.gl_next_sync_point.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    lib.gl_run_set_ids.argtypes = [vp, vp, C.c_int]
    lib.gl_get_run_ids.argtypes = [vp, vp, C.c_int]
    lib.gl_domain_exchange.argtypes = [vp, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    _lib = lib
    return lib


def make_params(p):
    """dict (ics.default_params) -> GLParams, applying set_softenings (gravtree.c:835-885)."""
    g = GLParams()
    g.ErrTolTheta = p["ErrTolTheta"]
    g.ErrTolForceAcc = p["ErrTolForceAcc"]
    g.TypeOfOpeningCriterion = int(p["TypeOfOpeningCriterion"])
    g.DesNumNgb = int(p["DesNumNgb"])
    g.MaxNumNgbDeviation = p["MaxNumNgbDeviation"]
    g.BlackHoleNgbFactor = p["BlackHoleNgbFactor"]
    g.ArtBulkViscConst = p["ArtBulkViscConst"]
    g.G = p["G"]
    for i in range(6):
        g.ForceSoftening[i] = 2.8 * p["Softening"][i]        # gravtree.c:881-882
        g.SofteningTable[i] = p["Softening"][i]
    g.MinGasHsml = p["MinGasHsmlFractional"] * g.ForceSoftening[0]  # gravtree.c:884
    g.Timebase_interval = p["Timebase_interval"]
    g.Ti_Current = int(p["Ti_Current"])
    g.ComovingIntegrationOn = 0
    g.OmegaLambdaHubble2 = 0.0
    return g


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class GLError(RuntimeError):
    pass
