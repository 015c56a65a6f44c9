/* gl_b200.h - C ABI of libgl_b200.so: the B200-native (sm_100a) force path of the
 * Leicester GADGET-3 fork.  Plain pointers and sizes only; no torch types.
 *
 * Each entry point names the reference interface it replaces (file:line into the
 * reference tree).  The reference has no plugin API: its boundary is a set of global C
 * symbols operating on the global arrays P[] / SphP[] (proto.h:281,295,331,332;
 * forcetree.h:48,67,86).  The functions below are what those symbols bind to; the glue
 * a maintainer adds on the reference side (which owns allvars.h and therefore the struct
 * layout) is shown in INTEGRATION.md and shipped as integration/gl_gadget_glue.c.
 *
 * Conventions: every function returns 0 on success, non-zero on error
 * (gl_last_error() gives the text).  The reference's own convention for unrecoverable
 * errors is endrun(code) (endrun.c:23); the glue maps non-zero returns onto it.
 * One gl_ctx <-> one GPU <-> one reference MPI rank.  Calls are synchronous: results
 * are complete in the caller's memory on return.
 */
#ifndef GL_B200_H
#define GL_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gl_ctx gl_ctx;
typedef unsigned long long gl_peanokey;	/* peanokey, allvars.h:144 (63 bits used) */

/* Run-time parameters the path reads out of the reference's global `All`
 * (SURVEY.md 8(b) "Inputs read").  Field names follow struct global_data_all_processes. */
typedef struct gl_params
{
  double ErrTolTheta;		/* All.ErrTolTheta, forcetree.c:2074-2076 (0 selects the relative criterion) */
  double ErrTolForceAcc;	/* All.ErrTolForceAcc, forcetree.c:1850 */
  int TypeOfOpeningCriterion;	/* gravtree.c:396-397 */
  int DesNumNgb;		/* density.c:158 */
  double MaxNumNgbDeviation;	/* density.c:559 */
  double BlackHoleNgbFactor;	/* density.c:551 */
  double ArtBulkViscConst;	/* hydra.c:1538 */
  double G;			/* gravtree.c:403 */
  double ForceSoftening[6];	/* All.ForceSoftening[] = 2.8 * SofteningTable[], gravtree.c:881-882 */
  double MinGasHsml;		/* gravtree.c:884 */
  double Timebase_interval;	/* density.c:491, hydra.c:1583 */
  int Ti_Current;		/* All.Ti_Current */
  int ComovingIntegrationOn;	/* must be 0: the comoving factors of hydra.c:196-206 are not implemented */
  double OmegaLambdaHubble2;	/* All.OmegaLambda*All.Hubble^2, gravtree.c:405-411 (0 for the shipped run) */
  double SofteningTable[6];	/* All.SofteningTable[], only for the self-potential term of potential.c:176 */
} gl_params;

/* ---- life cycle ---------------------------------------------------------------- */
/* max_part plays the role of All.MaxPart (allocate.c:56): device buffers are sized once. */
int gl_create(gl_ctx ** ctx, int device, long long max_part);
void gl_destroy(gl_ctx * ctx);
const char *gl_last_error(const gl_ctx * ctx);
int gl_set_params(gl_ctx * ctx, const gl_params * p);
/* stream the library launches on (a cudaStream_t); lets a caller time it with its own events */
void *gl_stream(gl_ctx * ctx);

/* ---- particle state, SoA form ---------------------------------------------------- */
/* Loads the fields of P[]/SphP[] the path reads (allvars.h:1131-1377, :1384-1639).  All arrays
 * have n elements and are indexed like P[]; the SphP fields are only read where type==0.
 * `where`: 0 = host pointers, 1 = device pointers (same device as the ctx).
 *   pos[3n] P.Pos, vel[3n] SphP.VelPred for gas else P.Vel, mass P.Mass, type P.Type,
 *   hsml P.Hsml, oldacc P.OldAcc, timebin P.TimeBin, ti_begstep P.Ti_begstep,
 *   entropy SphP.Entropy, dtentropy SphP.e.DtEntropy, divvel SphP.v.DivVel (may be NULL = 0). */
int gl_load_particles(gl_ctx * ctx, long long n, const double *pos, const double *vel, const double *mass,
		      const int *type, const double *hsml, const double *oldacc, const int *timebin,
		      const int *ti_begstep, const double *entropy, const double *dtentropy, int where);
/* update only P.OldAcc / P.Hsml between calls (used when chaining calls like accel.c:61-67 does) */
int gl_set_oldacc(gl_ctx * ctx, const double *oldacc, int where);
int gl_set_hsml(gl_ctx * ctx, const double *hsml, int where);

/* SphP state of particles that are NOT density targets of this rank/step (inactive particles, or particles
 * owned by another GPU after the allgather that replaces the hydro export loop, hydra.c:412): SphP.d.Density,
 * .Pressure, .h.DhsmlDensityFactor, .v.DivVel, .r.CurlVel; n elements each, indexed like P[]. */
int gl_set_sph_state(gl_ctx * ctx, const double *density, const double *pressure, const double *dhsml,
		     const double *divvel, const double *curlvel, int where);

/* ---- multi-GPU: the same two loads straight out of the NCCL all-gather receive buffers ---------------------
 * These replace the import side of the reference's export/import loops (gravtree.c:246-366, density.c:258-342,
 * hydra.c:412-520).  Every rank sends one block of doubles and one of ints for its Peano-Hilbert segment, padded to
 * m particles; rank r's block sits at f64buf + r*11*m = [pos 3m | mass | hsml | oldacc | vel 3m | entropy |
 * dtentropy] and i32buf + r*3*m = [type | timebin | ti_begstep].  full = 0: a gravity-only exchange sent just the
 * prefix the tree build and the walk read - blocks of 6*m doubles [pos | mass | hsml | oldacc] and m ints [type].
 * seg_start[world+1]: global index of each rank's first particle (seg_start[world] = n).  Device pointers. */
int gl_load_particles_gathered(gl_ctx * ctx, long long n, int world, long long m, const long long *seg_start,
			       const double *f64buf, const int *i32buf, int full);
/* split exchange: after a full = 0 load the tree build and the walk can start; the SPH-side inputs follow while they run
 * and are handed over here before gl_density - rank r's blocks at f64rest + r*5*m = [vel 3m | entropy | dtentropy] and
 * i32rest + r*2*m = [timebin | ti_begstep].  (Node vmax, which only gl_get_nodes reports, is not refreshed.) */
int gl_set_sph_inputs_gathered(gl_ctx * ctx, int world, long long m, const long long *seg_start, const double *f64rest,
			       const int *i32rest);
/* after density (replaces hydra.c:412 and the hmax exchange forcetree.c:1722-1747): rank r's block at
 * buf + r*6*m = [hsml | density | pressure | dhsml | divvel | curlvel] */
int gl_set_sph_state_gathered(gl_ctx * ctx, int world, long long m, const long long *seg_start, const double *buf);

/* ---- domain / keys: domain.c:1972 domain_findExtent, peano.c:300,320,335 ------------------ */
/* computes DomainCorner/Center/Len/Fac from the loaded positions; out[8] = len, fac, corner[3], center[3] */
int gl_domain_find_extent(gl_ctx * ctx, double *out8);
/* multi-GPU: impose the global extent (the ncclAllReduce min/max result, domain.c:1996-1997) */
int gl_domain_set_extent(gl_ctx * ctx, const double xmin[3], const double xmax[3]);
/* peano_and_morton_key for every loaded particle on the BITS_PER_DIMENSION=21 grid */
int gl_peano_keys(gl_ctx * ctx, gl_peanokey * ph_keys, gl_peanokey * morton_keys, int where);
/* peano_hilbert_order (peano.c:25): order[] such that new P[i] = old P[order[i]]; the gas block
 * [0,n_gas) and the rest are ordered separately as the reference does (peano.c:37-88). */
int gl_peano_hilbert_order(gl_ctx * ctx, long long n_gas, int *order, int where);
/* scalar (host) versions of peano.c:300 / :320, exported for the key parity tests */
gl_peanokey gl_peano_hilbert_key(int x, int y, int z, int bits);
gl_peanokey gl_morton_key(int x, int y, int z, int bits);

/* ---- tree: forcetree.c:67 force_treebuild (+ :468 force_update_node_recursive) ------------- */
/* returns 0 and stores the node count (Numnodestree analogue) in *numnodes */
int gl_force_treebuild(gl_ctx * ctx, long long *numnodes);
/* node table for parity tests: per node 12 doubles
 * {len, center[3], mass, s[3], maxsoft, hmax, vmax, level}; host buffer of 12*numnodes doubles */
int gl_get_nodes(gl_ctx * ctx, double *table, long long cap_nodes);

/* ---- targets ---------------------------------------------------------------------------------- */
/* Restrict the next force calls to a subset of P[] indices (the active list
 * FirstActiveParticle/NextActiveParticle, or a GPU's Peano-Hilbert segment).  n<0 = all. */
int gl_set_targets(gl_ctx * ctx, const int *index, long long n, int where);
/* Same, for a list that is already ordered along the space-filling curve (a GPU's Peano-Hilbert segment after
 * domain.c:1075, or the reference's P[] after peano_hilbert_order, peano.c:25): warps take 32 consecutive list
 * entries as they come and the library skips its own Peano-Hilbert sort.  Results do not depend on the order. */
int gl_set_targets_ordered(gl_ctx * ctx, const int *index, long long n, int where);

/* ---- the four force entry points -------------------------------------------------------------- */
int gl_gravity_tree(gl_ctx * ctx);	/* gravtree.c:27  (includes the tree build when flagged) */
int gl_density(gl_ctx * ctx);		/* density.c:89   (with the Hsml iteration) */
int gl_force_update_hmax(gl_ctx * ctx);	/* forcetree.c:1661 */
int gl_hydro_force(gl_ctx * ctx);	/* hydra.c:145 */
int gl_compute_accelerations(gl_ctx * ctx, int mode);	/* accel.c:27, force part (:61-106) */
int gl_compute_potential(gl_ctx * ctx);	/* potential.c:22 (all particles) */
void gl_set_tree_reconstruct_flag(gl_ctx * ctx, int flag);	/* TreeReconstructFlag, domain.c:135 */

/* ---- results (arrays indexed like P[], n elements; NULL pointers are skipped) ------------------- */
/* gravity: P.g.GravAccel (already x G), P.OldAcc, P.GravCost (gravtree.c:392-403, forcetree.c:2282) */
int gl_get_gravity(gl_ctx * ctx, double *gravaccel3, double *oldacc, float *gravcost, int where);
/* P.p.Potential (already x G, self term removed as potential.c:176 does) */
int gl_get_potential(gl_ctx * ctx, double *potential, int where);
/* density: P.Hsml, P.n.NumNgb, SphP.d.Density, .h.DhsmlDensityFactor, .v.DivVel, .r.CurlVel, .Pressure;
 * sink/dust targets: b1/d1 density, b2/d2 entropy, b3/d3 gas velocity (density.c:944-988, :520-546) */
int gl_get_density(gl_ctx * ctx, double *hsml, double *numngb, double *density, double *dhsml, double *divvel,
		   double *curlvel, double *pressure, double *sink_density, double *sink_entropy,
		   double *sink_gasvel3, int where);
/* hydro: SphP.a.HydroAccel, .e.DtEntropy (converted, hydra.c:583), .MaxSignalVel */
int gl_get_hydro(gl_ctx * ctx, double *hydroaccel3, double *dtentropy, double *maxsignalvel, int where);
/* restrict the gl_get_* calls (device or host destination) to the particles [first, first+count): a GPU that only
 * evaluated its own segment copies only that; count < 0 = all again.  Values land at the same indices of the
 * destination arrays. */
int gl_set_result_range(gl_ctx * ctx, long long first, long long count);
/* integer pair counts (not stored by the reference, density.c:858 is a dead local): number of j with
 * Mass_j != 0 and r2 < h_i^2 (density) / (r2 < h_i^2 or r2 < h_j^2) and r > 0 (hydro) */
int gl_get_pair_counts(gl_ctx * ctx, int *dens_pairs, int *hydro_pairs, int where);

/* ---- statistics --------------------------------------------------------------------------------- */
typedef struct gl_stats
{
  double ms_keys_sort, ms_treebuild, ms_moments, ms_gravity_walk, ms_density, ms_hmax, ms_hydro, ms_h2d, ms_d2h, ms_potential;
  long long interactions;	/* sum of P.GravCost over targets of the last gl_gravity_tree */
  long long density_targets, density_passes_max, hydro_targets;
  long long numnodes, numentries;
  long long density_fat_targets;	/* targets handed to the block-per-target density path */
  long long kernel_launches;	/* kernels launched by this library since gl_create / last reset */
  long long density_pairs;	/* sum over targets of the integer neighbour count (r2 < h_i^2) of their final pass */
  long long density_target_passes;	/* number of (target, pass) evaluations of the last gl_density (>= density_targets) */
  long long hydro_pairs;	/* sum over targets of the pair count (r2 < h_i^2 or r2 < h_j^2, r > 0) */
  long long density_slowest_index, density_slowest_passes;	/* P[] index of the target with the most Hsml passes in the last gl_density (-1: all took one) */
} gl_stats;
int gl_get_stats(gl_ctx * ctx, gl_stats * out);
void gl_reset_launch_count(gl_ctx * ctx);

/* ---- multi-rank mode: one rank <-> one GPU, NCCL inside the library -----------------------------------------------
 * Replaces the reference's MPI choreography of the force path: the pseudo-particle exchange of the tree build
 * (force_exchange_pseudodata forcetree.c:879-1016, force_treeupdate_pseudos :1025, force_insert_pseudo_particles :439)
 * and the export/import loops of gravity_tree() (gravtree.c:175-347), density() (density.c:193-342), hydro_force()
 * (hydra.c:274-520) and force_update_hmax() (forcetree.c:1722-1747).  With a communicator attached
 *   - gl_load_particles() loads the rank's OWN particles (any decomposition; contiguous Peano-Hilbert key ranges
 *     as domain.c:1075 produces them prune best), in the caller's order (P[] is Peano-Hilbert ordered, peano.c:25);
 *   - the tree build inside gl_gravity_tree / gl_force_treebuild / gl_density all-reduces the domain extent
 *     (domain.c:1996), exchanges the locally-essential tree (csrc/gl_let.cu) and builds the merged tree; every rank
 *     must make the same sequence of calls (they are collective, like the reference's);
 *   - gl_force_update_hmax() first exchanges Hsml, the SphP fields hydro reads and the cells' hmax;
 *   - results (gl_get_*) cover the own particles [0, n).  Interaction counts, neighbour sets and sums are those of a
 *     single-rank run on the union of all ranks' particles.
 * The 128-byte id comes from gl_comm_unique_id() on one rank and is broadcast by the CALLER (the reference's
 * MPI_Bcast; torch.distributed in bench.py): the library opens no rendezvous of its own. */
int gl_comm_unique_id(void *id128);
int gl_comm_init(gl_ctx * ctx, int rank, int world, const void *id128);
/* several contexts of ONE process as ranks (one host thread each): how the multi-rank path is tested on a single GPU */
void *gl_comm_local_group_create(int world);
void gl_comm_local_group_destroy(void *group);
int gl_comm_init_local(gl_ctx * ctx, int rank, void *group);
int gl_comm_finalize(gl_ctx * ctx);
/* want_gravity = 0: the export test leaves out the gravity opening criteria (a run that only calls density / hydro);
 * search_padding (>= 1, default 1.3): factor on the search radii in the export test (room for Hsml to grow) */
int gl_comm_set_options(gl_ctx * ctx, int want_gravity, double search_padding);
/* Collective partition (domain.c:1075 domain_findSplit_work_balanced with one stretch per rank): the global
 * Peano-Hilbert order of ALL ranks' particles is cut into `world` stretches of equal cost, and every rank learns where each
 * of its loaded particles belongs: dest[n] (where = 0 host / 1 device), to_go[world] (host; 0 for the rank itself) - the
 * toGo row of domain_countToGo (domain.c:1274).  cost[n]: the reference's (1 + GravCost) / 2^TimeBin (domain.c:378), NULL =
 * equal costs.  splitters[world-1] (may be NULL) <- the first Peano-Hilbert key of ranks 1 .. world-1, the same on every
 * rank.  No rank sees all keys: each describes its cost distribution by 2048 sample keys, which are all-gathered (balance
 * within world / 2048 of a stretch).  The struct migration (domain_exchange, domain.c:665) stays with the caller.  Every
 * rank must call it; the tree is invalidated. */
int gl_domain_split_collective(gl_ctx * ctx, const double *cost, int cost_where, int *dest, int dest_where, long long *to_go,
			       unsigned long long *splitters);
typedef struct gl_let_stats
{
  long long own_particles, exported_cells, exported_particles, imported_cells, imported_particles;
  long long builds, violations;	/* multi-rank builds since gl_create; phases repeated with a full export */
  int full;			/* the last build exported every particle */
  double ms_own_tree, ms_select, ms_exchange, ms_sph_exchange;
} gl_let_stats;
int gl_get_let_stats(gl_ctx * ctx, gl_let_stats * out);

/* ---- drop-in layer for the reference's AoS arrays ------------------------------------------------- */
/* Byte offsets inside struct particle_data / struct sph_particle_data; the glue fills them with
 * offsetof() so the library never needs allvars.h (SURVEY.md 8(b) "ABI caveat"). */
typedef struct gl_gadget_layout
{
  long long p_stride, sph_stride;
  int p_Pos, p_Vel, p_Mass, p_GravAccel, p_OldAcc, p_Hsml, p_NumNgb, p_GravCost, p_Ti_begstep, p_Type, p_TimeBin;
  int p_SinkDensity, p_SinkEntropy, p_SinkGasVel;	/* b1/b2/b3 (Type 5) */
  int p_DustDensity, p_DustEntropy, p_DustGasVel;	/* d1/d2/d3 (Type 2) */
  int p_SwallowID;
  int s_Entropy, s_Pressure, s_VelPred, s_MaxSignalVel, s_Density, s_DtEntropy, s_HydroAccel,
    s_DhsmlDensityFactor, s_DivVel, s_CurlVel;
  int p_Ti_current;		/* only gl_gadget_drift reads these two */
  int s_DragAccel;		/* SphP.da.DragAccel (-DDUST), < 0: the build has none */
} gl_gadget_layout;
int gl_gadget_bind(gl_ctx * ctx, const gl_gadget_layout * lay);
/* pack P[0..NumPart) / SphP[0..N_gas) into the device SoA (host threads + pinned staging) */
int gl_gadget_pull(gl_ctx * ctx, const void *P, const void *SphP, long long NumPart, long long N_gas);
/* compute_accelerations(mode) of accel.c:27 on the reference's own arrays in one call: pull, the force path,
 * and a software-pipelined push (results of one phase cross PCIe and are scattered while the next phase runs) */
int gl_gadget_compute_accelerations(gl_ctx * ctx, void *P, void *SphP, long long NumPart, long long N_gas, int mode);
/* The lazy drift of the sources (predict.c:129 drift_particle, called per particle from inside the reference's walks:
 * forcetree.c:1911, ngb.c:57,194): every particle with Ti_current != time1 is moved to time1 IN PLACE, with the
 * reference's statements in the reference's order (Pos; for gas VelPred from GravAccel / HydroAccel / DragAccel,
 * Density and Hsml from DivVel, Pressure from Entropy + DtEntropy * dt_entr; Type 3 is skipped, -DVIRTUAL), on all host
 * cores.  Host code only - no context, no GPU; non-comoving integration.  Returns non-zero if a particle is ahead of
 * time1 (endrun(12) in the reference).  *ndrifted <- number of particles moved. */
int gl_gadget_drift(const gl_gadget_layout * lay, void *P, void *SphP, long long NumPart, long long N_gas, int time1,
		    double Timebase_interval, double MinGasHsml, long long *ndrifted);
/* The same for the particles of a list only - the drift of the ACTIVE particles at a synchronisation point
 * (find_next_sync_point_and_drift, run.c:325-329); everybody else is left for the lazy drift. */
int gl_gadget_drift_list(const gl_gadget_layout * lay, void *P, void *SphP, const int *list, long long nlist, long long N_gas,
			 int time1, double Timebase_interval, double MinGasHsml, long long *ndrifted);
/* The second half of peano_hilbert_order() (peano.c:25-92 -> reorder_gas :106, reorder_particles :148): permute the
 * structs so that new P[k] = old P[order[k]] (order from gl_peano_hilbert_order: gas block and the rest ordered
 * separately), SphP[] alongside for the gas block.  The reference follows permutation cycles in place, one struct at a
 * time; here all host cores gather from a temporary copy.  Host code only - no context, no GPU. */
int gl_gadget_reorder(const gl_gadget_layout * lay, void *P, void *SphP, long long NumPart, long long N_gas, const int *order);
/* peano_hilbert_order() (peano.c:25-92) as a whole, on the host, from the keys the domain decomposition has already
 * computed (the reference's global Key[], domain.c:1750): stable sort of the gas block and of the rest by key (the
 * reference's merge sort is stable, peano.c:551) on all host cores, then gl_gadget_reorder.  order_out (may be NULL)
 * receives the permutation (new P[k] = old P[order[k]]).  Host code only - no context, no GPU.  (gl_peano_hilbert_order
 * above is the GPU form for particles that are already loaded.) */
int gl_gadget_peano_hilbert_order(const gl_gadget_layout * lay, void *P, void *SphP, long long NumPart, long long N_gas,
				  const gl_peanokey * Key, int *order_out);
/* The top-level tree of the domain decomposition for one rank: domain_determineTopTree
 * (domain.c:1738-1916) with domain_check_for_local_refine (domain.c:1524) and the leaf numbering of domain_walktoptree
 * (domain.c:1495), from the particles' Peano-Hilbert keys and cost factors (1 + GravCost) / 2^TimeBin (domain.c:378), both
 * in P[] order.  The node numbering, every Count and every Cost (summation order included) are the reference's; the key
 * sort runs on all host cores and the particle ranges are found by bisection.  nodes[] = struct local_topnode_data
 * (domain.c:48); returns 1 when max_nodes is too small (the reference then raises TopNodeAllocFactor and repeats,
 * domain.c:246-269).  Host code only - no context, no GPU. */
typedef struct gl_topnode
{
  gl_peanokey Size, StartKey;
  long long Count;
  double Cost;
  int Daughter, Leaf, Parent, PIndex;
} gl_topnode;
int gl_domain_top_tree(const gl_peanokey * Key, const double *costfactor, long long NumPart, int NTask, int multipledomains,
		       double topnodefactor, gl_topnode * nodes, int max_nodes, int *ntopnodes, int *ntopleaves);
/* The other host pieces of a one-rank domain_Decomposition() (domain.c:100): the cube and the keys (domain_findExtent
 * domain.c:1972 + the key loop domain.c:1748), the cost factors (domain.c:378), and the leaf work / leaf count tables
 * with the split into ncpu stretches of leaves (domain_sumCost domain.c:1920, domain_findSplit_work_balanced :1075). */
int gl_gadget_domain_keys(const gl_gadget_layout * lay, const void *P, long long NumPart, double DomainCorner[3],
			  double DomainCenter[3], double *DomainLen, double *DomainFac, gl_peanokey * Key);
int gl_gadget_costfactors(const gl_gadget_layout * lay, const void *P, long long NumPart, double *costfactor);
int gl_domain_split(const gl_topnode * nodes, int ntopleaves, const gl_peanokey * Key, const double *costfactor, long long NumPart,
		    int ncpu, int *start_list, int *end_list, float *leaf_work, int *leaf_count);
/* The multi-rank half of domain_decompose() (domain.c:470-487).  In the reference these functions run
 * identically on every rank, on leaf tables that domain_sumCost has summed over the ranks (the caller's MPI_Allreduce of
 * the leaf_work / leaf_count tables gl_domain_split returns, domain.c:1950-1960), so they are plain functions of those
 * tables and of NTask: the split into multipledomains * NTask stretches of about equal work (domain_findSplit_work_balanced
 * :1075), dealt to the ranks so that the particle numbers are even (domain_assign_load_or_work_balanced(1) :1153); if a rank
 * would then hold more than max_load particles or max_load_sph gas particles (domain_check_memory_bound :595) the split by
 * particle number (:1225) dealt by work (mode 0) instead.  Out: domain_task[ntopleaves] = DomainTask[], start_list /
 * end_list[multipledomains * NTask] = DomainStartList / DomainEndList (ordered by rank), list_load / list_loadsph /
 * list_work[NTask] (may be NULL).  Returns 0 (first variant), 1 (second), 2 (neither fits: endrun(0), domain.c:483-487),
 * 3 (bad arguments: multipledomains must be a power of two, ntopleaves >= multipledomains * NTask).  Host code only. */
int gl_domain_assign(int NTask, int multipledomains, int ntopleaves, const float *leaf_work, const int *leaf_count,
		     const int *leaf_count_sph, long long max_load, long long max_load_sph, int *domain_task, int *start_list,
		     int *end_list, int *list_load, int *list_loadsph, double *list_work);
/* domain_countToGo (domain.c:1274-1305, in front of its memory-limited iteration): dest[NumPart] (may be NULL) = the new
 * owner of every local particle, to_go / to_go_sph[NTask] = how many (gas) particles leave for each rank.  The struct
 * exchange itself (domain_exchange :665) stays with the caller's MPI. */
int gl_domain_count_to_go(const gl_topnode * nodes, const gl_peanokey * Key, const int *type, long long NumPart, const int *domain_task,
			  int ThisTask, int NTask, int *to_go, int *to_go_sph, int *dest);
/* Individual time steps: the per-particle part of advance_and_find_timesteps() (timestep.c:29) for
 * the particles of the active list - get_timestep (timestep.c:607: acceleration, Courant, dust, sink-accretion and
 * feedback criteria of the shipped -D list), the power-of-two bin with the "only onto a synchronised bin" rule
 * (timestep.c:136-154), P.TimeBin / P.Ti_begstep, and do_the_kick (timestep.c:364: Vel, VelPred, DragAccel = 0, Entropy
 * and DtEntropy with their limiters) - in place, on all host cores, with the reference's statements in the reference's
 * order.  Host code only - no context, no GPU; non-comoving integration, no dynamic tree update (the tree is rebuilt).
 * What stays with the caller is the serial bookkeeping of the time-bin lists (TimeBinCount[], First/Last/Next/
 * PrevInTimeBin[], timestep.c:170-233): binold[k] <- the bin particle active[k] had, so the caller relinks exactly the
 * particles whose P.TimeBin differs from it, in active-list order. */
typedef struct gl_timestep_params
{
  int Ti_Current;		/* All.Ti_Current */
  int TimeBinActive[32];	/* TimeBinActive[TIMEBINS] */
  double Timebase_interval, ErrTolIntAccuracy, CourantFac, MaxSizeTimestep, MinSizeTimestep, MinEgySpec;
  double GammaMinus1;		/* GAMMA_MINUS1 (allvars.h:64-65) */
  double SofteningTable[6];	/* All.SofteningTable[] */
  /* fork-specific criteria: Type 3 (virtual particles) and Type 5 (sinks), timestep.c:889-1000 */
  double OuterBoundary, UnitVelocity_in_cm_per_s, FeedBackVelocity, SMBHmass, InnerBoundary, SinkBoundary, AccDtBlackHole;
  int p_DeltaDustMomentum;	/* P.DeltaDustMomentum (-DDUST), < 0: the build has none */
  int p_NewDensity;		/* P.NewDensity (-DVIRTUAL), < 0: the build has none */
} gl_timestep_params;
/* Non-zero return where the reference calls endrun(): a step below MinSizeTimestep (888), of integer size 0 or 1 (818,
 * 112313), beyond the end of the time line (888); also for a Type-3 particle with NewDensity > 0, whose criterion
 * reads an uninitialised variable in the reference (`rho`, timestep.c:894).  Particles are independent: on error the ones
 * already processed keep their kick, like in the reference, which aborts. */
int gl_gadget_advance_and_find_timesteps(const gl_gadget_layout * lay, const gl_timestep_params * par, void *P, void *SphP,
					 const int *active, long long nactive, int *binold, char *err, int errlen);
/* scatter results back in place; what = bit mask: 1 gravity, 2 density, 4 hydro */
int gl_gadget_push(gl_ctx * ctx, void *P, void *SphP, int what);

/* ---- run()'s loop with the state resident on the device (device side) -------------------------
 * run.c:63-110 alternates find_next_sync_point_and_drift (run.c:190), domain_Decomposition, compute_accelerations
 * (accel.c:27) and advance_and_find_timesteps (timestep.c:29) on P[] / SphP[] in host memory.  These entry points run
 * the same loop on the SoA state the context already holds: drift_particle (predict.c:129) for every particle to the
 * next synchronisation point, the force path for the particles on active time bins, get_timestep (timestep.c:607) +
 * the bin rule (timestep.c:136-154) + do_the_kick (timestep.c:364) for them - nothing crosses PCIe but the 32
 * TimeBinCount[] values per step.  Same restrictions as the host-side pieces above (non-comoving, tree rebuilt every
 * step = TreeDomainUpdateFrequency 0, no sub-grid loops inside the loop: DragAccel and DeltaDustMomentum are zero).
 * Multi-rank mode (a communicator attached): every rank integrates its own particles, gl_run_steps is collective -
 * the ranks' TimeBinCount[] are all-gathered (the MPI_Allreduce of run.c:218 / :290), the force path imports its
 * locally-essential set in every step, an endrun() on one rank stops all.  The ranks keep the particles they have
 * (any assignment gives the same results; re-balancing is gl_domain_split_collective + the caller's exchange). */
/* SphP[].VelPred of the leading n_gas particles (SphP[i] belongs to P[i]); gas particles without a row start from the
 * loaded velocities in gl_run_begin */
int gl_set_velpred(gl_ctx * ctx, const double *velpred3, long long n_gas, int where);
/* P[].g.GravAccel, SphP[].a.HydroAccel (3n each), SphP[].MaxSignalVel (n), indexed like P[] (non-gas rows of the
 * last two are ignored); NULL = keep what the last force computation in this context left.  */
int gl_set_accelerations(gl_ctx * ctx, const double *gravaccel3, const double *hydroaccel3, const double *maxsignalvel, int where);
/* checks that the context holds the whole state (particles, SphP state, accelerations); every particle is taken to
 * be drifted to Ti_Current (P[].Ti_current == All.Ti_Current) */
int gl_run_begin(gl_ctx * ctx, int Ti_Current);
/* advance_and_find_timesteps() for the particles on the bins par->TimeBinActive[] marks, at par->Ti_Current.
 * commit_dtentropy != 0: SphP.e.DtEntropy of those gas particles is first taken from the last hydro_force() of this
 * context (hydra.c:785-813 writes it for the active particles). Errors as gl_gadget_advance_and_find_timesteps. */
int gl_run_kick(gl_ctx * ctx, const gl_timestep_params * par, int commit_dtentropy);
typedef struct gl_run_info
{
  int steps_done, Ti_Current;	/* iterations executed (fewer than asked for only at the end of the time line) */
  long long active_total, active_last;	/* force updates summed over the steps (NumForceUpdate), and of the last step */
  long long interactions;	/* gravity interactions summed over the steps */
  double ms_total;		/* device time of the whole call */
} gl_run_info;
/* The time-line arithmetic of find_next_sync_point_and_drift (run.c:199-216, :276-290) on the ranks' TimeBinCount[] rows -
 * hist[32 r + b] = particles of rank r on time bin b (b < 29), hist[32 r + 31] = particles rank r holds.  Returns the next
 * kick time (the reference's MPI_Allreduce(MPI_MIN) over the ranks, run.c:218), fills TimeBinActive[32] for it and
 * counts the force updates of `rank` and of all ranks.  Host code only - gl_run_steps calls it with the all-gathered rows. */
int gl_next_sync_point(const int *hist, int world, int rank, int Ti_Current, int *TimeBinActive, long long *nactive_own,
		       long long *nactive_all, long long *particles_all);
/* nsteps iterations of the loop; par->Ti_Current and par->TimeBinActive[] are advanced in place (the caller's
 * All.Ti_Current / TimeBinActive[]) */
int gl_run_steps(gl_ctx * ctx, gl_timestep_params * par, int nsteps, gl_run_info * info);
/* the state the loop integrates, indexed like P[] (velpred3: a row per particle, meaningful where the type is 0); NULL
 * = not wanted.  Forces,
 * densities and the other results come from gl_get_gravity / gl_get_density / gl_get_hydro as usual */
int gl_get_run_state(gl_ctx * ctx, double *pos3, double *vel3, double *velpred3, double *hsml, double *entropy, double *dtentropy,
		     int *timebin, int *ti_begstep, int where);
/* domain_exchange (domain.c:665) for the resident state, collective: the new owners come from
 * gl_domain_split_collective (by_cost != 0: the reference's cost (1 + GravCost) / 2^TimeBin, domain.c:378; 0: equal
 * counts), every particle's state crosses to its owner in one all-to-all of 256-byte records, and every rank stores what
 * it now holds in Peano-Hilbert order.  Results do not depend on the assignment; this restores the balance and the
 * pruning of the exchange after the particles have moved.  *n_new <- the particles this rank holds afterwards, *n_sent
 * <- how many left it.  Particle ids (gl_run_set_ids; default: the global index before the first exchange) travel with
 * the particles: gl_get_run_ids says which particle each row of gl_get_run_state / gl_get_* now is. */
int gl_run_set_ids(gl_ctx * ctx, const int *ids, int where);
int gl_get_run_ids(gl_ctx * ctx, int *ids, int where);
int gl_domain_exchange(gl_ctx * ctx, int by_cost, long long *n_new, long long *n_sent);
/* The same loop on the reference's own arrays: P[0..NumPart) / SphP[0..N_gas) are read once (everything the loop
 * carries: positions, velocities, VelPred, Hsml, entropies, time bins, the accelerations and SphP state of the last
 * force computation), nsteps iterations run on the device, and the structs are written once - where
 * gl_gadget_compute_accelerations moves them through host memory twice PER STEP.  Every particle must be drifted to
 * par->Ti_Current on entry (P.Ti_current, as after move_particles(); Type 3 excepted, predict.c:134); on return all
 * are drifted to the new par->Ti_Current, and P.TimeBin / P.Ti_begstep are current: the caller rebuilds its time-bin
 * lists from them (reconstruct_timebins(), timestep.c) before it touches them.  gl_gadget_bind first. */
int gl_gadget_run_steps(gl_ctx * ctx, void *P, void *SphP, long long NumPart, long long N_gas, gl_timestep_params * par, int nsteps,
			gl_run_info * info);

/* ---- the tree, in the reference's own storage (SURVEY.md 8(b) "Tree (if CPU consumers are kept)") ---------------
 * The sub-grid loops that follow the force path in the reference (sink accretion blackhole.c:794,1201,1351, dust
 * dust.c:60,263,748, setup_smoothinglengths init.c:719, its own force_update_hmax) walk Nodes[] / Nextnode[] with
 * ngb_treefind_* (ngb.c).  Instead of re-building that tree on the host (force_treebuild, ~0.4 us/particle on one
 * core) the GPU tree is written into the reference's arrays: struct NODE / struct extNODE records (allvars.h:1845,
 * :1887) with the threading of force_update_node_recursive (forcetree.c:468-872): nextnode = first daughter,
 * sibling = next entry behind the subtree, Nextnode[i] = entry behind particle i, Father[i] = its parent cell;
 * node k of the library is Nodes[MaxPart + k], the root is Nodes[MaxPart] (density.c:804).  Offsets come from the
 * reference side via offsetof(), like gl_gadget_layout.  extNODE.vs / .divVmax (forcetree.c:684-706, what force_drift_node
 * / force_kick_node read) are formed here - a bottom-up pass in child-slot order from P.Vel and the SphP.v.DivVel of the
 * last density() - when their offsets are >= 0; dp = 0.  Single rank only: no pseudo-particles are inserted. */
typedef struct gl_tree_layout
{
  long long node_stride, ext_stride;	/* sizeof(struct NODE), sizeof(struct extNODE) */
  int n_len, n_center, n_maxsoft;	/* n_maxsoft < 0: the build has no ADAPTIVE_GRAVSOFT_FORGAS */
  int n_s, n_mass, n_bitflags, n_sibling, n_nextnode, n_father, n_Ti_current;
  int e_vs, e_vmax, e_divVmax, e_hmax, e_Ti_lastkicked, e_Flag;
  unsigned int bitflags_value;		/* what a cell with >= 2 particles carries: 1 << BITFLAG_MULTIPLEPARTICLES */
} gl_tree_layout;
/* nodes_base / extnodes_base: first allocated node (= Nodes + MaxPart); nextnode, father: MaxPart ints each, indexed
 * like P[]; fails if more than max_nodes cells exist.  *numnodes <- Numnodestree. */
int gl_export_tree(gl_ctx * ctx, const gl_tree_layout * lay, void *nodes_base, void *extnodes_base, int *nextnode,
		   int *father, long long MaxPart, long long max_nodes, int Ti_Current, long long *numnodes);

/* ---- the sub-grid neighbour loops on the same tree -------------------------------------------
 * What accel.c:189-196 calls right behind hydro_force(): sink accretion (blackhole_accretion blackhole.c:70 with its two
 * neighbour loops blackhole_evaluate :794 and blackhole_evaluate_swallow :1201 over ngb_treefind_blackhole :1351) and the
 * dust loops (dust_density dust.c:60 -> dust_evaluate_density :748; dust_drag :263 -> dust_evaluate_select :889 over
 * ngb_treefind_dust_active :1235), under the shipped Makefile's switches.  They search the tree of the last build
 * (gl_gravity_tree / gl_density of this step; positions, masses, types, Hsml and time bins are the loaded ones, Hsml
 * the one gl_density left) - no second tree, no Nodes[] on the host.  Single rank.
 * Arrays are indexed like P[] (n elements, gas first), `where` as in gl_load_particles; arrays marked in/out are updated
 * in place.  `active`: the active list in FirstActiveParticle order (host pointer; NULL = every particle in index
 * order) - the order decides which sink gets a particle two sinks claim (blackhole.c:991, :1030: the last one). */
typedef struct gl_subgrid_params
{
  double SMBHmass, InnerBoundary, SinkBoundary, SofteningBndry;	/* All.* of the same names (.txt:167-171, :102) */
  double ViscousTimescale, BlackHoleFeedbackFactor;	/* .txt:138, :126 */
  double UnitTime_in_s, UnitMass_in_g, UnitLength_in_cm, UnitVelocity_in_cm_per_s, UnitEnergy_in_cgs, UnitDensity_in_cgs;
  double MeanWeight, MinEgySpec;
  double xbh[3];		/* All.xbh, ybh, zbh (FindQuasars, blackhole.c:1482) */
} gl_subgrid_params;
int gl_set_subgrid_params(gl_ctx * ctx, const gl_subgrid_params * par);

typedef struct gl_sink_state
{
  double *vel;			/* P.Vel[3n], every type (NOT SphP.VelPred); in/out: sinks that accreted */
  const int *id;		/* P.ID[n] */
  int *swallow_id;		/* P.SwallowID[n]; in/out */
  double *mass;			/* out (may be NULL): P.Mass[n] afterwards - swallowed particles 0, sinks grown.  The input
				   is the loaded mass; the library's own copy is updated, so the dust loops see it */
  double *bh_mass, *bh_mdot, *accdisc_mass, *dust_mass, *total_mass;	/* P.BH_Mass ... P.Total_Mass [n]; in/out at Type 5 */
  const double *sink_density;	/* P.b1.BH_Density[n] at Type 5; NULL = what gl_density computed */
  double *injected_bh_energy;	/* SphP.i.Injected_BH_Energy, indexed like P[] (n elements, gas entries used); in/out */
} gl_sink_state;
/* n_swallowed[3] <- N_gas_swallowed, N_BH_swallowed, N_dust_swallowed (blackhole.c:300, :648).
 * Two statements of the reference are restated as written, not as probably meant: the bound-orbit test for dust near a
 * sink uses the distance of the LAST GAS neighbour visited - of this sink or, before it has met gas, of the sinks visited
 * before it (blackhole.c:1024 reads `r`, set only at :1042; the compiled reference keeps it in a stack slot across calls) -, and
 * BH_accreted_BHMass / BH_accreted_DustMass share one union (allvars.h:1261-1267), so a sink's BH_Mass grows by the
 * accreted DUST mass (blackhole.c:699-701). */
int gl_blackhole_accretion(gl_ctx * ctx, const int *active, long long nactive, const gl_sink_state * st, int where,
			   long long n_swallowed[3]);

/* dust_density(): P.d7.DUST_particle_density of the active Type-2 particles (other entries untouched) */
int gl_dust_density(gl_ctx * ctx, const int *active, long long nactive, double *dust_particle_density, int where);

typedef struct gl_dust_state
{
  double *vel;			/* P.Vel[3n]; in/out: active dust (drag) and the gas around it (momentum returned) */
  const double *gravaccel;	/* P.g.GravAccel[3n]; NULL = the last gl_gravity_tree result */
  const double *gas_density, *gas_entropy, *gas_vel;	/* P.d1/d2/d3 at Type 2 ([n], [n], [3n]); NULL = gl_density's */
  const double *dust_particle_density;	/* P.d7 [n] (gl_dust_density) */
  double *dust_particle_velocity;	/* P.d9 [3n]; in/out (divided by d7 in place, dust.c:368-370); NULL = zeros */
  const double *dust_radius;	/* P.DustRadius[n], cm */
  double *dust_vcoll, *delta_dust_momentum, *new_drag_acc, *delta_drag_energy;	/* out at active Type 2: [n], [3n], [3n], [n] */
  double *entropy;		/* SphP.Entropy, indexed like P[]; in/out at gas; NULL = the loaded one, updated in the library */
  double *drag_heating;		/* SphP.dh.DragHeating, indexed like P[]; in/out at gas */
} gl_dust_state;
int gl_dust_drag(gl_ctx * ctx, const int *active, long long nactive, const gl_dust_state * st, int where);

/* The same three on the reference's own P[] / SphP[] (after gl_gadget_compute_accelerations / gl_gadget_pull of this
 * step; gl_gadget_bind supplies the offsets the force path already uses, this struct the ones only these loops need).
 * The search runs first and returns the indices of the particles the loops touch; only their fields are gathered from
 * the structs, sent, updated and scattered back - for the two sinks of the shipped run a few hundred particles. */
typedef struct gl_gadget_subgrid_layout
{
  int p_ID, p_BH_Mass, p_BH_Mdot, p_AccDisc_Mass, p_Dust_Mass, p_Total_Mass;
  int p_DustParticleDensity, p_DustParticleVelocity;	/* d7, d9 */
  int p_DustRadius, p_DustVcoll, p_DeltaDustMomentum, p_NewDragAcc, p_DeltaDragEnergy;
  int s_Injected_BH_Energy, s_DragHeating;
} gl_gadget_subgrid_layout;
int gl_gadget_blackhole_accretion(gl_ctx * ctx, const gl_gadget_subgrid_layout * lay, void *P, void *SphP, const int *active,
				  long long nactive, long long n_swallowed[3]);
int gl_gadget_dust_density(gl_ctx * ctx, const gl_gadget_subgrid_layout * lay, void *P, const int *active, long long nactive);
int gl_gadget_dust_drag(gl_ctx * ctx, const gl_gadget_subgrid_layout * lay, void *P, void *SphP, const int *active, long long nactive);

/* ---- multi-GPU drop-in: every rank keeps its own P[] / SphP[] (its Peano-Hilbert segment), as every MPI rank of the
 * reference does; the sources are replicated with gl_load_particles_gathered --------------------------------------- */
/* pack this rank's arrays into its all-gather block (layout of gl_load_particles_gathered, padded to m particles) and
 * send it to the device buffers f64_dev[11*m], i32_dev[3*m]; global_offset = index of the rank's first particle */
int gl_gadget_pack_segment(gl_ctx * ctx, const void *P, const void *SphP, long long NumPart, long long N_gas, long long m,
			   long long global_offset, double *f64_dev, int *i32_dev);
/* return one result block (what = 1 gravity, 2 density, 4 hydro) of the rank's own particles to its arrays while the
 * next phase computes; gl_gadget_push_wait() blocks until every pending block has been scattered */
int gl_gadget_push_segment_async(gl_ctx * ctx, void *P, void *SphP, int what);
int gl_gadget_push_wait(gl_ctx * ctx);

/* ---- on-disk formats: initial-conditions reader and snapshot writer --------------------------
 * The reference's binary formats 1 and 2 and its HDF5 format 3 (All.ICFormat / All.SnapFormat; the shipped parameter file
 * uses 3), single file, as the shipped -D list stores them: 4-byte floats (neither INPUT_ nor
 * OUTPUT_IN_DOUBLEPRECISION), 4-byte IDs (no LONGIDS), Fortran record markers, blocks in `enum iofields` order
 * (allvars.h:1762).  Files written here are byte-identical to what savepositions() (io.c:33 -> write_file :2521 ->
 * fill_write_buffer :173) writes from the same P[]/SphP[]; files are read the way read_ic() (read_ic.c:28 -> read_file
 * :634 -> empty_read_buffer :238) reads them.  Format 3 (io.c:2619-2633, :2715-2800, :2866; read_ic.c:648-1060, :1333): the
 * container is written and parsed by the library itself (csrc/gl_hdf5.h - the image has no HDF5 library): "/Header" with
 * the header fields as attributes, "/PartType<t>/<get_dataset_name()>" contiguous datasets of the same 4-byte values the
 * binary formats hold, in the structures libhdf5 emits for the reference's calls; the writer appends ".hdf5" to fname as
 * the reference does.  The reader takes files libhdf5 wrote with default settings (old-style groups, contiguous or compact
 * datasets, any integer / float width - converted like H5Dread does) and names what it does not handle (chunked or
 * compressed datasets, libver='latest' files, big-endian data).  Host code only - no context, no GPU. */
typedef struct gl_io_header	/* struct io_header, allvars.h:1733 (256 bytes) */
{
  int npart[6];
  double mass[6];
  double time, redshift;
  int flag_sfr, flag_feedback;
  unsigned int npartTotal[6];
  int flag_cooling, num_files;
  double BoxSize, Omega0, OmegaLambda, HubbleParam;
  int flag_stellarage, flag_metals;
  unsigned int npartTotalHighWord[6];
  int flag_entropy_instead_u, flag_doubleprecision;
  char fill[56];
} gl_io_header;
/* offsets of the fields only the files carry, next to gl_gadget_layout; < 0: the build has no such field and its
 * block is absent (blockpresent(), io.c:1639) */
typedef struct gl_io_layout
{
  int p_ID;			/* IO_ID */
  int p_DustRadius;		/* IO_DUST (-DDUST), every type */
  int p_StellarAge;		/* IO_AGE (-DSTELLARAGE), Type 4, only with StarformationOn */
  int p_BH_Mass, p_BH_Mdot, p_AccDisc_Mass, p_BH_Mseed;	/* IO_BHMASS ... IO_BHMSEED (-DBLACK_HOLES), Type 5 */
  int p_NewDensity, p_DeltaPhotonMomentum, p_OldPhotonMomentum, p_BirthPhotonEnergy;	/* -DVIRTUAL, Type 3 */
  int s_Ne;			/* IO_NE (-DCOOLING), gas, only with CoolingOn */
  int s_accreted_mass, s_accreted_momentum;	/* IO_ACCMASS, IO_ACCMOM (-DBLACK_HOLES -DVIRTUAL), gas */
} gl_io_layout;
#define GL_IO_BUILD_COOLING 1	/* the reference's compile-time switches that reach the files */
#define GL_IO_BUILD_SFR 2
#define GL_IO_BUILD_STELLARAGE 4
#define GL_IO_BUILD_METALS 8
#define GL_IO_BUILD_BH_SWALLOWGAS 16
typedef struct gl_io_params
{
  int format;			/* All.SnapFormat / All.ICFormat: 1, 2 or 3 */
  int build;			/* GL_IO_BUILD_* mask: header flags (io.c:2589-2608), MassTable rules (read_ic.c:162-179) */
  int CoolingOn, StarformationOn;	/* All.*: NE/NH and SFR/AGE blocks exist (io.c:1652-1680) */
  int RestartFlag;		/* reader: 0 = initial conditions (blocks behind U are not read, read_ic.c:891), 2 = snapshot */
  int Ti_Current;		/* writer: velocities are synchronised to All.Ti_Current (io.c:253-299) */
  double Timebase_interval;
  double Time;			/* writer: header.time */
  double MassTable[6];		/* All.MassTable: types with an entry != 0 have no MASS record; the reader sets it from the file */
  double MinEgySpec;		/* All.MinEgySpec: floor of U (io.c:322) and of the entropy read (read_ic.c:215) */
  double GammaMinus1;		/* GAMMA_MINUS1 (allvars.h:64-65): U = A / (gamma-1) * rho^(gamma-1) */
  double InitGasTemp, InitGasU;	/* reader, RestartFlag 0: gas read with u == 0 gets All.InitGasU if InitGasTemp > 0 (read_ic.c:196-210) */
  double BoxSize, Omega0, OmegaLambda, HubbleParam;	/* header only */
} gl_io_params;
/* savepositions(): nh / sfr = the two derived records that come out of the cooling / star-formation modules
 * (AbundanceRatios, get_starformation_rate: outside the path), indexed like SphP[]; NULL writes zeros.
 * err (may be NULL) receives the message of a non-zero return. */
int gl_snapshot_write(const char *fname, const gl_gadget_layout * lay, const gl_io_layout * io, const gl_io_params * par,
		      const void *P, const void *SphP, long long NumPart, const double *nh, const double *sfr, char *err,
		      int errlen);
/* read_ic(): fills P[0..*NumPart) / SphP[0..*N_gas) (gas first, types in order) and only the fields the reference
 * sets; par->MassTable is updated the way read_file / read_ic leave All.MassTable; *hdr <- the file header.
 * Non-zero return where the reference calls endrun(): unreadable file, bad header record (890), double-precision
 * file (read_ic.c:744-750), not enough room (172, 173), wrong format-2 label (1890); also for a record whose length does
 * not match the header's particle numbers (the reference reads on regardless) and for multi-file snapshots. */
/* only the header record - what read_file() looks at before P[] / SphP[] exist (read_ic.c:688-697, :754-790) */
int gl_snapshot_read_header(const char *fname, int format, gl_io_header * hdr, char *err, int errlen);
int gl_snapshot_read(const char *fname, const gl_gadget_layout * lay, const gl_io_layout * io, gl_io_params * par, void *P,
		     void *SphP, long long MaxPart, long long MaxPartSph, long long *NumPart, long long *N_gas,
		     gl_io_header * hdr, char *err, int errlen);

/* any numeric dataset ("/PartType0/Coordinates") or attribute ("/Header@Time") of an HDF5 file as doubles; dims[4] and
 * *rank <- its shape, out (may be NULL) <- up to cap values in row-major order */
int gl_hdf5_read_f64(const char *fname, const char *path, double *out, long long cap, long long dims[4], int *rank, char *err,
		     int errlen);

#ifdef __cplusplus
}
#endif
#endif
