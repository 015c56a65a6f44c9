/* gl_gadget_glue.c - the reference-side binding of libgl_b200.so.
 *
 * This file is what a maintainer of the reference ADDS to its tree (it is compiled with the reference's own
 * allvars.h/proto.h and -D list, so every offsetof() below is exact for that build) and then REMOVES
 * gravtree.o density.o hydra.o from OBJS (Makefile:845-856), linking with -lgl_b200 -lcudart instead.
 * It defines the reference's own global entry points for the force path
 *     void gravity_tree(void)            proto.h / gravtree.c:27
 *     void density(void)                 proto.h / density.c:89
 *     void force_update_hmax(void)       forcetree.h / forcetree.c:1661   (only when GL_TAKE_OVER_HMAX is set)
 *     void hydro_force(void)             proto.h / hydra.c:145
 *     void compute_accelerations(int)    proto.h / accel.c:27             (only when GL_TAKE_OVER_ACCEL is set)
 * the sub-grid neighbour loops that follow it in accel.c:189-196, on the GPU tree of the step (-DGL_SUBGRID, one rank;
 * remove blackhole_accretion / dust_density / dust_drag from blackhole.c / dust.c or rename them with objcopy):
 *     void blackhole_accretion(void)     proto.h / blackhole.c:70
 *     void dust_density(void)            proto.h / dust.c:60
 *     void dust_drag(void)               proto.h / dust.c:263
 * and, each behind its own switch, the host-side neighbours of the path (no GPU involved):
 *     void advance_and_find_timesteps(void)   proto.h / timestep.c:29     -DGL_TIMESTEP
 *     void peano_hilbert_order(void)          proto.h / peano.c:25        -DGL_PEANO_ORDER
 *     void domain_Decomposition(void)         proto.h / domain.c:100      -DGL_DOMAIN (one rank)
 *     void read_ic(char *), void savepositions(int)   read_ic.c:28, io.c:33   -DGL_SNAPSHOT_IO (formats 1, 2 and 3, one rank)
 * on top of the C ABI of include/gl_b200.h, keeps the reference's side effects (All.ErrTolTheta = 0 after the
 * first call, TreeReconstructFlag = 0, All.TotNumOfForces, CPU_Step[] buckets), leaves the mymalloc stack
 * untouched (mymalloc.c:159-164) and maps any library error onto endrun() (endrun.c:23).
 *
 * It is compiled in the build container by oracle/refbuild (target `dropin`) into oracle/_ref/ref_dropin,
 * where the reference's own harness runs with THESE symbols instead of gravtree.c/density.c/hydra.c; targets
 * `host_dropin` and `io_dropin` do the same for the host-side entry points (ref_host_dropin, ref_io_dropin; with
 * -DGL_NO_FORCE_PATH the force path stays the reference's, so those binaries run without a GPU).
 */
#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>
#include <math.h>
#include <mpi.h>
#include "allvars.h"
#include "proto.h"
#include "forcetree.h"
#include "gl_b200.h"

static gl_ctx *GL;
static gl_gadget_layout GL_Layout;

static void gl_fail(const char *where)
{
  printf("libgl_b200: %s failed: %s\n", where, GL ? gl_last_error(GL) : "no context");
  fflush(stdout);
  endrun(8800);
}

static void gl_sync_params(void)
{
  gl_params p;
  int i;
  p.ErrTolTheta = All.ErrTolTheta;
  p.ErrTolForceAcc = All.ErrTolForceAcc;
  p.TypeOfOpeningCriterion = All.TypeOfOpeningCriterion;
  p.DesNumNgb = All.DesNumNgb;
  p.MaxNumNgbDeviation = All.MaxNumNgbDeviation;
  p.BlackHoleNgbFactor = All.BlackHoleNgbFactor;
  p.ArtBulkViscConst = All.ArtBulkViscConst;
  p.G = All.G;
  for(i = 0; i < 6; i++) p.ForceSoftening[i] = All.ForceSoftening[i];
  p.MinGasHsml = All.MinGasHsml;
  p.Timebase_interval = All.Timebase_interval;
  p.Ti_Current = All.Ti_Current;
  p.ComovingIntegrationOn = All.ComovingIntegrationOn;
  p.OmegaLambdaHubble2 = All.OmegaLambda * All.Hubble * All.Hubble;
  for(i = 0; i < 6; i++) p.SofteningTable[i] = All.SofteningTable[i];
  if(gl_set_params(GL, &p)) gl_fail("gl_set_params");
}

static void gl_fill_layout(void)
{
  gl_gadget_layout L;
  L.p_stride = sizeof(struct particle_data);
  L.sph_stride = sizeof(struct sph_particle_data);
  L.p_Pos = offsetof(struct particle_data, Pos);
  L.p_Vel = offsetof(struct particle_data, Vel);
  L.p_Mass = offsetof(struct particle_data, Mass);
  L.p_GravAccel = offsetof(struct particle_data, g.GravAccel);
  L.p_OldAcc = offsetof(struct particle_data, OldAcc);
  L.p_Hsml = offsetof(struct particle_data, Hsml);
  L.p_NumNgb = offsetof(struct particle_data, n.NumNgb);
  L.p_GravCost = offsetof(struct particle_data, GravCost);
  L.p_Ti_begstep = offsetof(struct particle_data, Ti_begstep);
  L.p_Type = offsetof(struct particle_data, Type);
  L.p_TimeBin = offsetof(struct particle_data, TimeBin);
  L.p_SinkDensity = offsetof(struct particle_data, b1.BH_Density);
  L.p_SinkEntropy = offsetof(struct particle_data, b2.BH_Entropy);
  L.p_SinkGasVel = offsetof(struct particle_data, b3.BH_SurroundingGasVel);
  L.p_DustDensity = offsetof(struct particle_data, d1.DUST_Density);
  L.p_DustEntropy = offsetof(struct particle_data, d2.DUST_Entropy);
  L.p_DustGasVel = offsetof(struct particle_data, d3.DUST_SurroundingGasVel);
  L.p_SwallowID = offsetof(struct particle_data, SwallowID);
  L.s_Entropy = offsetof(struct sph_particle_data, Entropy);
  L.s_Pressure = offsetof(struct sph_particle_data, Pressure);
  L.s_VelPred = offsetof(struct sph_particle_data, VelPred);
  L.s_MaxSignalVel = offsetof(struct sph_particle_data, MaxSignalVel);
  L.s_Density = offsetof(struct sph_particle_data, d.Density);
  L.s_DtEntropy = offsetof(struct sph_particle_data, e.DtEntropy);
  L.s_HydroAccel = offsetof(struct sph_particle_data, a.HydroAccel);
  L.s_DhsmlDensityFactor = offsetof(struct sph_particle_data, h.DhsmlDensityFactor);
  L.s_DivVel = offsetof(struct sph_particle_data, v.DivVel);
  L.s_CurlVel = offsetof(struct sph_particle_data, r.CurlVel);
  L.p_Ti_current = offsetof(struct particle_data, Ti_current);
#ifdef DUST
  L.s_DragAccel = offsetof(struct sph_particle_data, da.DragAccel);
#else
  L.s_DragAccel = -1;
#endif
  GL_Layout = L;
}

#ifndef GL_GPUS_PER_NODE
#define GL_GPUS_PER_NODE 8
#endif
static void gl_attach(void)
{
  if(GL) return;
  /* one rank <-> one GPU of the node */
  if(gl_create(&GL, ThisTask % GL_GPUS_PER_NODE, All.MaxPart)) gl_fail("gl_create");
  gl_fill_layout();
  if(gl_gadget_bind(GL, &GL_Layout)) gl_fail("gl_gadget_bind");
#ifndef GL_SINGLE_RANK_ONLY
  if(NTask > 1)
    {
      /* the communicator of the force path: the id travels the way the reference broadcasts All (begrun.c:1922) */
      char id[128];
      if(ThisTask == 0 && gl_comm_unique_id(id)) gl_fail("gl_comm_unique_id");
      MPI_Bcast(id, 128, MPI_BYTE, 0, MPI_COMM_WORLD);
      if(gl_comm_init(GL, ThisTask, NTask, id)) gl_fail("gl_comm_init");
    }
#endif
}

/* the library evaluates the particles it is told to; the reference's active list is
 * FirstActiveParticle / NextActiveParticle[] (run.c:190-343) */
static int gl_send_active_list(void)
{
  int i, n = 0, all, *idx;
  for(i = FirstActiveParticle; i >= 0; i = NextActiveParticle[i]) n++;
  all = (n == NumPart);
  if(all)
    {
      if(gl_set_targets(GL, NULL, -1, 0)) gl_fail("gl_set_targets");
      return n;
    }
  idx = (int *) malloc(sizeof(int) * (n > 0 ? n : 1));	/* private scratch: the mymalloc stack stays as found */
  for(i = FirstActiveParticle, n = 0; i >= 0; i = NextActiveParticle[i]) idx[n++] = i;
  if(gl_set_targets(GL, idx, n, 0)) gl_fail("gl_set_targets");
  free(idx);
  return n;
}

static void gl_pull(void)
{
  int i;
  gl_attach();
  gl_sync_params();
  /* sources must be current: the reference drifts lazily inside its walks (forcetree.c:1911, ngb.c:57) */
#ifdef GL_REFERENCE_DRIFT
  for(i = 0; i < NumPart; i++)
    if(P[i].Ti_current != All.Ti_Current) drift_particle(i, All.Ti_Current);
#else
  /* the same statements (predict.c:129), on all host cores */
  if(All.ComovingIntegrationOn)
    gl_fail("comoving drift");
  if(gl_gadget_drift(&GL_Layout, P, SphP, NumPart, N_gas, All.Ti_Current, All.Timebase_interval, All.MinGasHsml, NULL))
    endrun(12);
  (void) i;
#endif
  if(gl_gadget_pull(GL, P, SphP, NumPart, N_gas)) gl_fail("gl_gadget_pull");
  gl_set_tree_reconstruct_flag(GL, 1);
}

static int gl_loaded_for = -1;	/* All.NumCurrentTiStep the device copy belongs to */

/* NTask > 1.  The reference's own functions work at any NTask: they export targets that touch a pseudo-particle to its
 * owner and add the partial results (gravtree.c:175-347, density.c:193-342, hydra.c:274-520).  The library does the
 * opposite - every rank imports the locally-essential part of the other ranks' trees over NCCL and evaluates its own
 * particles completely (include/gl_b200.h "multi-rank mode") - behind the same three entry points: gl_attach() creates
 * the communicator from an id that rank 0 draws and MPI_Bcast distributes, and from then on gl_gravity_tree /
 * gl_density / gl_force_update_hmax / gl_hydro_force are collective, exactly like the functions they replace.
 * One rank <-> one GPU: ranks of a node take the GPUs in PTask-local order (GL_GPUS_PER_NODE, default 8).
 * -DGL_SINGLE_RANK_ONLY keeps the old behaviour of stopping at NTask != 1 (a build without NCCL on the nodes). */
static void gl_require_single_rank(const char *where)
{
#ifdef GL_SINGLE_RANK_ONLY
  if(NTask != 1)
    {
      printf("libgl_b200 glue: %s() of this build handles NTask == 1 only (NTask = %d)\n", where, NTask);
      fflush(stdout);
      endrun(8801);
    }
#else
  (void) where;
#endif
}

static void gl_ensure_loaded(void)
{
  gl_attach();
  /* the active list first: gl_gadget_pull() uploads the SphP state of the INACTIVE particles (hydro reads the density /
   * pressure of every neighbour, hydra.c:1236) only when the list of THIS step is a subset - with the previous step's
   * list still set, a full step followed by a sub-step would read stale, undrifted values */
  gl_send_active_list();
  if(gl_loaded_for != All.NumCurrentTiStep || TreeReconstructFlag)
    {
      gl_pull();
      gl_loaded_for = All.NumCurrentTiStep;
    }
}

#ifndef GL_NO_FORCE_PATH	/* (a build that only takes the host-side pieces below: the CPU test binaries of oracle/refbuild) */
void gravity_tree(void)
{
  double t0 = second();
  int nactive;
  gl_require_single_rank("gravity_tree");
  gl_ensure_loaded();
  nactive = gl_send_active_list();
  gl_sync_params();
  if(gl_gravity_tree(GL)) gl_fail("gl_gravity_tree");
  if(gl_gadget_push(GL, P, SphP, 1)) gl_fail("gl_gadget_push");
#ifdef GL_EXPORT_TREE
  /* The reference's remaining tree walkers (sink / dust loops, setup_smoothinglengths, its own force_update_hmax) find
   * the GPU tree in Nodes[] / Extnodes[] / Nextnode[] / Father[] - no host tree build at all. */
  if(TreeReconstructFlag)
    {
      gl_tree_layout T;
      long long nn = 0;
      T.node_stride = sizeof(struct NODE);
      T.ext_stride = sizeof(struct extNODE);
      T.n_len = offsetof(struct NODE, len);
      T.n_center = offsetof(struct NODE, center);
#ifdef ADAPTIVE_GRAVSOFT_FORGAS
      T.n_maxsoft = offsetof(struct NODE, maxsoft);
#else
      T.n_maxsoft = -1;
#endif
      T.n_s = offsetof(struct NODE, u.d.s);
      T.n_mass = offsetof(struct NODE, u.d.mass);
      T.n_bitflags = offsetof(struct NODE, u.d.bitflags);
      T.n_sibling = offsetof(struct NODE, u.d.sibling);
      T.n_nextnode = offsetof(struct NODE, u.d.nextnode);
      T.n_father = offsetof(struct NODE, u.d.father);
      T.n_Ti_current = offsetof(struct NODE, Ti_current);
      T.e_vs = offsetof(struct extNODE, vs);
      T.e_vmax = offsetof(struct extNODE, vmax);
      T.e_divVmax = offsetof(struct extNODE, divVmax);
      T.e_hmax = offsetof(struct extNODE, hmax);
      T.e_Ti_lastkicked = offsetof(struct extNODE, Ti_lastkicked);
      T.e_Flag = offsetof(struct extNODE, Flag);
      T.bitflags_value = 1u << BITFLAG_MULTIPLEPARTICLES;
      if(gl_export_tree(GL, &T, Nodes_base, Extnodes_base, Nextnode, Father, All.MaxPart, MaxNodes, All.Ti_Current, &nn))
	gl_fail("gl_export_tree");
      Numnodestree = (int) nn;
    }
#endif
#ifdef GL_KEEP_CPU_TREE
  /* The reference's sub-grid physics (blackhole.c:1351, dust.c:1235), setup_smoothinglengths (init.c:719) and its own
   * force_update_hmax still walk the CPU Nodes[] tree.  Until those loops are ported too a host
   * that needs them keeps building it here - at ~0.4 us/particle this dominates the step, so leave it off when only
   * the force path is used. */
  if(TreeReconstructFlag) force_treebuild(NumPart, NULL);
#endif
  if(All.TypeOfOpeningCriterion == 1) All.ErrTolTheta = 0;	/* gravtree.c:396-397 */
  TreeReconstructFlag = 0;
  All.TotNumOfForces += nactive;	/* gravtree.c:788 (GlobNumForceUpdate) */
  CPU_Step[CPU_TREEWALK1] += timediff(t0, second());
  if(ThisTask == 0 && FdTimings)
    {
      /* the block gravity_tree() appends to timings.txt (gravtree.c:803-825): same lines, the walk time is the GPU
         kernel's, nothing is exported (ex-frac 0), one iteration, balance 1 (one rank = one GPU) */
      gl_stats st;
      double tw;
      long long nf = nactive, tot = All.TotNumOfForces;
      gl_get_stats(GL, &st);
      tw = st.ms_gravity_walk * 1e-3 + 1.0e-20;
      fprintf(FdTimings, "Step= %d  t= %g  dt= %g \n", All.NumCurrentTiStep, All.Time, All.TimeStep);
      fprintf(FdTimings, "Nf= %d%09d  total-Nf= %d%09d  ex-frac= %g (%g) iter= %d\n", (int) (nf / 1000000000), (int) (nf % 1000000000),
	      (int) (tot / 1000000000), (int) (tot % 1000000000), 0.0, 0.0, 1);
      fprintf(FdTimings, "work-load balance: %g (%g %g) rel1to2=%g   max=%g avg=%g\n", 1.0, 1.0, 1.0, 1.0, tw, tw);
      fprintf(FdTimings, "particle-load balance: %g\n", (NumPart / ((double) All.TotNumPart)) * NTask);
      fprintf(FdTimings, "max. nodes: %d, filled: %g\n", (int) st.numnodes, st.numnodes / (All.TreeAllocFactor * All.MaxPart + NTopnodes));
      fprintf(FdTimings, "part/sec=%g | %g  ia/part=%g (%g)\n", nf / tw, nf / tw, ((double) st.interactions) / (nf > 0 ? nf : 1), 0.0);
      fprintf(FdTimings, "\n");
      fflush(FdTimings);
    }
}

void density(void)
{
  double t0 = second();
  gl_require_single_rank("density");
  gl_ensure_loaded();
  if(gl_density(GL)) gl_fail("gl_density");
  if(gl_gadget_push(GL, P, SphP, 2)) gl_fail("gl_gadget_push");
  CPU_Step[CPU_DENSCOMPUTE] += timediff(t0, second());
}

void hydro_force(void)
{
  double t0 = second();
  gl_require_single_rank("hydro_force");
  gl_ensure_loaded();
  if(gl_force_update_hmax(GL)) gl_fail("gl_force_update_hmax");
  if(gl_hydro_force(GL)) gl_fail("gl_hydro_force");
  if(gl_gadget_push(GL, P, SphP, 4)) gl_fail("gl_gadget_push");
  CPU_Step[CPU_HYDCOMPUTE] += timediff(t0, second());
}

#endif /* GL_NO_FORCE_PATH */

#ifdef GL_TAKE_OVER_ACCEL
/* the whole force part of accel.c:27-106 in one pipelined call; the sub-grid physics that follows in the
 * reference (accel.c:176-256) still runs on the host and still finds its inputs in P[]/SphP[] */
void gl_compute_force_path(int mode)
{
  gl_require_single_rank("compute_accelerations");
  gl_attach();
  gl_sync_params();
  gl_send_active_list();
  gl_set_tree_reconstruct_flag(GL, 1);
  if(gl_gadget_compute_accelerations(GL, P, SphP, NumPart, N_gas, mode)) gl_fail("gl_gadget_compute_accelerations");
  if(All.TypeOfOpeningCriterion == 1) All.ErrTolTheta = 0;
  TreeReconstructFlag = 0;
}
#endif

#ifdef GL_SUBGRID
/* ---- blackhole_accretion() / dust_density() / dust_drag() (accel.c:189-196) on the tree the force path built ----------
 * The reference walks Nodes[] on the host for these (blackhole.c:1351, dust.c:1235); here the searches run on the device
 * copy of the same step, and only the particles they touch are gathered from P[] / SphP[] and written back.  Kept from
 * the reference: the log lines, the swallow counters, the TimeBin_BH_* sums and the blackholes.txt record.  One rank. */
static gl_gadget_subgrid_layout GL_SubLayout;
static void gl_subgrid_setup(const char *where)
{
  gl_subgrid_params sp;
  gl_gadget_subgrid_layout *S = &GL_SubLayout;
  if(NTask != 1)
    {
      printf("libgl_b200 glue: %s() on the GPU tree handles NTask == 1 only (NTask = %d)\n", where, NTask);
      fflush(stdout);
      endrun(8802);
    }
  if(All.ComovingIntegrationOn) gl_fail("comoving sub-grid loops");
  gl_ensure_loaded();
  S->p_ID = offsetof(struct particle_data, ID);
  S->p_BH_Mass = offsetof(struct particle_data, BH_Mass);
  S->p_BH_Mdot = offsetof(struct particle_data, BH_Mdot);
  S->p_AccDisc_Mass = offsetof(struct particle_data, AccDisc_Mass);
  S->p_Dust_Mass = offsetof(struct particle_data, Dust_Mass);
  S->p_Total_Mass = offsetof(struct particle_data, Total_Mass);
  S->p_DustParticleDensity = offsetof(struct particle_data, d7.DUST_particle_density);
  S->p_DustParticleVelocity = offsetof(struct particle_data, d9.DUST_particle_velocity);
  S->p_DustRadius = offsetof(struct particle_data, DustRadius);
  S->p_DustVcoll = offsetof(struct particle_data, DustVcoll);
  S->p_DeltaDustMomentum = offsetof(struct particle_data, DeltaDustMomentum);
  S->p_NewDragAcc = offsetof(struct particle_data, NewDragAcc);
  S->p_DeltaDragEnergy = offsetof(struct particle_data, DeltaDragEnergy);
  S->s_Injected_BH_Energy = offsetof(struct sph_particle_data, i.Injected_BH_Energy);
  S->s_DragHeating = offsetof(struct sph_particle_data, dh.DragHeating);
  sp.SMBHmass = All.SMBHmass; sp.InnerBoundary = All.InnerBoundary; sp.SinkBoundary = All.SinkBoundary;
  sp.SofteningBndry = All.SofteningBndry; sp.ViscousTimescale = All.ViscousTimescale;
  sp.BlackHoleFeedbackFactor = All.BlackHoleFeedbackFactor;
  sp.UnitTime_in_s = All.UnitTime_in_s; sp.UnitMass_in_g = All.UnitMass_in_g; sp.UnitLength_in_cm = All.UnitLength_in_cm;
  sp.UnitVelocity_in_cm_per_s = All.UnitVelocity_in_cm_per_s; sp.UnitEnergy_in_cgs = All.UnitEnergy_in_cgs;
  sp.UnitDensity_in_cgs = All.UnitDensity_in_cgs; sp.MeanWeight = All.MeanWeight; sp.MinEgySpec = All.MinEgySpec;
  sp.xbh[0] = All.xbh; sp.xbh[1] = All.ybh; sp.xbh[2] = All.zbh;
  gl_sync_params();
  if(gl_set_subgrid_params(GL, &sp)) gl_fail("gl_set_subgrid_params");
}
static int *gl_active_array(int *n)
{
  int i, k = 0, *idx;
  for(i = FirstActiveParticle; i >= 0; i = NextActiveParticle[i]) k++;
  idx = (int *) malloc(sizeof(int) * (k > 0 ? k : 1));
  for(i = FirstActiveParticle, k = 0; i >= 0; i = NextActiveParticle[i]) idx[k++] = i;
  *n = k;
  return idx;
}

void blackhole_accretion(void)
{
  int n, bin, nact, *act;
  long long sw[3];
  double mass_holes = 0, mass_real = 0, mdot = 0;
  printf("Beginning black-hole accretion\n");
  fflush(stdout);
  CPU_Step[CPU_MISC] += measure_time();
  gl_subgrid_setup("blackhole_accretion");
  for(n = 0; n < TIMEBINS; n++)
    if(TimeBinActive[n])
      {
	TimeBin_BH_mass[n] = 0;
	TimeBin_BH_dynamicalmass[n] = 0;
	TimeBin_BH_Mdot[n] = 0;
      }
  act = gl_active_array(&nact);
  if(gl_gadget_blackhole_accretion(GL, &GL_SubLayout, P, SphP, act, nact, sw)) gl_fail("gl_gadget_blackhole_accretion");
  printf("Accretion done: %d gas particles swallowed, %d BH particles swallowed, %d Dust particles swallowed\n", (int) sw[0], (int) sw[1],
	 (int) sw[2]);
  fflush(stdout);
  for(n = 0; n < nact; n++)
    if(P[act[n]].Type == 5)
      {
	bin = P[act[n]].TimeBin;
	TimeBin_BH_mass[bin] += P[act[n]].BH_Mass;
	TimeBin_BH_dynamicalmass[bin] += P[act[n]].Mass;
	TimeBin_BH_Mdot[bin] += P[act[n]].BH_Mdot;
      }
  free(act);
  for(bin = 0; bin < TIMEBINS; bin++)
    if(TimeBinCount[bin])
      {
	mass_holes += TimeBin_BH_mass[bin];
	mass_real += TimeBin_BH_dynamicalmass[bin];
	mdot += TimeBin_BH_Mdot[bin];
      }
  if(FdBlackHoles)
    {
      /* blackhole.c:744-760 */
      double mdot_in_msun_per_year = mdot * (All.UnitMass_in_g / SOLAR_MASS) / (All.UnitTime_in_s / SEC_PER_YEAR);
      double total_mdoteddington = 0;
      if(mass_holes > 0)
	total_mdoteddington = mdot / ((4 * M_PI * GRAVITY * C * PROTONMASS / (0.1 * C * C * THOMPSON)) * mass_holes * All.UnitTime_in_s);
      fprintf(FdBlackHoles, "%g %d %g %g %g %g %g\n", All.Time, All.TotBHs, mass_holes, mdot, mdot_in_msun_per_year, mass_real,
	      total_mdoteddington);
      fflush(FdBlackHoles);
    }
  CPU_Step[CPU_BLACKHOLES] += measure_time();
}

void dust_density(void)
{
  int nact, *act;
  printf("Beginning dust density estimation\n");
  fflush(stdout);
  CPU_Step[CPU_MISC] += measure_time();
  gl_subgrid_setup("dust_density");
  act = gl_active_array(&nact);
  if(gl_gadget_dust_density(GL, &GL_SubLayout, P, act, nact)) gl_fail("gl_gadget_dust_density");
  free(act);
}

void dust_drag(void)
{
  int nact, *act;
  printf("Beginning dust management\n");
  fflush(stdout);
  CPU_Step[CPU_MISC] += measure_time();
  gl_subgrid_setup("dust_drag");
  act = gl_active_array(&nact);
  if(gl_gadget_dust_drag(GL, &GL_SubLayout, P, SphP, act, nact)) gl_fail("gl_gadget_dust_drag");
  free(act);
}
#endif /* GL_SUBGRID */

#ifdef GL_SNAPSHOT_IO
/* ---- read_ic() / savepositions() (proto.h; read_ic.c:28, io.c:33) on the library's reader / writer -----------------
 * Formats 1, 2 and 3 (HDF5 through the library's own container code), one file, one rank (a multi-rank run keeps the reference's own
 * read_ic.o / io.o).  Needs no GPU and no context.  The two derived records of a snapshot (NH, SFR) are computed here
 * with the reference's own cooling / star-formation functions and handed to the writer. */
#include <math.h>
#include <string.h>
#include "cooling.h"

static void gl_io_setup(gl_io_layout * io, gl_io_params * par)
{
  int i;
  memset(par, 0, sizeof(*par));
  io->p_ID = offsetof(struct particle_data, ID);
  io->p_DustRadius = io->p_StellarAge = io->p_BH_Mass = io->p_BH_Mdot = io->p_AccDisc_Mass = io->p_BH_Mseed = -1;
  io->p_NewDensity = io->p_DeltaPhotonMomentum = io->p_OldPhotonMomentum = io->p_BirthPhotonEnergy = -1;
  io->s_Ne = io->s_accreted_mass = io->s_accreted_momentum = -1;
#ifdef DUST
  io->p_DustRadius = offsetof(struct particle_data, DustRadius);
#endif
#ifdef STELLARAGE
  io->p_StellarAge = offsetof(struct particle_data, StellarAge);
  par->build |= GL_IO_BUILD_STELLARAGE;
#endif
#ifdef BLACK_HOLES
  io->p_BH_Mass = offsetof(struct particle_data, BH_Mass);
  io->p_BH_Mdot = offsetof(struct particle_data, BH_Mdot);
  io->p_AccDisc_Mass = offsetof(struct particle_data, AccDisc_Mass);
  io->p_BH_Mseed = offsetof(struct particle_data, BH_Mseed);
#ifdef VIRTUAL
  io->s_accreted_mass = offsetof(struct sph_particle_data, accreted_mass);
  io->s_accreted_momentum = offsetof(struct sph_particle_data, accreted_momentum);
#endif
#ifdef SWALLOWGAS
  par->build |= GL_IO_BUILD_BH_SWALLOWGAS;
#endif
#endif
#ifdef VIRTUAL
  io->p_NewDensity = offsetof(struct particle_data, NewDensity);
  io->p_DeltaPhotonMomentum = offsetof(struct particle_data, DeltaPhotonMomentum);
  io->p_OldPhotonMomentum = offsetof(struct particle_data, OldPhotonMomentum);
  io->p_BirthPhotonEnergy = offsetof(struct particle_data, BirthPhotonEnergy);
#endif
#ifdef COOLING
  io->s_Ne = offsetof(struct sph_particle_data, Ne);
  par->build |= GL_IO_BUILD_COOLING;
#endif
#ifdef SFR
  par->build |= GL_IO_BUILD_SFR;
#endif
#ifdef METALS
  par->build |= GL_IO_BUILD_METALS;
#endif
  par->CoolingOn = All.CoolingOn;
  par->StarformationOn = All.StarformationOn;
  par->RestartFlag = RestartFlag;
  par->Ti_Current = All.Ti_Current;
  par->Timebase_interval = All.Timebase_interval;
  par->Time = All.Time;
  for(i = 0; i < 6; i++)
    par->MassTable[i] = All.MassTable[i];
  par->MinEgySpec = All.MinEgySpec;
  par->GammaMinus1 = GAMMA_MINUS1;
  par->InitGasTemp = All.InitGasTemp;
  par->BoxSize = All.BoxSize;
  par->Omega0 = All.Omega0;
  par->OmegaLambda = All.OmegaLambda;
  par->HubbleParam = All.HubbleParam;
}

static void gl_io_fail(const char *what, const char *err, int code)
{
  printf("%s: %s\n", what, err);
  fflush(stdout);
  endrun(code);
}

void read_ic(char *fname)
{
  gl_io_layout io;
  gl_io_params par;
  gl_io_header h;
  char err[512];
  long long n = 0, ngas = 0;
  double u_init, molecular_weight;
  int i;

  CPU_Step[CPU_MISC] += measure_time();
  if(NTask > 1 || All.ICFormat < 1 || All.ICFormat > 3)
    gl_io_fail("read_ic", "the library reads formats 1, 2 and 3 on a single rank; link the reference's read_ic.o otherwise", 8801);
  if(All.ICFormat == 3)
    {
      /* read_ic.c:88: the single-file name of an HDF5 initial-conditions file */
      static char h5name[512];
      snprintf(h5name, sizeof(h5name), "%s.hdf5", fname);
      fname = h5name;
    }
  gl_fill_layout();
  gl_io_setup(&io, &par);
  par.format = All.ICFormat;
  NumPart = 0;
  N_gas = 0;
  if(gl_snapshot_read_header(fname, All.ICFormat, &h, err, sizeof(err)))
    gl_io_fail("read_ic", err, 123);
  /* read_file(): particle storage is sized from the header (read_ic.c:754-790) */
  All.TotN_gas = (h.num_files <= 1 ? (unsigned int) h.npart[0] : h.npartTotal[0]);
  for(i = 0, All.TotNumPart = 0; i < 6; i++)
    All.TotNumPart += (h.num_files <= 1 ? (unsigned int) h.npart[i] : h.npartTotal[i]);
  All.MaxPart = (int) (All.PartAllocFactor * (All.TotNumPart / NTask));
  All.MaxPartSph = (int) (All.PartAllocFactor * (All.TotN_gas / NTask));
#ifdef INHOMOG_GASDISTR_HINT
  All.MaxPartSph = All.MaxPart;
#endif
  allocate_memory();
  if(RestartFlag >= 2)
    All.Time = All.TimeBegin = h.time;
  /* read_ic.c:181-192 */
  u_init = (1.0 / GAMMA_MINUS1) * (BOLTZMANN / PROTONMASS) * All.InitGasTemp;
  u_init *= All.UnitMass_in_g / All.UnitEnergy_in_cgs;
  if(All.InitGasTemp > 1.0e4)
    molecular_weight = 4 / (8 - 5 * (1 - HYDROGEN_MASSFRAC));
  else
    molecular_weight = 4 / (1 + 3 * HYDROGEN_MASSFRAC);
  u_init /= molecular_weight;
  All.InitGasU = u_init;
  par.InitGasU = All.InitGasU;
  if(gl_snapshot_read(fname, &GL_Layout, &io, &par, P, SphP, All.MaxPart, All.MaxPartSph, &n, &ngas, &h, err, sizeof(err)))
    gl_io_fail("read_ic", err, 890);
  NumPart = (int) n;
  N_gas = (int) ngas;
  for(i = 0; i < 6; i++)
    All.MassTable[i] = par.MassTable[i];
  memcpy(&header, &h, sizeof(header));
  if(ThisTask == 0)
    {
      printf("reading done.\n");
      printf("Total number of particles :  %d%09d\n\n", (int) (All.TotNumPart / 1000000000), (int) (All.TotNumPart % 1000000000));
      fflush(stdout);
    }
  CPU_Step[CPU_SNAPSHOT] += measure_time();
}

void savepositions(int num)
{
  gl_io_layout io;
  gl_io_params par;
  char err[512], buf[500];
  double *nh = NULL, *sfr = NULL;
  int i;

  if(ThisTask == 0)
    printf("\nwriting snapshot file... \n");
  CPU_Step[CPU_MISC] += measure_time();
#if defined(SFR) || defined(BLACK_HOLES)
  rearrange_particle_sequence();
  /* ensures that new tree will be constructed (io.c:47-49) */
  All.NumForcesSinceLastDomainDecomp = (long long) (1 + All.TreeDomainUpdateFrequency * All.TotNumPart);
#endif
  if(NTask > 1 || All.NumFilesPerSnapshot > 1 || All.SnapFormat < 1 || All.SnapFormat > 3)
    gl_io_fail("savepositions", "the library writes formats 1, 2 and 3, one file, on a single rank; link the reference's io.o otherwise", 8802);
  gl_fill_layout();
  gl_io_setup(&io, &par);
  par.format = All.SnapFormat;
#ifdef COOLING
  if(All.CoolingOn && N_gas > 0)
    {
      double ne, nh0, nHeII, dmax1, dmax2;
      nh = (double *) malloc(sizeof(double) * N_gas);
      for(i = 0; i < N_gas; i++)
	{
	  /* io.c:349-360 (non-comoving: a3inv = 1) */
	  ne = SphP[i].Ne;
	  AbundanceRatios(DMAX(All.MinEgySpec, SphP[i].Entropy / GAMMA_MINUS1 * pow(SphP[i].d.Density, GAMMA_MINUS1)),
			  SphP[i].d.Density, &ne, &nh0, &nHeII);
	  nh[i] = nh0;
	}
    }
#endif
#ifdef SFR
  if(All.StarformationOn && N_gas > 0)
    {
      sfr = (double *) malloc(sizeof(double) * N_gas);
      for(i = 0; i < N_gas; i++)
	sfr[i] = get_starformation_rate(i);	/* io.c:469-477 */
    }
#endif
  sprintf(buf, "%s%s_%03d", All.OutputDir, All.SnapshotFileBase, num);
  i = gl_snapshot_write(buf, &GL_Layout, &io, &par, P, SphP, NumPart, nh, sfr, err, sizeof(err));
  free(nh);
  free(sfr);
  if(i)
    gl_io_fail("savepositions", err, 123);
  if(ThisTask == 0)
    printf("done with snapshot.\n");
  CPU_Step[CPU_SNAPSHOT] += measure_time();
}
#endif /* GL_SNAPSHOT_IO */

#if defined(GL_TIMESTEP) || defined(GL_RESIDENT_RUN)
/* All.* -> gl_timestep_params (timestep.c:607-1122 reads these) */
static void gl_fill_timestep_params(gl_timestep_params * t)
{
  int i;
  t->Ti_Current = All.Ti_Current;
  for(i = 0; i < 32; i++)
    t->TimeBinActive[i] = i < TIMEBINS ? TimeBinActive[i] : 0;
  t->Timebase_interval = All.Timebase_interval;
  t->ErrTolIntAccuracy = All.ErrTolIntAccuracy;
  t->CourantFac = All.CourantFac;
  t->MaxSizeTimestep = All.MaxSizeTimestep;
  t->MinSizeTimestep = All.MinSizeTimestep;
  t->MinEgySpec = All.MinEgySpec;
  t->GammaMinus1 = GAMMA_MINUS1;
  for(i = 0; i < 6; i++)
    t->SofteningTable[i] = All.SofteningTable[i];
  t->OuterBoundary = All.OuterBoundary;
  t->UnitVelocity_in_cm_per_s = All.UnitVelocity_in_cm_per_s;
  t->FeedBackVelocity = All.FeedBackVelocity;
  t->SMBHmass = All.SMBHmass;
  t->InnerBoundary = All.InnerBoundary;
  t->SinkBoundary = All.SinkBoundary;
  t->AccDtBlackHole = All.AccDtBlackHole;
#ifdef DUST
  t->p_DeltaDustMomentum = offsetof(struct particle_data, DeltaDustMomentum);
#else
  t->p_DeltaDustMomentum = -1;
#endif
#ifdef VIRTUAL
  t->p_NewDensity = offsetof(struct particle_data, NewDensity);
#else
  t->p_NewDensity = -1;
#endif
}
#endif

#ifdef GL_RESIDENT_RUN
/* ---- run.c:63-110 for nsteps iterations with the particle state resident on the GPU ----------------------------------
 * run()'s loop body is  find_next_sync_point_and_drift(); domain_Decomposition(); compute_accelerations(0);
 * advance_and_find_timesteps();  on P[] / SphP[] in host memory.  A maintainer who does not need the structs between
 * two outputs replaces `nsteps` iterations by ONE call of this function (between outputs: nsteps = until
 * All.Ti_nextoutput; the statistics / restart-file hooks of run.c:66-110 then run once per call): the structs are read
 * once, the loop runs in the library (gl_run_steps: drift, tree + gravity + density + hydro for the active particles,
 * kick), the structs are written once, and the time line of the host is brought up to date - All.Ti_Current, All.Time,
 * TimeBinActive[], the time-bin lists (reconstruct_timebins(), the reference's own function for exactly this).
 * Restrictions: include/gl_b200.h "run()'s loop".  Returns the iterations executed. */
int gl_glue_run_steps(int nsteps)
{
  gl_timestep_params t;
  gl_run_info info;
  int i;
  if(All.ComovingIntegrationOn)
    gl_fail("gl_glue_run_steps: comoving integration");
  if(NTask != 1)
    gl_fail("gl_glue_run_steps: the struct-level entry point handles NTask == 1 (at NTask > 1 keep the state in the library: gl_run_steps)");
  gl_attach();
  gl_sync_params();
  /* everybody to All.Ti_Current first (the loop drifts all particles from ONE time) */
  if(gl_gadget_drift(&GL_Layout, P, SphP, NumPart, N_gas, All.Ti_Current, All.Timebase_interval, All.MinGasHsml, NULL))
    endrun(12);
  gl_fill_timestep_params(&t);
  if(gl_gadget_run_steps(GL, P, SphP, NumPart, N_gas, &t, nsteps, &info))
    gl_fail("gl_gadget_run_steps");
  All.Ti_Current = t.Ti_Current;
  All.Time = All.TimeBegin + All.Ti_Current * All.Timebase_interval;	/* run.c:262-265 */
  for(i = 0; i < TIMEBINS; i++)
    TimeBinActive[i] = t.TimeBinActive[i];
  All.NumCurrentTiStep += info.steps_done;
  if(All.TypeOfOpeningCriterion == 1 && info.steps_done > 0)
    All.ErrTolTheta = 0;	/* gravtree.c:396-397 */
  /* reconstruct_timebins() checks its active list against NumForceUpdate (predict.c:104-123): the particles of the
   * last step's active bins are still on active bins after their kick (a bin only grows onto a synchronised bin) */
  NumForceUpdate = (int) info.active_last;
  reconstruct_timebins();
  TreeReconstructFlag = 1;
  gl_loaded_for = -1;
  return info.steps_done;
}
#endif

#ifdef GL_TIMESTEP
/* ---- advance_and_find_timesteps() (proto.h; timestep.c:29) --------------------------------------------------------
 * The per-particle work - get_timestep, the bin decision, do_the_kick - runs in the library on all host cores
 * (gl_gadget_advance_and_find_timesteps); what stays here is the reference's serial bookkeeping of the time-bin lists
 * (timestep.c:170-233), executed for the particles whose bin changed, in active-list order, with the reference's own
 * statements.  Non-comoving integration without PMGRID / MAKEGLASS (the shipped build); the tree is rebuilt after the
 * kicks, so no dynamic tree update is prepared (All.DoDynamicUpdate = 0). */
void advance_and_find_timesteps(void)
{
  gl_timestep_params t;
  char err[512];
  int i, k, n, bin, binold, prev, next, *act, *old;

  CPU_Step[CPU_MISC] += measure_time();
  if(All.ComovingIntegrationOn)
    gl_fail("advance_and_find_timesteps: comoving integration");
  gl_fill_layout();
  All.DoDynamicUpdate = 0;
  gl_fill_timestep_params(&t);
  for(i = FirstActiveParticle, n = 0; i >= 0; i = NextActiveParticle[i]) n++;
  act = (int *) malloc(sizeof(int) * (n > 0 ? n : 1));	/* private scratch: the mymalloc stack stays as found */
  old = (int *) malloc(sizeof(int) * (n > 0 ? n : 1));
  for(i = FirstActiveParticle, n = 0; i >= 0; i = NextActiveParticle[i]) act[n++] = i;
  if(gl_gadget_advance_and_find_timesteps(&GL_Layout, &t, P, SphP, act, n, old, err, sizeof(err)))
    {
      printf("%s\n", err);
      fflush(stdout);
      endrun(888);
    }
  for(k = 0; k < n; k++)
    {
      i = act[k];
      binold = old[k];
      bin = P[i].TimeBin;
      if(bin == binold) continue;
      /* timestep.c:170-233 */
      TimeBinCount[binold]--;
      if(P[i].Type == 0)
	{
	  TimeBinCountSph[binold]--;
#ifdef SFR
	  TimeBinSfr[binold] -= SphP[i].Sfr;
	  TimeBinSfr[bin] += SphP[i].Sfr;
#endif
	}
#ifdef BLACK_HOLES
      if(P[i].Type == 5)
	{
#ifndef DUST
	  TimeBin_BH_mass[binold] -= P[i].BH_Mass;
	  TimeBin_BH_dynamicalmass[binold] -= P[i].Mass;
	  TimeBin_BH_Mdot[binold] -= P[i].BH_Mdot;
	  TimeBin_BH_mass[bin] += P[i].BH_Mass;
	  TimeBin_BH_dynamicalmass[bin] += P[i].Mass;
	  TimeBin_BH_Mdot[bin] += P[i].BH_Mdot;
#else
	  TimeBin_BH_mass[binold] -= P[i].Dust_Mass;
	  TimeBin_BH_dynamicalmass[binold] -= P[i].Total_Mass;
	  TimeBin_BH_Mdot[binold] -= P[i].Mass;
	  TimeBin_BH_mass[bin] += P[i].Dust_Mass;
	  TimeBin_BH_dynamicalmass[bin] += P[i].Total_Mass;
	  TimeBin_BH_Mdot[bin] += P[i].Mass;
#endif
	}
#endif
      prev = PrevInTimeBin[i];
      next = NextInTimeBin[i];
      if(FirstInTimeBin[binold] == i) FirstInTimeBin[binold] = next;
      if(LastInTimeBin[binold] == i) LastInTimeBin[binold] = prev;
      if(prev >= 0) NextInTimeBin[prev] = next;
      if(next >= 0) PrevInTimeBin[next] = prev;
      if(TimeBinCount[bin] > 0)
	{
	  PrevInTimeBin[i] = LastInTimeBin[bin];
	  NextInTimeBin[LastInTimeBin[bin]] = i;
	  NextInTimeBin[i] = -1;
	  LastInTimeBin[bin] = i;
	}
      else
	{
	  FirstInTimeBin[bin] = LastInTimeBin[bin] = i;
	  PrevInTimeBin[i] = NextInTimeBin[i] = -1;
	}
      TimeBinCount[bin]++;
      if(P[i].Type == 0) TimeBinCountSph[bin]++;
    }
  free(act);
  free(old);
  CPU_Step[CPU_TIMELINE] += measure_time();
}

/* The one loop of find_next_sync_point_and_drift() that scales with the particle number (run.c:325-329):
 *     for(i = FirstActiveParticle, NumForceUpdate = 0; i >= 0; i = NextActiveParticle[i])
 *       { drift_particle(i, All.Ti_Current); NumForceUpdate++; }
 * A maintainer replaces it by   NumForceUpdate = gl_drift_active_particles();   (the function itself is all
 * bookkeeping of the time line and stays the reference's). */
int gl_drift_active_particles(void)
{
  int i, n, *act;
  for(i = FirstActiveParticle, n = 0; i >= 0; i = NextActiveParticle[i]) n++;
  act = (int *) malloc(sizeof(int) * (n > 0 ? n : 1));
  for(i = FirstActiveParticle, n = 0; i >= 0; i = NextActiveParticle[i]) act[n++] = i;
  gl_fill_layout();
  if(gl_gadget_drift_list(&GL_Layout, P, SphP, act, n, N_gas, All.Ti_Current, All.Timebase_interval, All.MinGasHsml, NULL))
    endrun(12);
  free(act);
  return n;
}
#endif /* GL_TIMESTEP */

#ifdef GL_PEANO_ORDER
/* ---- peano_hilbert_order() (proto.h; peano.c:25) ------------------------------------------------------------------
 * Called by domain_Decomposition() after the particle exchange (domain.c:132) with the keys in the global Key[]
 * (domain.c:1750).  The reference sorts 16-byte records with a serial merge sort and then follows permutation cycles
 * through 536-byte structs one at a time; the library does both on all host cores.  The mymalloc stack is left as
 * found (the library uses malloc). */
void peano_hilbert_order(void)
{
  if(ThisTask == 0)
    printf("begin Peano-Hilbert order...\n");
  gl_fill_layout();
  if(gl_gadget_peano_hilbert_order(&GL_Layout, P, SphP, NumPart, N_gas, Key, NULL))
    {
      printf("libgl_b200: gl_gadget_peano_hilbert_order failed\n");
      fflush(stdout);
      endrun(8803);
    }
  if(ThisTask == 0)
    printf("Peano-Hilbert done.\n");
}
#endif /* GL_PEANO_ORDER */

#ifdef GL_DOMAIN
/* ---- domain_Decomposition() (proto.h; domain.c:100) for a run on ONE rank (one GPU) ----------------------------------
 * With a single rank nothing is exchanged; what the reference's function leaves behind is the cube (DomainCorner / Center /
 * Len / Fac), the top-level tree TopNodes[] with NTopnodes / NTopleaves, DomainTask[] / DomainStartList[] /
 * DomainEndList[], P[] and SphP[] in Peano-Hilbert order, a freshly allocated tree and rebuilt time-bin lists.  All of
 * that is reproduced here with the heavy loops - drift, extent, keys, key sort, top-level tree, struct permutation - in
 * the library on all host cores (the reference spends 0.65 s per call on 4e5 particles, most of it in reorder_gas and in
 * two serial merge sorts).  The mymalloc stack is left exactly as the reference leaves it: DomainStartList,
 * DomainEndList, TopNodes (shrunk), then the tree. */
void domain_Decomposition(void)
{
  int i, nn = 0, nl = 0, rc;
  double *cost, dlen, dfac, tm[8];
  gl_topnode *nodes;

  if(NTask != 1)
    gl_fail("domain_Decomposition: the library's version is for one rank; link the reference's domain.o otherwise");
  if(!(All.NumForcesSinceLastDomainDecomp >= All.TotNumPart * All.TreeDomainUpdateFrequency || All.DoDynamicUpdate == 0))
    return;
  CPU_Step[CPU_MISC] += measure_time();
  gl_fill_layout();
  tm[0] = second();
  if(gl_gadget_drift(&GL_Layout, P, SphP, NumPart, N_gas, All.Ti_Current, All.Timebase_interval, All.MinGasHsml, NULL))
    endrun(12);			/* domain.c:120-122 */
  force_treefree();
  domain_free();
#if defined(SFR) || defined(BLACK_HOLES) || defined(DUST)
  rearrange_particle_sequence();
#endif
  All.NumForcesSinceLastDomainDecomp = 0;
  TreeReconstructFlag = 1;	/* ensures that new tree will be constructed */
  if(ThisTask == 0)
    printf("domain decomposition... (presently allocated=%g MB)\n", AllocatedBytes / (1024.0 * 1024.0));
  cost = (double *) malloc(sizeof(double) * (NumPart > 0 ? NumPart : 1));
  tm[1] = second();
  do
    {
      domain_allocate();
      Key = (peanokey *) mymalloc_msg(sizeof(peanokey) * All.MaxPart, "domain_key");
      /* domain_decompose(), domain.c:404-427 */
      gl_gadget_costfactors(&GL_Layout, P, NumPart, cost);
      for(i = 0; i < 6; i++)
	NtypeLocal[i] = 0;
      for(i = 0; i < NumPart; i++)
	{
	  NtypeLocal[P[i].Type]++;
	  All.Cadj_Cost += cost[i];
	}
      for(i = 0; i < 6; i++)
	Ntype[i] = NtypeLocal[i];
      All.Cadj_Cpu *= 0.9;
      All.Cadj_Cost *= 0.9;
      if(gl_gadget_domain_keys(&GL_Layout, P, NumPart, DomainCorner, DomainCenter, &dlen, &dfac, (gl_peanokey *) Key))
	gl_fail("gl_gadget_domain_keys");
      DomainLen = dlen;
      DomainFac = dfac;
      tm[2] = second();
      nodes = (gl_topnode *) malloc(sizeof(gl_topnode) * MaxTopNodes);
      rc = gl_domain_top_tree((gl_peanokey *) Key, cost, NumPart, NTask, MULTIPLEDOMAINS, TOPNODEFACTOR, nodes, MaxTopNodes, &nn, &nl);
      tm[3] = second();
      if(rc == 0)
	{
	  NTopnodes = nn;
	  NTopleaves = nl;
	  if(NTopleaves < MULTIPLEDOMAINS * NTask)
	    endrun(112);
	  for(i = 0; i < NTopnodes; i++)
	    {
	      TopNodes[i].StartKey = nodes[i].StartKey;
	      TopNodes[i].Size = nodes[i].Size;
	      TopNodes[i].Daughter = nodes[i].Daughter;
	      TopNodes[i].Leaf = nodes[i].Leaf;
	    }
	  if(gl_domain_split(nodes, NTopleaves, (gl_peanokey *) Key, cost, NumPart, MULTIPLEDOMAINS * NTask, DomainStartList, DomainEndList,
			     NULL, NULL))
	    endrun(112);
	  for(i = 0; i < NTopleaves; i++)
	    DomainTask[i] = 0;	/* domain_assign_load_or_work_balanced on one rank */
	}
      free(nodes);
      if(rc == 1)
	{
	  /* out of top nodes: domain.c:246-269 */
	  myfree(Key);
	  domain_free();
	  if(ThisTask == 0)
	    printf("Increasing TopNodeAllocFactor=%g  ", All.TopNodeAllocFactor);
	  All.TopNodeAllocFactor *= 1.3;
	  if(ThisTask == 0)
	    printf("new value=%g\n", All.TopNodeAllocFactor);
	  if(All.TopNodeAllocFactor > 1000)
	    endrun(781);
	}
      else if(rc)
	gl_fail("gl_domain_top_tree");
    }
  while(rc);
  free(cost);
  if(ThisTask == 0)
    printf("domain decomposition done. \n");
  CPU_Step[CPU_DOMAIN] += measure_time();
  tm[4] = second();
  peano_hilbert_order();
  tm[5] = second();
  CPU_Step[CPU_PEANO] += measure_time();
  myfree(Key);
  memmove(TopNodes + NTopnodes, DomainTask, NTopnodes * sizeof(int));
  TopNodes = (struct topnode_data *) myrealloc(TopNodes, NTopnodes * sizeof(struct topnode_data) + NTopnodes * sizeof(int));
  DomainTask = (int *) (TopNodes + NTopnodes);
  force_treeallocate((int) (All.TreeAllocFactor * All.MaxPart) + NTopnodes, All.MaxPart);
  tm[6] = second();
  reconstruct_timebins();
  tm[7] = second();
  if(getenv("GL_DEBUG_E2E"))
    fprintf(stderr, "domain_Decomposition: drift+free %.1f ms, costs+keys %.1f, top tree %.1f, split %.1f, peano order %.1f, "
	    "tree allocate %.1f, time bins %.1f\n", (tm[1] - tm[0]) * 1e3, (tm[2] - tm[1]) * 1e3, (tm[3] - tm[2]) * 1e3,
	    (tm[4] - tm[3]) * 1e3, (tm[5] - tm[4]) * 1e3, (tm[6] - tm[5]) * 1e3, (tm[7] - tm[6]) * 1e3);
}
#endif /* GL_DOMAIN */
