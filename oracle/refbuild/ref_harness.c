/* ref_harness.c - drives the UNMODIFIED reference hot path (gravity_tree, density,
 * force_update_hmax, hydro_force, domain_Decomposition, force_treebuild) on a
 * particle set read from a file and dumps every output the parity tests compare.
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled together with the reference's own sources
 * where they lie under /root/reference (see Makefile in this directory); the
 * binary goes to oracle/_ref/ and is never linked into, or called by, the product.
 *
 * The set-up sequence follows what the reference does at start-up
 * (begrun.c:32-, init.c:36-, allocate.c:20, init.c:150-157, init.c:250-252) reduced
 * to the fields the hot path reads.
 *
 * usage: ref_harness <in.glin> <out.glout> "<ops>"
 *   ops is a blank-separated list, executed in order:
 *     T  ngb_treebuild()                      (init.c:529)
 *     S  initial Hsml guess                   (restates init.c:719-745)
 *     D  density()                            (accel.c:84)
 *     R  TreeReconstructFlag = 1              (what domain_Decomposition sets, domain.c:135)
 *     G  gravity_tree()                       (accel.c:61)
 *     U  force_update_hmax()                  (accel.c:96)
 *     H  hydro_force()                        (accel.c:106)
 *     N  dump the node table of the current tree
 *     C  integer pair counts at the current Hsml via ngb_treefind_variable / _pairs
 *     P  compute_potential()                  (potential.c:22; run.c:74 at statistics times)
 *     K  Peano-Hilbert and Morton key of every particle (peano_and_morton_key, peano.c:335)
 *     A  compute_accelerations-order block:  R G D U H   (accel.c:61-106 without sub-grid physics)
 *     L  eight iterations of run()'s loop: find_next_sync_point_and_drift, domain_Decomposition, force path, kick
 *     M  op L with its eight iterations run by the glue's gl_glue_run_steps() (drop-in builds: the device-resident loop)
 *     B  blackhole_accretion()               (accel.c:189, blackhole.c:70: sink accretion loops blackhole_evaluate :794 and
 *        blackhole_evaluate_swallow :1201 over ngb_treefind_blackhole :1351) on the current tree; sink state is derived
 *        from the input (see set_subgrid_state), everything read and written is dumped (pre / post)
 *     Y  dust_density()                       (accel.c:192, dust.c:60 -> dust_evaluate_density :748)
 *     Z  dust_drag()                          (accel.c:196, dust.c:263 -> dust_evaluate_select :889)
 *     V  advance_and_find_timesteps()         (timestep.c:29) at a synchronisation point of an individual-time-step
 *        state built from the input's time bins; dumps everything it reads and everything it leaves behind
 *   every op that computes dumps its outputs under the prefix "<k>_<op>_" with k the
 *   position in the list.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <mpi.h>

#include "allvars.h"
#include "proto.h"

/* ---- symbols the 13 hot-path files leave undefined ---- */
double get_drift_factor(int t0, int t1) { (void) t0; (void) t1; return 0; }
double get_gravkick_factor(int t0, int t1) { (void) t0; (void) t1; return 0; }
double get_hydrokick_factor(int t0, int t1) { (void) t0; (void) t1; return 0; }
double hubble_function(double a) { (void) a; return 0; }
void rearrange_particle_sequence(void) { /* no Mass==0 particles in harness inputs (sfr_eff.c:1018) */ }
static unsigned long long lcg_state = 42;
double gsl_rng_uniform(gsl_rng * r)
{
  (void) r;
  lcg_state = lcg_state * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double) (lcg_state >> 11) / 9007199254740992.0;
}

/* ---- tiny tagged-array container: see oracle/glio.py for the reader ---- */
static FILE *fout;
static void put(const char *prefix, const char *name, char kind, long long n, int ncomp, const void *data)
{
  char tag[48];
  int esz = (kind == 'd' || kind == 'q') ? 8 : 4;
  memset(tag, 0, sizeof(tag));
  snprintf(tag, sizeof(tag), "%s%s", prefix, name);
  fwrite(tag, 1, 48, fout);
  fwrite(&kind, 1, 1, fout);
  fwrite(&ncomp, 4, 1, fout);
  fwrite(&n, 8, 1, fout);
  fwrite(data, esz, (size_t) n * ncomp, fout);
}
static double now(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

#define NPAR 32
static double par[NPAR];
enum { P_THETA, P_ERRTOLFORCEACC, P_OPENCRIT, P_DESNUMNGB, P_MAXNGBDEV, P_ARTVISC, P_G, P_SOFT0,
  P_MINGASHSMLFRAC = P_SOFT0 + 6, P_TIMEBASE_INTERVAL, P_TI_CURRENT, P_BHNGBFACTOR, P_PARTALLOC, P_MINEGYSPEC,
  P_BHFEEDBACK, P_SMBHMASS, P_INNERBOUNDARY, P_SINKBOUNDARY, P_VISCOUSTIMESCALE, P_MEANWEIGHT
};

static void dump_particles_out(const char *prefix, int what)
{
  long long n = NumPart, i;
  double *d3 = malloc(sizeof(double) * 3 * n), *d1 = malloc(sizeof(double) * n);
  int *i1 = malloc(sizeof(int) * n);
  for(i = 0; i < n; i++) i1[i] = (int) P[i].ID;
  put(prefix, "id", 'i', n, 1, i1);
  if(what == 'G')
    {
      for(i = 0; i < n; i++) { d3[3 * i] = P[i].g.GravAccel[0]; d3[3 * i + 1] = P[i].g.GravAccel[1]; d3[3 * i + 2] = P[i].g.GravAccel[2]; }
      put(prefix, "GravAccel", 'd', n, 3, d3);
      for(i = 0; i < n; i++) d1[i] = P[i].OldAcc;
      put(prefix, "OldAcc", 'd', n, 1, d1);
      for(i = 0; i < n; i++) d1[i] = P[i].GravCost;
      put(prefix, "GravCost", 'd', n, 1, d1);
    }
  if(what == 'D' || what == 'S')
    {
      for(i = 0; i < n; i++) d1[i] = P[i].Hsml;
      put(prefix, "Hsml", 'd', n, 1, d1);
    }
  if(what == 'D')
    {
      long long ng = N_gas;
      for(i = 0; i < n; i++) d1[i] = P[i].n.NumNgb;
      put(prefix, "NumNgb", 'd', n, 1, d1);
      for(i = 0; i < ng; i++) d1[i] = SphP[i].d.Density;
      put(prefix, "Density", 'd', ng, 1, d1);
      for(i = 0; i < ng; i++) d1[i] = SphP[i].h.DhsmlDensityFactor;
      put(prefix, "DhsmlDensityFactor", 'd', ng, 1, d1);
      for(i = 0; i < ng; i++) d1[i] = SphP[i].v.DivVel;
      put(prefix, "DivVel", 'd', ng, 1, d1);
      for(i = 0; i < ng; i++) d1[i] = SphP[i].r.CurlVel;
      put(prefix, "CurlVel", 'd', ng, 1, d1);
      for(i = 0; i < ng; i++) d1[i] = SphP[i].Pressure;
      put(prefix, "Pressure", 'd', ng, 1, d1);
      /* sink (Type 5) / dust (Type 2) targets: density.c:944-988 */
      for(i = 0; i < n; i++) d1[i] = (P[i].Type == 5) ? P[i].b1.BH_Density : (P[i].Type == 2 ? P[i].d1.DUST_Density : 0);
      put(prefix, "SinkDensity", 'd', n, 1, d1);
      for(i = 0; i < n; i++) d1[i] = (P[i].Type == 5) ? P[i].b2.BH_Entropy : (P[i].Type == 2 ? P[i].d2.DUST_Entropy : 0);
      put(prefix, "SinkEntropy", 'd', n, 1, d1);
      for(i = 0; i < n; i++)
	{
	  int k;
	  for(k = 0; k < 3; k++)
	    d3[3 * i + k] = (P[i].Type == 5) ? P[i].b3.BH_SurroundingGasVel[k] : (P[i].Type == 2 ? P[i].d3.DUST_SurroundingGasVel[k] : 0);
	}
      put(prefix, "SinkGasVel", 'd', n, 3, d3);
    }
  if(what == 'H')
    {
      long long ng = N_gas;
      for(i = 0; i < ng; i++) { d3[3 * i] = SphP[i].a.HydroAccel[0]; d3[3 * i + 1] = SphP[i].a.HydroAccel[1]; d3[3 * i + 2] = SphP[i].a.HydroAccel[2]; }
      put(prefix, "HydroAccel", 'd', ng, 3, d3);
      for(i = 0; i < ng; i++) d1[i] = SphP[i].e.DtEntropy;
      put(prefix, "DtEntropy", 'd', ng, 1, d1);
      for(i = 0; i < ng; i++) d1[i] = SphP[i].MaxSignalVel;
      put(prefix, "MaxSignalVel", 'd', ng, 1, d1);
    }
  free(d3); free(d1); free(i1);
}

/* everything drift_particle (predict.c:129) reads or writes, in P[] order */
static void dump_drift_state(const char *prefix, const char *tag)
{
  long long n = NumPart, ng = N_gas, i;
  double *d3 = malloc(sizeof(double) * 3 * n), *d1 = malloc(sizeof(double) * n);
  int *i1 = malloc(sizeof(int) * n);
  char nm[40];
#define PUT3(field, expr, cnt) do { for(i = 0; i < (cnt); i++) { d3[3 * i] = expr[0]; d3[3 * i + 1] = expr[1]; d3[3 * i + 2] = expr[2]; } \
    snprintf(nm, sizeof(nm), "%s_%s", tag, field); put(prefix, nm, 'd', (cnt), 3, d3); } while(0)
#define PUT1(field, expr, cnt) do { for(i = 0; i < (cnt); i++) d1[i] = expr; snprintf(nm, sizeof(nm), "%s_%s", tag, field); put(prefix, nm, 'd', (cnt), 1, d1); } while(0)
#define PUTI(field, expr, cnt) do { for(i = 0; i < (cnt); i++) i1[i] = expr; snprintf(nm, sizeof(nm), "%s_%s", tag, field); put(prefix, nm, 'i', (cnt), 1, i1); } while(0)
  PUTI("id", (int) P[i].ID, n); PUTI("Type", P[i].Type, n); PUTI("TimeBin", P[i].TimeBin, n); PUTI("Ti_begstep", P[i].Ti_begstep, n);
  PUTI("Ti_current", P[i].Ti_current, n);
  PUT3("Pos", P[i].Pos, n); PUT3("Vel", P[i].Vel, n); PUT3("GravAccel", P[i].g.GravAccel, n);
  PUT1("Mass", P[i].Mass, n); PUT1("Hsml", P[i].Hsml, n);
  PUT3("VelPred", SphP[i].VelPred, ng); PUT3("HydroAccel", SphP[i].a.HydroAccel, ng); PUT3("DragAccel", SphP[i].da.DragAccel, ng);
  PUT1("DivVel", SphP[i].v.DivVel, ng); PUT1("Density", SphP[i].d.Density, ng); PUT1("Entropy", SphP[i].Entropy, ng);
  PUT1("DtEntropy", SphP[i].e.DtEntropy, ng); PUT1("Pressure", SphP[i].Pressure, ng);
#undef PUT3
#undef PUT1
#undef PUTI
  free(d3); free(d1); free(i1);
}

/* everything advance_and_find_timesteps (timestep.c:29 -> get_timestep :607, do_the_kick :364) reads or writes */
static void dump_kick_state(const char *prefix, const char *tag)
{
  long long n = NumPart, ng = N_gas, i;
  double *d3 = malloc(sizeof(double) * 3 * n), *d1 = malloc(sizeof(double) * n);
  int *i1 = malloc(sizeof(int) * n);
  char nm[40];
#define PUT3(field, expr, cnt) do { for(i = 0; i < (cnt); i++) { d3[3 * i] = expr[0]; d3[3 * i + 1] = expr[1]; d3[3 * i + 2] = expr[2]; } \
    snprintf(nm, sizeof(nm), "%s_%s", tag, field); put(prefix, nm, 'd', (cnt), 3, d3); } while(0)
#define PUT1(field, expr, cnt) do { for(i = 0; i < (cnt); i++) d1[i] = expr; snprintf(nm, sizeof(nm), "%s_%s", tag, field); put(prefix, nm, 'd', (cnt), 1, d1); } while(0)
#define PUTI(field, expr, cnt) do { for(i = 0; i < (cnt); i++) i1[i] = expr; snprintf(nm, sizeof(nm), "%s_%s", tag, field); put(prefix, nm, 'i', (cnt), 1, i1); } while(0)
  PUTI("id", (int) P[i].ID, n); PUTI("Type", P[i].Type, n); PUTI("TimeBin", P[i].TimeBin, n); PUTI("Ti_begstep", P[i].Ti_begstep, n);
  PUT3("Pos", P[i].Pos, n); PUT3("Vel", P[i].Vel, n); PUT3("GravAccel", P[i].g.GravAccel, n); PUT3("DeltaDustMomentum", P[i].DeltaDustMomentum, n);
  PUT1("Mass", P[i].Mass, n); PUT1("Hsml", P[i].Hsml, n); PUT1("NewDensity", P[i].NewDensity, n);
  PUT3("VelPred", SphP[i].VelPred, ng); PUT3("HydroAccel", SphP[i].a.HydroAccel, ng); PUT3("DragAccel", SphP[i].da.DragAccel, ng);
  PUT1("Density", SphP[i].d.Density, ng); PUT1("Entropy", SphP[i].Entropy, ng); PUT1("DtEntropy", SphP[i].e.DtEntropy, ng);
  PUT1("Pressure", SphP[i].Pressure, ng); PUT1("MaxSignalVel", SphP[i].MaxSignalVel, ng);
  PUTI("NextInTimeBin", NextInTimeBin[i], n); PUTI("PrevInTimeBin", PrevInTimeBin[i], n);
  PUTI("TimeBinCount", TimeBinCount[i], TIMEBINS); PUTI("TimeBinCountSph", TimeBinCountSph[i], TIMEBINS);
  PUTI("FirstInTimeBin", FirstInTimeBin[i], TIMEBINS); PUTI("LastInTimeBin", LastInTimeBin[i], TIMEBINS);
  PUTI("TimeBinActive", TimeBinActive[i], TIMEBINS);
#undef PUT3
#undef PUT1
#undef PUTI
  free(d3); free(d1); free(i1);
}


/* what a force evaluation leaves behind beyond dump_kick_state's fields: the rest of the state run()'s loop carries
   from one step to the next (drift: DivVel; hydro sources: DhsmlDensityFactor, CurlVel; relative criterion: OldAcc) */
static void dump_loop_extra(const char *prefix, const char *tag)
{
  long long n = NumPart, ng = N_gas, i;
  double *d1 = malloc(sizeof(double) * (n + 1));
  char nm[40];
  double q[8];
#define PUT1(field, expr, cnt) do { for(i = 0; i < (cnt); i++) d1[i] = expr; snprintf(nm, sizeof(nm), "%s_%s", tag, field); put(prefix, nm, 'd', (cnt), 1, d1); } while(0)
  PUT1("OldAcc", P[i].OldAcc, n); PUT1("NumNgb", P[i].n.NumNgb, n);
  PUT1("DivVel", SphP[i].v.DivVel, ng); PUT1("CurlVel", SphP[i].r.CurlVel, ng);
  PUT1("DhsmlDensityFactor", SphP[i].h.DhsmlDensityFactor, ng);
#undef PUT1
  q[0] = All.Ti_Current; q[1] = All.Timebase_interval; q[2] = All.ErrTolTheta; q[3] = All.ErrTolForceAcc; q[4] = All.MinGasHsml;
  q[5] = All.MinEgySpec; q[6] = All.MaxSizeTimestep; q[7] = All.G;
  snprintf(nm, sizeof(nm), "%s_%s", tag, "globals");
  put(prefix, nm, 'd', 8, 1, q);
  free(d1);
}


/* ---- sub-grid neighbour loops: blackhole.c, dust.c ---- */
/* run-time parameters of the shipped run (Pd_N16e5_*.txt:76,86-88,126,138,167-171) unless the input overrides them */
static int subgrid_ready;
static void set_subgrid_state(void)
{
  int i;
  if(subgrid_ready) return;
  subgrid_ready = 1;
  All.UnitLength_in_cm = 1.496e15; All.UnitMass_in_g = 1.989e33; All.UnitVelocity_in_cm_per_s = 297837.66;
  All.UnitTime_in_s = All.UnitLength_in_cm / All.UnitVelocity_in_cm_per_s;			/* begrun.c */
  All.UnitDensity_in_cgs = All.UnitMass_in_g / pow(All.UnitLength_in_cm, 3);
  All.UnitEnergy_in_cgs = All.UnitMass_in_g * pow(All.UnitLength_in_cm, 2) / pow(All.UnitTime_in_s, 2);
  All.BlackHoleFeedbackFactor = par[P_BHFEEDBACK];
  All.SMBHmass = par[P_SMBHMASS] > 0 ? par[P_SMBHMASS] : 0.5;
  All.InnerBoundary = par[P_INNERBOUNDARY] > 0 ? par[P_INNERBOUNDARY] : 2.0e-2;
  All.SinkBoundary = par[P_SINKBOUNDARY] > 0 ? par[P_SINKBOUNDARY] : 1.0e-4;
  All.ViscousTimescale = par[P_VISCOUSTIMESCALE] > 0 ? par[P_VISCOUSTIMESCALE] : 3.15e9;
  All.MeanWeight = par[P_MEANWEIGHT] > 0 ? par[P_MEANWEIGHT] : 2.45;
  All.MinEgySpec = par[P_MINEGYSPEC];
  All.xbh = All.ybh = All.zbh = 0;
  FdBlackHoles = fopen("/dev/null", "w");
  FdBlackHolesDetails = fopen("/dev/null", "w");
  /* the per-particle state no input array carries: a fixed function of what the input does carry */
  for(i = 0; i < NumPart; i++)
    {
      if(P[i].Type == 5)
	{
	  P[i].BH_Mass = 0.8 * P[i].Mass; P[i].AccDisc_Mass = 0.1 * P[i].Mass;
	  P[i].Dust_Mass = 0.01 * P[i].Mass; P[i].Total_Mass = P[i].Mass;
	}
      if(P[i].Type == 2)
	P[i].DustRadius = pow(10.0, (double) (P[i].ID % 7) - 3.0);	/* cm: both drag regimes of dust.c:340-366 */
    }
}
static void dump_subgrid_params(const char *prefix)
{
  double q[16];
  q[0] = All.SMBHmass; q[1] = All.InnerBoundary; q[2] = All.SinkBoundary; q[3] = All.SofteningBndry; q[4] = All.ViscousTimescale;
  q[5] = All.UnitTime_in_s; q[6] = All.UnitMass_in_g; q[7] = All.UnitEnergy_in_cgs; q[8] = All.BlackHoleFeedbackFactor;
  q[9] = All.Timebase_interval; q[10] = All.UnitLength_in_cm; q[11] = All.UnitDensity_in_cgs; q[12] = All.MeanWeight;
  q[13] = All.MinEgySpec; q[14] = All.UnitVelocity_in_cm_per_s; q[15] = All.Time;
  put(prefix, "params", 'd', 16, 1, q);
}
/* everything blackhole_accretion / dust_density / dust_drag read or write, in P[] order */
static void dump_subgrid_state(const char *prefix, const char *tag)
{
  long long n = NumPart, ng = N_gas, i;
  double *d3 = malloc(sizeof(double) * 3 * n), *d1 = malloc(sizeof(double) * n);
  int *i1 = malloc(sizeof(int) * n);
  char nm[40];
#define PUT3(field, expr, cnt) do { for(i = 0; i < (cnt); i++) { d3[3 * i] = expr[0]; d3[3 * i + 1] = expr[1]; d3[3 * i + 2] = expr[2]; } \
    snprintf(nm, sizeof(nm), "%s_%s", tag, field); put(prefix, nm, 'd', (cnt), 3, d3); } while(0)
#define PUT1(field, expr, cnt) do { for(i = 0; i < (cnt); i++) d1[i] = expr; snprintf(nm, sizeof(nm), "%s_%s", tag, field); put(prefix, nm, 'd', (cnt), 1, d1); } while(0)
#define PUTI(field, expr, cnt) do { for(i = 0; i < (cnt); i++) i1[i] = expr; snprintf(nm, sizeof(nm), "%s_%s", tag, field); put(prefix, nm, 'i', (cnt), 1, i1); } while(0)
  PUTI("id", (int) P[i].ID, n); PUTI("Type", P[i].Type, n); PUTI("TimeBin", P[i].TimeBin, n); PUTI("SwallowID", P[i].SwallowID, n);
  PUT3("Pos", P[i].Pos, n); PUT3("Vel", P[i].Vel, n); PUT3("GravAccel", P[i].g.GravAccel, n);
  PUT1("Mass", P[i].Mass, n); PUT1("Hsml", P[i].Hsml, n);
  PUT1("BH_Mass", P[i].BH_Mass, n); PUT1("BH_Mdot", P[i].BH_Mdot, n); PUT1("AccDisc_Mass", P[i].AccDisc_Mass, n);
  PUT1("Dust_Mass", P[i].Dust_Mass, n); PUT1("Total_Mass", P[i].Total_Mass, n);
  PUT1("SinkDensity", ((P[i].Type == 5) ? P[i].b1.BH_Density : (P[i].Type == 2 ? P[i].d1.DUST_Density : 0)), n);
  PUT1("SinkEntropy", ((P[i].Type == 5) ? P[i].b2.BH_Entropy : (P[i].Type == 2 ? P[i].d2.DUST_Entropy : 0)), n);
  for(i = 0; i < n; i++)
    {
      int k;
      for(k = 0; k < 3; k++)
	d3[3 * i + k] = (P[i].Type == 5) ? P[i].b3.BH_SurroundingGasVel[k] : (P[i].Type == 2 ? P[i].d3.DUST_SurroundingGasVel[k] : 0);
    }
  snprintf(nm, sizeof(nm), "%s_SinkGasVel", tag); put(prefix, nm, 'd', n, 3, d3);
  PUT1("DustParticleDensity", (P[i].Type == 2 ? P[i].d7.DUST_particle_density : 0), n);
  PUT1("DustRadius", P[i].DustRadius, n); PUT1("DustVcoll", P[i].DustVcoll, n);
  PUT3("DeltaDustMomentum", P[i].DeltaDustMomentum, n); PUT3("NewDragAcc", P[i].NewDragAcc, n);
  PUT1("DeltaDragEnergy", P[i].DeltaDragEnergy, n);
  PUT1("Entropy", SphP[i].Entropy, ng); PUT1("Density", SphP[i].d.Density, ng);
  PUT1("Injected_BH_Energy", SphP[i].i.Injected_BH_Energy, ng); PUT1("DragHeating", SphP[i].dh.DragHeating, ng);
  {
    int nact = 0, k = 0, *act;
    for(i = FirstActiveParticle; i >= 0; i = NextActiveParticle[i]) nact++;
    act = malloc(sizeof(int) * (nact + 1));
    for(i = FirstActiveParticle; i >= 0; i = NextActiveParticle[i]) act[k++] = (int) i;
    snprintf(nm, sizeof(nm), "%s_active", tag); put(prefix, nm, 'i', nact, 1, act);
    free(act);
  }
#undef PUT3
#undef PUT1
#undef PUTI
  free(d3); free(d1); free(i1);
}

/* accel.c:27 reduced to the force path (accel.c:61-106; the sub-grid physics behind it is outside the path): what
 * run.c's loop calls between the drift and the kick */
void compute_accelerations(int mode)
{
  (void) mode;
  gravity_tree();
  if(All.TotN_gas > 0)
    {
      density();
      force_update_hmax();
      hydro_force();
    }
}

void ref_density(void) __attribute__((weak));
int gl_glue_run_steps(int nsteps) __attribute__((weak));	/* integration/gl_gadget_glue.c -DGL_RESIDENT_RUN (drop-in builds) */
void ref_hydro_force(void) __attribute__((weak));

static void dump_nodes(const char *prefix)
{
  long long nn = Numnodestree, k;
  double *buf = malloc(sizeof(double) * 11 * nn);
  int *fl = malloc(sizeof(int) * nn);
  double dom[8];
  for(k = 0; k < nn; k++)
    {
      int no = All.MaxPart + (int) k;
      buf[11 * k + 0] = Nodes[no].len;
      buf[11 * k + 1] = Nodes[no].center[0];
      buf[11 * k + 2] = Nodes[no].center[1];
      buf[11 * k + 3] = Nodes[no].center[2];
      buf[11 * k + 4] = Nodes[no].u.d.mass;
      buf[11 * k + 5] = Nodes[no].u.d.s[0];
      buf[11 * k + 6] = Nodes[no].u.d.s[1];
      buf[11 * k + 7] = Nodes[no].u.d.s[2];
      buf[11 * k + 8] = Nodes[no].maxsoft;
      buf[11 * k + 9] = Extnodes[no].hmax;
      buf[11 * k + 10] = Extnodes[no].vmax;
      fl[k] = Nodes[no].u.d.bitflags;
    }
  put(prefix, "nodes", 'd', nn, 11, buf);
  put(prefix, "nodeflags", 'i', nn, 1, fl);
  /* the velocity moments of extNODE (forcetree.c:684-706): what the dynamic tree update reads */
  for(k = 0; k < nn; k++)
    {
      int no = All.MaxPart + (int) k;
      buf[4 * k + 0] = Extnodes[no].vs[0];
      buf[4 * k + 1] = Extnodes[no].vs[1];
      buf[4 * k + 2] = Extnodes[no].vs[2];
      buf[4 * k + 3] = Extnodes[no].divVmax;
    }
  put(prefix, "nodes_vel", 'd', nn, 4, buf);
  dom[0] = DomainLen; dom[1] = DomainFac;
  dom[2] = DomainCorner[0]; dom[3] = DomainCorner[1]; dom[4] = DomainCorner[2];
  dom[5] = DomainCenter[0]; dom[6] = DomainCenter[1]; dom[7] = DomainCenter[2];
  put(prefix, "domain", 'd', 8, 1, dom);
  free(buf); free(fl);
}

/* integer pair counts the reference never stores (density.c:858 is a dead local): re-evaluate the
 * caller-side tests of density.c:835-856 and hydra.c:1236-1266 on the lists the reference's own
 * neighbour searches return. */
static void pair_counts(const char *prefix)
{
  long long n = NumPart, i;
  int *cd = calloc(n, sizeof(int)), *ch = calloc(n, sizeof(int));
  int dummy = 0, dummy2 = 0;
  Ngblist = (int *) mymalloc(NumPart * sizeof(int));
  for(i = 0; i < n; i++)
    {
      int startnode, nn, k;
      double h = P[i].Hsml, h2 = h * h;
      if(!(P[i].Type == 0 || P[i].Type == 5 || P[i].Type == 2))
	continue;
      startnode = All.MaxPart;
      while(startnode >= 0)
	{
	  nn = ngb_treefind_variable(P[i].Pos, h, (int) i, &startnode, 0, &dummy, &dummy2);
	  for(k = 0; k < nn; k++)
	    {
	      int j = Ngblist[k];
	      double dx, dy, dz, r2;
	      if(P[j].Mass == 0) continue;
	      dx = P[i].Pos[0] - P[j].Pos[0]; dy = P[i].Pos[1] - P[j].Pos[1]; dz = P[i].Pos[2] - P[j].Pos[2];
	      r2 = dx * dx + dy * dy + dz * dz;
	      if(r2 < h2) cd[i]++;
	    }
	}
      if(P[i].Type != 0)
	continue;
      startnode = All.MaxPart;
      while(startnode >= 0)
	{
	  nn = ngb_treefind_pairs(P[i].Pos, h, (int) i, &startnode, 0, &dummy, &dummy2);
	  for(k = 0; k < nn; k++)
	    {
	      int j = Ngblist[k];
	      double dx, dy, dz, r2, hj;
	      if(P[j].Mass == 0) continue;
	      dx = P[i].Pos[0] - P[j].Pos[0]; dy = P[i].Pos[1] - P[j].Pos[1]; dz = P[i].Pos[2] - P[j].Pos[2];
	      r2 = dx * dx + dy * dy + dz * dz;
	      hj = P[j].Hsml;
	      if((r2 < h2 || r2 < hj * hj) && sqrt(r2) > 0) ch[i]++;
	    }
	}
    }
  myfree(Ngblist);
  {
    int *ids = malloc(sizeof(int) * n);
    for(i = 0; i < n; i++) ids[i] = (int) P[i].ID;
    put(prefix, "id", 'i', n, 1, ids);
    free(ids);
  }
  put(prefix, "DensPairs", 'i', n, 1, cd);
  put(prefix, "HydroPairs", 'i', n, 1, ch);
  free(cd); free(ch);
}

/* keys of every particle from the reference's own peano_and_morton_key (peano.c:335) */
static void dump_keys(const char *prefix)
{
  long long n = NumPart, i;
  long long *ph = malloc(8 * n), *mo = malloc(8 * n);
  int *ids = malloc(4 * n);
  for(i = 0; i < n; i++)
    {
      peanokey m;
      ph[i] = (long long) peano_and_morton_key((int) ((P[i].Pos[0] - DomainCorner[0]) * DomainFac),
						(int) ((P[i].Pos[1] - DomainCorner[1]) * DomainFac),
						(int) ((P[i].Pos[2] - DomainCorner[2]) * DomainFac), BITS_PER_DIMENSION, &m);
      mo[i] = (long long) m;
      ids[i] = (int) P[i].ID;
    }
  put(prefix, "id", 'i', n, 1, ids);
  put(prefix, "phkey", 'q', n, 1, ph);
  put(prefix, "mortonkey", 'q', n, 1, mo);
  free(ph); free(mo); free(ids);
}

static void initial_hsml(void)
{
  /* restates init.c:719-745 (RestartFlag == 0) */
  int i, no, p;
  for(i = 0; i < N_gas; i++)
    {
      no = Father[i];
      while(10 * All.DesNumNgb * P[i].Mass > Nodes[no].u.d.mass)
	{
	  p = Nodes[no].u.d.father;
	  if(p < 0) break;
	  no = p;
	}
      P[i].Hsml = pow(3.0 / (4 * M_PI) * All.DesNumNgb * P[i].Mass / Nodes[no].u.d.mass, 1.0 / 3) * Nodes[no].len;
      if(P[i].Hsml > 10.0 * All.SofteningTable[0])
	P[i].Hsml = All.SofteningTable[0];
    }
  for(i = 0; i < NumPart; i++)
    {
      if(P[i].Type == 5) P[i].Hsml = All.SofteningTable[5];
      if(P[i].Type == 2) P[i].Hsml = All.SofteningTable[2];
    }
}

int main(int argc, char **argv)
{
  FILE *fin;
  char magic[4];
  int version, i, k;
  long long N, Ngas;
  double *pos, *vel, *mass, *hsml, *oldacc, *entropy, *dtentropy;
  int *type, *id, *timebin;
  char *ops, *tok;
  int opidx = 0;
  double interactions;

  if(argc < 4) { fprintf(stderr, "usage: %s in out \"ops\"\n", argv[0]); return 2; }
  fin = fopen(argv[1], "rb");
  if(!fin) { perror(argv[1]); return 2; }
  if(fread(magic, 1, 4, fin) != 4 || memcmp(magic, "GLIN", 4)) { fprintf(stderr, "bad magic\n"); return 2; }
  if(fread(&version, 4, 1, fin) != 1 || fread(&N, 8, 1, fin) != 1 || fread(&Ngas, 8, 1, fin) != 1) return 2;
  if(fread(par, 8, NPAR, fin) != NPAR) return 2;
#define RD(ptr, T, cnt) do { ptr = malloc(sizeof(T) * (size_t) (cnt) + 8); if(fread(ptr, sizeof(T), (size_t) (cnt), fin) != (size_t) (cnt)) { fprintf(stderr, "short read %s\n", #ptr); return 2; } } while(0)
  RD(pos, double, 3 * N); RD(vel, double, 3 * N); RD(mass, double, N); RD(type, int, N);
  RD(hsml, double, N); RD(id, int, N); RD(timebin, int, N); RD(oldacc, double, N);
  RD(entropy, double, Ngas); RD(dtentropy, double, Ngas);
  fclose(fin);
  fout = fopen(argv[2], "wb");
  if(!fout) { perror(argv[2]); return 2; }
  fwrite("GLOU", 1, 4, fout);

  MPI_Init(&argc, &argv);
  ThisTask = 0; NTask = 1; PTask = 0;
  mymalloc_init();
  memset(&All, 0, sizeof(All));
  All.PartAllocFactor = par[P_PARTALLOC] > 0 ? par[P_PARTALLOC] : 1.2;
  All.TotNumPart = N; All.TotN_gas = Ngas;
  All.MaxPart = (int) (All.PartAllocFactor * N) + 64;
  All.MaxPartSph = (int) (All.PartAllocFactor * Ngas) + 64;
  All.BufferSize = 200;
  All.TopNodeAllocFactor = 0.02;	/* init.c:156 */
  All.TreeAllocFactor = 2.0;		/* init.c:157 */
  All.ErrTolTheta = par[P_THETA];
  All.ErrTolForceAcc = par[P_ERRTOLFORCEACC];
  All.TypeOfOpeningCriterion = (int) par[P_OPENCRIT];
  All.DesNumNgb = (int) par[P_DESNUMNGB];
  All.MaxNumNgbDeviation = par[P_MAXNGBDEV];
  All.ArtBulkViscConst = par[P_ARTVISC];
  All.G = par[P_G];
  All.SofteningGas = par[P_SOFT0 + 0]; All.SofteningHalo = par[P_SOFT0 + 1]; All.SofteningDisk = par[P_SOFT0 + 2];
  All.SofteningBulge = par[P_SOFT0 + 3]; All.SofteningStars = par[P_SOFT0 + 4]; All.SofteningBndry = par[P_SOFT0 + 5];
  All.MinGasHsmlFractional = par[P_MINGASHSMLFRAC];
  All.Timebase_interval = par[P_TIMEBASE_INTERVAL];
  All.Ti_Current = (int) par[P_TI_CURRENT];
  All.BlackHoleNgbFactor = par[P_BHNGBFACTOR];
  All.TreeDomainUpdateFrequency = 0.1;
  All.Cadj_Cost = 1.0e-30; All.Cadj_Cpu = 1.0e-3;
  All.Time = All.TimeBegin = 0; All.TimeMax = 1;
  All.ComovingIntegrationOn = 0;
  All.NumForcesSinceLastDomainDecomp = (long long) (1 + All.TotNumPart * All.TreeDomainUpdateFrequency);
  set_softenings();
  allocate_memory();
  set_random_numbers();
  FdTimings = fopen("/dev/null", "w");
  FdCPU = fopen("/dev/null", "w");

  NumPart = (int) N; N_gas = (int) Ngas;
  memset(P, 0, sizeof(struct particle_data) * (size_t) All.MaxPart);
  memset(SphP, 0, sizeof(struct sph_particle_data) * (size_t) All.MaxPartSph);
  for(i = 0; i < NumPart; i++)
    {
      for(k = 0; k < 3; k++) { P[i].Pos[k] = pos[3 * i + k]; P[i].Vel[k] = vel[3 * i + k]; }
      P[i].Mass = mass[i]; P[i].Type = type[i]; P[i].ID = id[i]; P[i].Hsml = hsml[i];
      P[i].TimeBin = timebin[i]; P[i].OldAcc = oldacc[i]; P[i].GravCost = 1;	/* init.c:250-252 */
      P[i].Ti_current = All.Ti_Current; P[i].Ti_begstep = 0;
      if(i < N_gas)
	{
	  if(type[i] != 0) { fprintf(stderr, "gas must come first\n"); return 2; }
	  for(k = 0; k < 3; k++) SphP[i].VelPred[k] = vel[3 * i + k];
	  SphP[i].Entropy = entropy[i]; SphP[i].e.DtEntropy = dtentropy[i]; SphP[i].d.Density = -1;
	}
    }
  for(i = 0; i < TIMEBINS; i++) TimeBinActive[i] = 1;
  reconstruct_timebins();
  TreeReconstructFlag = 1;
  {
    double t0 = now(), t1;
    domain_Decomposition();	/* also peano_hilbert_order() and force_treeallocate() */
    t1 = now() - t0;
    put("", "time_domain", 'd', 1, 1, &t1);
  }
  {
    /* the top-level tree the decomposition built (domain.c:1738 domain_determineTopTree -> TopNodes[], domain.c:220-226) */
    long long *sk = malloc(8 * (NTopnodes + 1)), *sz = malloc(8 * (NTopnodes + 1));
    int *da = malloc(4 * (NTopnodes + 1)), *lf = malloc(4 * (NTopnodes + 1)), nn[2];
    for(i = 0; i < NTopnodes; i++)
      {
	sk[i] = (long long) TopNodes[i].StartKey; sz[i] = (long long) TopNodes[i].Size;
	da[i] = TopNodes[i].Daughter; lf[i] = TopNodes[i].Daughter < 0 ? TopNodes[i].Leaf : -1;
      }
    nn[0] = NTopnodes; nn[1] = NTopleaves;
    put("", "top_n", 'i', 2, 1, nn);
    put("", "top_StartKey", 'q', NTopnodes, 1, sk); put("", "top_Size", 'q', NTopnodes, 1, sz);
    put("", "top_Daughter", 'i', NTopnodes, 1, da); put("", "top_Leaf", 'i', NTopnodes, 1, lf);
    /* the split of the leaves into MULTIPLEDOMAINS * NTask stretches and their owners (domain.c:1075, :1153) */
    put("", "top_StartList", 'i', MULTIPLEDOMAINS * NTask, 1, DomainStartList);
    put("", "top_EndList", 'i', MULTIPLEDOMAINS * NTask, 1, DomainEndList);
    put("", "top_DomainTask", 'i', NTopleaves, 1, DomainTask);
    {
      double dom[5] = { DomainCorner[0], DomainCorner[1], DomainCorner[2], DomainFac, (double) (int) (All.TopNodeAllocFactor * All.MaxPart + 1) };
      put("", "top_domain", 'd', 5, 1, dom);
    }
    free(sk); free(sz); free(da); free(lf);
  }

  ops = strdup(argv[3]);
  for(tok = strtok(ops, " "); tok; tok = strtok(NULL, " "), opidx++)
    {
      char prefix[32];
      double t0, dt;
      const char *seq = tok;
      if(tok[0] == 'A') seq = "RGDUH";
      for(; *seq; seq++)
	{
	  char op = *seq;
	  snprintf(prefix, sizeof(prefix), "%d_%c_", opidx, op);
	  t0 = now();
	  switch (op)
	    {
	    case 'T': ngb_treebuild(); TreeReconstructFlag = 0; break;
	    case 'S':
	      initial_hsml();
	      /* input convention: OldAcc < 0 marks a gas particle that is swallowed (Mass = 0, sfr/blackhole code) after
	         the initial-Hsml step but is still in the tree - the state density.c:832 / hydra.c:1236 guard against */
	      for(i = 0; i < NumPart; i++)
		if(P[i].OldAcc < 0)
		  {
		    P[i].Mass = 0;
		    P[i].OldAcc = 0;
		  }
	      break;
	    case 'D': density(); break;
	    /* drop-in binaries keep the reference's own CPU loops under ref_* (objcopy, see Makefile): 'd' / 'h' run them on
	       whatever tree Nodes[] holds - the GPU tree written by gl_export_tree in the ref_dropin_tree build */
	    case 'd': if(!ref_density) { fprintf(stderr, "op d needs a drop-in build\n"); return 2; } ref_density(); break;
	    case 'h': if(!ref_hydro_force) { fprintf(stderr, "op h needs a drop-in build\n"); return 2; } ref_hydro_force(); break;
	    case 'R': TreeReconstructFlag = 1; break;
	    case 'G': gravity_tree(); break;
	    case 'U': force_update_hmax(); break;
	    case 'H': hydro_force(); break;
	    case 'X':
	      /* lazy drift of the sources to a later time (predict.c:129, called from the walks: forcetree.c:1911, ngb.c:57):
	         give the dust drag a non-zero value, leave every third particle already current, dump, drift, dump */
	      {
		int k, t1 = All.Ti_Current + 24;
		for(i = 0; i < N_gas; i++)
		  for(k = 0; k < 3; k++) SphP[i].da.DragAccel[k] = 0.01 * SphP[i].a.HydroAccel[(k + 1) % 3];
		for(i = 0; i < NumPart; i += 3) P[i].Ti_current = t1;
		{
		  double q[3] = { (double) All.Ti_Current, (double) t1, All.Timebase_interval };
		  put(prefix, "times", 'd', 3, 1, q);
		  q[0] = All.MinGasHsml;
		  put(prefix, "mingashsml", 'd', 1, 1, q);
		}
		dump_drift_state(prefix, "pre");
		All.Ti_Current = t1;
		for(i = 0; i < NumPart; i++) drift_particle(i, All.Ti_Current);
		dump_drift_state(prefix, "post");
	      }
	      break;
	    case 'V':
	      /* advance_and_find_timesteps (timestep.c:29) at the synchronisation point Ti_Current = 64 of an individual-time-
	         step state: a particle on bin b (1..6, from the input) began its step at 64 - 2^b; bins above 6 are not
	         synchronised there, so a particle that wants a longer step than 64 has to stay where it is (timestep.c:147-154) */
	      {
		int k, nact = 0;
		double q[20];
		All.Ti_Current = 64;
		All.ErrTolIntAccuracy = 0.025; All.CourantFac = 0.15; All.TypeOfTimestepCriterion = 0;
		All.MaxSizeTimestep = 200 * All.Timebase_interval; All.MinSizeTimestep = 0;
		All.MinEgySpec = par[P_MINEGYSPEC];
		All.OuterBoundary = 1.4e10; All.UnitVelocity_in_cm_per_s = 2.98e6; All.FeedBackVelocity = 1.0;
		All.SMBHmass = 0.5; All.InnerBoundary = 2.0e-2; All.SinkBoundary = 1.0e-4; All.AccDtBlackHole = 0.007;
		for(i = 0; i < NumPart; i++)
		  {
		    int b = P[i].TimeBin < 1 ? 1 : (P[i].TimeBin > 6 ? 6 : P[i].TimeBin);
		    P[i].TimeBin = b;
		    P[i].Ti_begstep = All.Ti_Current - (1 << b);
		    P[i].Ti_current = All.Ti_Current;
		    if(i % 5 == 0)
		      for(k = 0; k < 3; k++) P[i].DeltaDustMomentum[k] = 1.0e-3 * P[i].Mass * P[i].g.GravAccel[(k + 2) % 3];
		  }
		for(i = 0; i < N_gas; i++)
		  {
		    for(k = 0; k < 3; k++) SphP[i].da.DragAccel[k] = 0.01 * SphP[i].a.HydroAccel[(k + 1) % 3];
		    /* strong cooling for every seventh particle: both entropy limiters of do_the_kick (timestep.c:546-583) */
		    if(i % 7 == 0) SphP[i].e.DtEntropy = -50 * fabs(SphP[i].e.DtEntropy) - 1.0;
		  }
		for(i = 0; i < TIMEBINS; i++) TimeBinActive[i] = (i <= 6);
		reconstruct_timebins();
		for(i = FirstActiveParticle; i >= 0; i = NextActiveParticle[i]) nact++;
		Flag_FullStep = 1;
		q[0] = All.Ti_Current; q[1] = All.Timebase_interval; q[2] = All.ErrTolIntAccuracy; q[3] = All.CourantFac;
		q[4] = All.MaxSizeTimestep; q[5] = All.MinSizeTimestep; q[6] = All.MinEgySpec; q[7] = All.OuterBoundary;
		q[8] = All.UnitVelocity_in_cm_per_s; q[9] = All.FeedBackVelocity; q[10] = All.SMBHmass; q[11] = All.InnerBoundary;
		q[12] = All.SinkBoundary; q[13] = All.AccDtBlackHole; q[14] = nact;
		for(k = 0; k < 6; k++) q[15 + k < 20 ? 15 + k : 19] = 0;
		put(prefix, "params", 'd', 15, 1, q);
		put(prefix, "SofteningTable", 'd', 6, 1, All.SofteningTable);
		{
		  int *act = malloc(sizeof(int) * (nact + 1));
		  for(i = FirstActiveParticle, k = 0; i >= 0; i = NextActiveParticle[i]) act[k++] = i;
		  put(prefix, "active", 'i', nact, 1, act);
		  free(act);
		}
		dump_kick_state(prefix, "pre");
		advance_and_find_timesteps();
		dump_kick_state(prefix, "post");
	      }
	      break;
	    case 'L':
	      /* several iterations of run()'s loop (run.c:63-110) on individual time steps, starting at the synchronisation
	         point of op V: find_next_sync_point_and_drift (run.c:190), domain_Decomposition, the force path for the
	         active particles, advance_and_find_timesteps.  TreeDomainUpdateFrequency = 0: a new decomposition and a new
	         tree in every step (no dynamic tree update), which is how the GPU path runs. */
	      {
		int k, step;
		double q[4];
		All.Ti_Current = 64;
		All.ErrTolIntAccuracy = 0.025; All.CourantFac = 0.15; All.TypeOfTimestepCriterion = 0;
		All.MaxSizeTimestep = 200 * All.Timebase_interval; All.MinSizeTimestep = 0;
		All.MinEgySpec = par[P_MINEGYSPEC];
		All.OuterBoundary = 1.4e10; All.UnitVelocity_in_cm_per_s = 2.98e6; All.FeedBackVelocity = 1.0;
		All.SMBHmass = 0.5; All.InnerBoundary = 2.0e-2; All.SinkBoundary = 1.0e-4; All.AccDtBlackHole = 0.007;
		All.TreeDomainUpdateFrequency = 0; All.Ti_nextoutput = -1; All.TimeBegin = 0;
		for(i = 0; i < NumPart; i++)
		  {
		    int b = P[i].TimeBin < 1 ? 1 : (P[i].TimeBin > 6 ? 6 : P[i].TimeBin);
		    P[i].TimeBin = b;
		    P[i].Ti_begstep = All.Ti_Current - (1 << b);
		    P[i].Ti_current = All.Ti_Current;
		  }
		for(i = 0; i < TIMEBINS; i++) TimeBinActive[i] = (i <= 6);
		reconstruct_timebins();
		Flag_FullStep = 1;
		/* the state the loop starts from, complete enough to start the same loop elsewhere (the device-resident
		   loop gl_run_steps, tests/test_gpu_resident.py) */
		dump_kick_state(prefix, "begin");
		dump_loop_extra(prefix, "begin");
		advance_and_find_timesteps();
		for(step = 0; step < 8; step++)
		  {
		    find_next_sync_point_and_drift();
		    domain_Decomposition();
		    compute_accelerations(0);
		    advance_and_find_timesteps();
		    All.NumCurrentTiStep++;
		  }
		for(i = 0, k = 0; i < NumPart; i++) k += (P[i].Ti_current == All.Ti_Current);
		q[0] = All.Ti_Current; q[1] = All.Time; q[2] = k; q[3] = step;
		put(prefix, "loop", 'd', 4, 1, q);
		dump_kick_state(prefix, "end");
		dump_loop_extra(prefix, "end");
	      }
	      break;
	    case 'M':
	      /* op L with the eight iterations handed to the glue's gl_glue_run_steps() (device-resident loop); the same dumps.
	         several iterations of run()'s loop (run.c:63-110) on individual time steps, starting at the synchronisation
	         point of op V: find_next_sync_point_and_drift (run.c:190), domain_Decomposition, the force path for the
	         active particles, advance_and_find_timesteps.  TreeDomainUpdateFrequency = 0: a new decomposition and a new
	         tree in every step (no dynamic tree update), which is how the GPU path runs. */
	      {
		int k, step;
		double q[4];
		All.Ti_Current = 64;
		All.ErrTolIntAccuracy = 0.025; All.CourantFac = 0.15; All.TypeOfTimestepCriterion = 0;
		All.MaxSizeTimestep = 200 * All.Timebase_interval; All.MinSizeTimestep = 0;
		All.MinEgySpec = par[P_MINEGYSPEC];
		All.OuterBoundary = 1.4e10; All.UnitVelocity_in_cm_per_s = 2.98e6; All.FeedBackVelocity = 1.0;
		All.SMBHmass = 0.5; All.InnerBoundary = 2.0e-2; All.SinkBoundary = 1.0e-4; All.AccDtBlackHole = 0.007;
		All.TreeDomainUpdateFrequency = 0; All.Ti_nextoutput = -1; All.TimeBegin = 0;
		for(i = 0; i < NumPart; i++)
		  {
		    int b = P[i].TimeBin < 1 ? 1 : (P[i].TimeBin > 6 ? 6 : P[i].TimeBin);
		    P[i].TimeBin = b;
		    P[i].Ti_begstep = All.Ti_Current - (1 << b);
		    P[i].Ti_current = All.Ti_Current;
		  }
		for(i = 0; i < TIMEBINS; i++) TimeBinActive[i] = (i <= 6);
		reconstruct_timebins();
		Flag_FullStep = 1;
		/* the state the loop starts from, complete enough to start the same loop elsewhere (the device-resident
		   loop gl_run_steps, tests/test_gpu_resident.py) */
		dump_kick_state(prefix, "begin");
		dump_loop_extra(prefix, "begin");
		advance_and_find_timesteps();
		if(!gl_glue_run_steps) { fprintf(stderr, "op M needs a drop-in build with -DGL_RESIDENT_RUN\n"); return 2; }
		step = gl_glue_run_steps(8);	/* the same eight iterations, in the library, on the device-resident state */
		for(i = 0, k = 0; i < NumPart; i++) k += (P[i].Ti_current == All.Ti_Current);
		q[0] = All.Ti_Current; q[1] = All.Time; q[2] = k; q[3] = step;
		put(prefix, "loop", 'd', 4, 1, q);
		dump_kick_state(prefix, "end");
		dump_loop_extra(prefix, "end");
	      }
	      break;
	    case 'B':
	      set_subgrid_state(); dump_subgrid_params(prefix); dump_subgrid_state(prefix, "pre");
	      { double tc = now(); blackhole_accretion(); tc = now() - tc; put(prefix, "time", 'd', 1, 1, &tc); }
	      dump_subgrid_state(prefix, "post");
	      break;
	    case 'Y':
	      set_subgrid_state(); dump_subgrid_params(prefix); dump_subgrid_state(prefix, "pre");
	      { double tc = now(); dust_density(); tc = now() - tc; put(prefix, "time", 'd', 1, 1, &tc); }
	      dump_subgrid_state(prefix, "post");
	      break;
	    case 'Z':
	      set_subgrid_state(); dump_subgrid_params(prefix); dump_subgrid_state(prefix, "pre");
	      { double tc = now(); dust_drag(); tc = now() - tc; put(prefix, "time", 'd', 1, 1, &tc); }
	      dump_subgrid_state(prefix, "post");
	      break;
	    case 'N': dump_nodes(prefix); break;
	    case 'C': pair_counts(prefix); break;
	    case 'K': dump_keys(prefix); break;
	    case 'P': compute_potential(); break;
	    default: fprintf(stderr, "unknown op %c\n", op); return 2;
	    }
	  dt = now() - t0;
	  if(strchr("TDGUHPdh", op)) put(prefix, "time", 'd', 1, 1, &dt);
	  if(op == 'P')
	    {
	      long long q;
	      double *pt = malloc(sizeof(double) * NumPart);
	      int *ids = malloc(sizeof(int) * NumPart);
	      for(q = 0; q < NumPart; q++) { pt[q] = P[q].p.Potential; ids[q] = (int) P[q].ID; }
	      put(prefix, "id", 'i', NumPart, 1, ids);
	      put(prefix, "Potential", 'd', NumPart, 1, pt);
	      free(pt); free(ids);
	    }
	  if(strchr("GDHSdh", op)) dump_particles_out(prefix, (op == 'd') ? 'D' : (op == 'h') ? 'H' : op);
	  if(op == 'G')
	    {
	      interactions = 0;
	      for(i = 0; i < NumPart; i++) interactions += P[i].GravCost;
	      put(prefix, "interactions", 'd', 1, 1, &interactions);
	      put(prefix, "theta_after", 'd', 1, 1, &All.ErrTolTheta);
	    }
	  if(op == 'T' || op == 'G')
	    {
	      int nn = Numnodestree;
	      put(prefix, "numnodes", 'i', 1, 1, &nn);
	    }
	}
    }
  /* reference time buckets (allvars.h:205-238) */
  put("", "CPU_Step", 'd', CPU_PARTS, 1, CPU_Step);
  fclose(fout);
  return 0;
}
