/* ref_io.c - drives the UNMODIFIED reference snapshot writer and initial-conditions reader
 * (savepositions io.c:33 -> write_file io.c:2521 -> fill_write_buffer io.c:173;
 *  read_ic read_ic.c:28 -> read_file read_ic.c:634 -> empty_read_buffer read_ic.c:238)
 * on raw images of P[] / SphP[], so that the library's gl_snapshot_write / gl_snapshot_read can be
 * compared with them byte for byte.
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled together with the reference's own sources where they lie under
 * /root/reference (Makefile target `io`); the binary goes to oracle/_ref/ and is never linked into, or called
 * by, the product.
 *
 * usage:
 *   ref_io W key=value ...   P/SphP images in (p=, sph=, n=, ngas=), snapshot out: <dir>/<base>_000
 *   ref_io R key=value ...   IC / snapshot file in (file=), P/SphP images out (p=, sph=) + meta text (meta=)
 * keys: format, cooling, sfr, restart, minegy, initgastemp, initgasu-independent units (unitmass, unitenergy),
 *       ti_current, timebase, time, mt0..mt5 (All.MassTable), dir, base, box, omega0, omegal, hubble
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <mpi.h>

#include "allvars.h"
#include "proto.h"

/* ---- symbols the linked reference files leave undefined ---- */
double get_drift_factor(int t0, int t1) { (void) t0; (void) t1; return 0; }
double get_gravkick_factor(int t0, int t1) { (void) t0; (void) t1; return 0; }
double get_hydrokick_factor(int t0, int t1) { (void) t0; (void) t1; return 0; }
double hubble_function(double a) { (void) a; return 0; }
void rearrange_particle_sequence(void) { /* no Mass==0 particles in the test images (sfr_eff.c:1018) */ }
double gsl_rng_uniform(gsl_rng * r) { (void) r; return 0.5; }
void compute_accelerations(int mode) { (void) mode; }	/* run.o is linked with the other reference objects */

static const char *arg(int argc, char **argv, const char *key, const char *dflt)
{
  size_t l = strlen(key);
  int i;
  for(i = 2; i < argc; i++)
    if(!strncmp(argv[i], key, l) && argv[i][l] == '=')
      return argv[i] + l + 1;
  return dflt;
}

#define ARGD(k, d) atof(arg(argc, argv, k, d))
#define ARGI(k, d) atoi(arg(argc, argv, k, d))

static void slurp(const char *fn, void *dst, size_t bytes)
{
  FILE *f = fopen(fn, "rb");
  if(!f || fread(dst, 1, bytes, f) != bytes) { fprintf(stderr, "ref_io: cannot read %zu bytes from %s\n", bytes, fn); exit(2); }
  fclose(f);
}

static void spill(const char *fn, const void *src, size_t bytes)
{
  FILE *f = fopen(fn, "wb");
  if(!f || fwrite(src, 1, bytes, f) != bytes) { fprintf(stderr, "ref_io: cannot write %s\n", fn); exit(2); }
  fclose(f);
}

int main(int argc, char **argv)
{
  int i;
  char key[8];
  if(argc < 2) { fprintf(stderr, "usage: %s W|R key=value ...\n", argv[0]); return 2; }
  MPI_Init(&argc, &argv);
  ThisTask = 0; NTask = 1; PTask = 0;
  mymalloc_init();
  memset(&All, 0, sizeof(All));
  All.BufferSize = 20;
  All.PartAllocFactor = 1.5;
  All.TopNodeAllocFactor = 0.02;
  All.TreeAllocFactor = 2.0;
  All.TreeDomainUpdateFrequency = 0.1;
  All.NumFilesPerSnapshot = 1;
  All.NumFilesWrittenInParallel = 1;
  All.SnapFormat = All.ICFormat = ARGI("format", "1");
  All.CoolingOn = ARGI("cooling", "0");
  All.StarformationOn = ARGI("sfr", "0");
  All.MinEgySpec = ARGD("minegy", "0");
  All.InitGasTemp = ARGD("initgastemp", "0");
  All.UnitMass_in_g = ARGD("unitmass", "1.989e33");
  All.UnitEnergy_in_cgs = ARGD("unitenergy", "1.989e43");
  All.Ti_Current = ARGI("ti_current", "0");
  All.Timebase_interval = ARGD("timebase", "0");
  All.Time = ARGD("time", "0");
  All.BoxSize = ARGD("box", "0");
  All.Omega0 = ARGD("omega0", "0");
  All.OmegaLambda = ARGD("omegal", "0");
  All.HubbleParam = ARGD("hubble", "1");
  All.ComovingIntegrationOn = 0;
  for(i = 0; i < 6; i++)
    {
      sprintf(key, "mt%d", i);
      All.MassTable[i] = ARGD(key, "0");
    }
  RestartFlag = ARGI("restart", "0");

  if(argv[1][0] == 'W')
    {
      long long n = atoll(arg(argc, argv, "n", "0")), ngas = atoll(arg(argc, argv, "ngas", "0"));
      All.TotNumPart = n; All.TotN_gas = ngas;
      All.MaxPart = (int) n + 16; All.MaxPartSph = (int) n + 16;
      allocate_memory();
      NumPart = (int) n; N_gas = (int) ngas;
      slurp(arg(argc, argv, "p", ""), P, sizeof(struct particle_data) * (size_t) n);
      if(ngas > 0) slurp(arg(argc, argv, "sph", ""), SphP, sizeof(struct sph_particle_data) * (size_t) ngas);
      snprintf(All.OutputDir, sizeof(All.OutputDir), "%s/", arg(argc, argv, "dir", "."));
      snprintf(All.SnapshotFileBase, sizeof(All.SnapshotFileBase), "%s", arg(argc, argv, "base", "snap"));
      savepositions(0);
    }
  else
    {
      char fname[500];
      FILE *m;
      snprintf(fname, sizeof(fname), "%s", arg(argc, argv, "file", ""));
      read_ic(fname);
      spill(arg(argc, argv, "p", "P.out"), P, sizeof(struct particle_data) * (size_t) NumPart);
      spill(arg(argc, argv, "sph", "SphP.out"), SphP, sizeof(struct sph_particle_data) * (size_t) (N_gas > 0 ? N_gas : 0));
      m = fopen(arg(argc, argv, "meta", "meta.txt"), "w");
      fprintf(m, "NumPart %d\nN_gas %d\nTotNumPart %lld\nTotN_gas %lld\nInitGasU %.17g\ntime %.17g\n", NumPart, N_gas,
	      (long long) All.TotNumPart, (long long) All.TotN_gas, All.InitGasU, header.time);
      for(i = 0; i < 6; i++) fprintf(m, "MassTable%d %.17g\n", i, All.MassTable[i]);
      for(i = 0; i < 6; i++) fprintf(m, "npart%d %d\n", i, header.npart[i]);
      fclose(m);
    }
  return 0;
}
