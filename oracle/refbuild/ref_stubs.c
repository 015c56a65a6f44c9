/* ref_stubs.c - the handful of symbols run.c / io.c reference that live in reference files outside the path
 * (cooling.c, sfr_eff.c, restart.c, global.c, begrun.c).  TEST INFRASTRUCTURE ONLY (oracle/_ref binaries). */
#include <stdio.h>
#include <stdlib.h>
#include <mpi.h>
#include "allvars.h"
#include "proto.h"

/* cooling.c / sfr_eff.c physics is outside the path: the NH and SFR records of a snapshot get a recognisable function
 * of the particle so that record placement can still be compared (the library takes these two records from the caller) */
double AbundanceRatios(double u, double rho, double *ne_guess, double *nH0_pointer, double *nHeII_pointer)
{
  (void) u; (void) ne_guess;
  *nH0_pointer = 0.25 + 0.5 * rho / (1.0 + rho);
  *nHeII_pointer = 0;
  return 0;
}
double get_starformation_rate(int i) { return 1.0e-3 * (i % 17); }
void IonizeParams(void) { }
void close_outputfiles(void) { }
void compute_global_quantities_of_system(void) { }
void restart(int mod) { (void) mod; }
