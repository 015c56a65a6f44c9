/* layout_probe.c - prints the byte offsets of the P[] / SphP[] fields the force path touches, as the
 * reference's allvars.h defines them under the shipped Makefile's -D list (SURVEY.md 8(b) "ABI caveat").
 * Build-container tool: `make layout` writes gadget_leicester_b200/gadget_layout.json. */
#include <stdio.h>
#include <stddef.h>
#include <mpi.h>
#include "allvars.h"

#define PO(name, field) printf("  \"%s\": %zu,\n", name, offsetof(struct particle_data, field))
#define SO(name, field) printf("  \"%s\": %zu,\n", name, offsetof(struct sph_particle_data, field))
int main(void)
{
  printf("{\n");
  printf("  \"p_stride\": %zu,\n  \"sph_stride\": %zu,\n", sizeof(struct particle_data), sizeof(struct sph_particle_data));
  PO("p_Pos", Pos); PO("p_Vel", Vel); PO("p_Mass", Mass); PO("p_GravAccel", g.GravAccel); PO("p_OldAcc", OldAcc);
  PO("p_Hsml", Hsml); PO("p_NumNgb", n.NumNgb); PO("p_GravCost", GravCost); PO("p_Ti_begstep", Ti_begstep);
  PO("p_Type", Type); PO("p_TimeBin", TimeBin);
  PO("p_SinkDensity", b1.BH_Density); PO("p_SinkEntropy", b2.BH_Entropy); PO("p_SinkGasVel", b3.BH_SurroundingGasVel);
  PO("p_DustDensity", d1.DUST_Density); PO("p_DustEntropy", d2.DUST_Entropy); PO("p_DustGasVel", d3.DUST_SurroundingGasVel);
  PO("p_SwallowID", SwallowID); PO("p_ID", ID); PO("p_Ti_current", Ti_current);
  SO("s_DragAccel", da.DragAccel);
  SO("s_Entropy", Entropy); SO("s_Pressure", Pressure); SO("s_VelPred", VelPred); SO("s_MaxSignalVel", MaxSignalVel);
  SO("s_Density", d.Density); SO("s_DtEntropy", e.DtEntropy); SO("s_HydroAccel", a.HydroAccel);
  SO("s_DhsmlDensityFactor", h.DhsmlDensityFactor); SO("s_DivVel", v.DivVel); SO("s_CurlVel", r.CurlVel);
  /* fields only the snapshot / IC files carry (io.c:173, read_ic.c:238) */
  PO("io_p_DustRadius", DustRadius); PO("io_p_StellarAge", StellarAge); PO("io_p_BH_Mass", BH_Mass); PO("io_p_BH_Mdot", BH_Mdot);
  PO("io_p_AccDisc_Mass", AccDisc_Mass); PO("io_p_BH_Mseed", BH_Mseed); PO("io_p_NewDensity", NewDensity);
  PO("io_p_DeltaPhotonMomentum", DeltaPhotonMomentum); PO("io_p_OldPhotonMomentum", OldPhotonMomentum);
  PO("io_p_BirthPhotonEnergy", BirthPhotonEnergy);
  PO("ts_p_DeltaDustMomentum", DeltaDustMomentum);	/* get_timestep / do_the_kick, timestep.c:686, :459 */
  SO("io_s_Ne", Ne); SO("io_s_accreted_mass", accreted_mass); SO("io_s_accreted_momentum", accreted_momentum);
  printf("  \"io_sizeof_ID\": %zu,\n  \"io_sizeof_header\": %zu,\n", sizeof(MyIDType), sizeof(struct io_header));
#define NO(name, field) printf("  \"%s\": %zu,\n", name, offsetof(struct NODE, field))
#define EO(name, field) printf("  \"%s\": %zu,\n", name, offsetof(struct extNODE, field))
  NO("n_len", len); NO("n_center", center); NO("n_maxsoft", maxsoft); NO("n_s", u.d.s); NO("n_mass", u.d.mass);
  NO("n_bitflags", u.d.bitflags); NO("n_sibling", u.d.sibling); NO("n_nextnode", u.d.nextnode); NO("n_father", u.d.father);
  NO("n_Ti_current", Ti_current);
  EO("e_vs", vs); EO("e_vmax", vmax); EO("e_divVmax", divVmax); EO("e_hmax", hmax); EO("e_Ti_lastkicked", Ti_lastkicked);
  EO("e_Flag", Flag);
  printf("  \"bitflags_value\": %u,\n", 1u << 7);	/* 1 << BITFLAG_MULTIPLEPARTICLES, forcetree.h:18 */
  printf("  \"sizeof_NODE\": %zu,\n  \"sizeof_extNODE\": %zu\n}\n", sizeof(struct NODE), sizeof(struct extNODE));
  return 0;
}
