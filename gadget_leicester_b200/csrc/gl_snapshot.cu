// gl_snapshot.cu - the reference's binary snapshot / initial-conditions files (formats 1 and 2) written from and read
// into its own P[] / SphP[] arrays.  Host code only: no context, no GPU - the file is compiled
// with the rest of the library so that the reference-side glue links one .so.
//
// The record list is the reference's `enum iofields` (allvars.h:1762) restricted to what blockpresent() (io.c:1639)
// lets through under the shipped -D list; every record is described by one row of the table below (which types carry
// it: get_particles_in_block io.c:1460; how wide it is: get_bytes_per_blockelement io.c:1184; its format-2 label:
// get_Tab_IO_Label io.c:2038).  Values are 4-byte floats, IDs 4-byte unsigned.
#include "gl_internal.h"
#include "gl_host.h"
#include "gl_hdf5.h"
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

namespace
{
double gl_now()
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

enum Block
{
  B_POS, B_VEL, B_ID, B_MASS, B_U, B_RHO, B_NE, B_NH, B_HSML, B_DUST, B_SFR, B_AGE, B_BHMASS, B_BHMDOT, B_ACCDISCMASS, B_BHMSEED,
  B_NEWDENS, B_DELPHOTONMOM, B_OLDPHOTONMOM, B_BIPHOTONEN, B_ACCMASS, B_ACCMOM, B_COUNT
};

struct BlockRow
{
  Block b;
  const char *label;		// format 2
  const char *h5name;		// format 3: get_dataset_name() io.c:2270
  int ncomp;
  int types;			// bit mask of the particle types that carry the record (before the MASS / npart rules)
  bool behind_u;		// not read from initial conditions (read_ic.c:891)
};

const int ALL_TYPES = 63, GAS = 1, STARS = 1 << 4, SINKS = 1 << 5, VIRTUALS = 1 << 3;
const BlockRow ROWS[B_COUNT] = {
  {B_POS, "POS ", "Coordinates", 3, ALL_TYPES, false}, {B_VEL, "VEL ", "Velocities", 3, ALL_TYPES, false}, {B_ID, "ID  ", "ParticleIDs", 1, ALL_TYPES, false},
  {B_MASS, "MASS", "Masses", 1, ALL_TYPES, false}, {B_U, "U   ", "InternalEnergy", 1, GAS, false}, {B_RHO, "RHO ", "Density", 1, GAS, true},
  {B_NE, "NE  ", "ElectronAbundance", 1, GAS, true}, {B_NH, "NH  ", "NeutralHydrogenAbundance", 1, GAS, true}, {B_HSML, "HSML", "SmoothingLength", 1, GAS, true},
  {B_DUST, "DUST", "DustRadius", 1, ALL_TYPES, true}, {B_SFR, "SFR ", "StarFormationRate", 1, GAS, true}, {B_AGE, "AGE ", "StellarFormationTime", 1, STARS, true},
  {B_BHMASS, "BHMA", "BH_Mass", 1, SINKS, true}, {B_BHMDOT, "BHMD", "BH_Mdot", 1, SINKS, true}, {B_ACCDISCMASS, "ACDM", "AccDisc_Mass", 1, SINKS, true},
  {B_BHMSEED, "BHMS", "BH_Mseed", 1, SINKS, true}, {B_NEWDENS, "NEWD", "NewDensity", 1, VIRTUALS, true}, {B_DELPHOTONMOM, "DPhM", "DeltaPhotonMomentum", 1, VIRTUALS, true},
  {B_OLDPHOTONMOM, "OPhM", "OldPhotonMomentum", 1, VIRTUALS, true}, {B_BIPHOTONEN, "BPhE", "BirthPhotonEnergy", 1, VIRTUALS, true}, {B_ACCMASS, "ACMA", "accreted_mass", 1, GAS, true},
  // ACMO: get_bytes_per_blockelement() (io.c:1184) gives this record ONE value per particle although fill_write_buffer
  // (io.c:1003-1014) produces three, so write_file puts only the first N_gas values of the (particle, component)
  // stream into the file: the momenta of the first N_gas/3 gas particles.  Kept as it is - the record length is part
  // of the format the reference and its readers expect.  (One CommBuffer pass assumed: N_gas <= BufferSize * 2^18.)
  {B_ACCMOM, "ACMO", "accreted_momentum", 1, GAS, true},
};

struct Ctx
{
  gl_gadget_layout L;
  gl_io_layout IO;
  gl_io_params par;
  char *err;
  int errlen;
  int fail(const char *fmt, ...)
  {
    if(err && errlen > 0)
      {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(err, (size_t) errlen, fmt, ap);
	va_end(ap);
      }
    return 1;
  }
};

// blockpresent() io.c:1639 under the shipped -D list, expressed through the layout (a build without the field has no
// such record) and the two run-time switches
bool present(const Ctx &c, Block b)
{
  switch (b)
    {
    case B_POS: case B_VEL: case B_ID: case B_MASS: case B_U: case B_RHO: case B_HSML:
      return true;
    case B_NE: case B_NH:
      return c.par.CoolingOn != 0 && c.IO.s_Ne >= 0;
    case B_SFR:
      return c.par.StarformationOn != 0 && (c.par.build & GL_IO_BUILD_SFR);
    case B_AGE:
      return c.par.StarformationOn != 0 && c.IO.p_StellarAge >= 0;
    case B_DUST:
      return c.IO.p_DustRadius >= 0;
    case B_BHMASS: return c.IO.p_BH_Mass >= 0;
    case B_BHMDOT: return c.IO.p_BH_Mdot >= 0;
    case B_ACCDISCMASS: return c.IO.p_AccDisc_Mass >= 0;
    case B_BHMSEED: return c.IO.p_BH_Mseed >= 0;
    case B_NEWDENS: return c.IO.p_NewDensity >= 0;
    case B_DELPHOTONMOM: return c.IO.p_DeltaPhotonMomentum >= 0;
    case B_OLDPHOTONMOM: return c.IO.p_OldPhotonMomentum >= 0;
    case B_BIPHOTONEN: return c.IO.p_BirthPhotonEnergy >= 0;
    case B_ACCMASS: return c.IO.s_accreted_mass >= 0;
    case B_ACCMOM: return c.IO.s_accreted_momentum >= 0;
    default: return false;
    }
}

// get_particles_in_block() io.c:1460: which types the record holds in THIS file, and how many elements
long long particles_in_block(const Ctx &c, const BlockRow &r, const int npart[6], int typelist[6])
{
  long long n = 0;
  for(int t = 0; t < 6; t++)
    {
      typelist[t] = (npart[t] > 0) && ((r.types >> t) & 1);
      if(r.b == B_MASS && c.par.MassTable[t] != 0) typelist[t] = 0;	// io.c:1492-1498
      if(typelist[t]) n += npart[t];
    }
  return n;
}

// a P[] field of particle i / an SphP[] field of gas particle i as double
inline double pf(const Ctx &c, const char *P, long long i, int off, int k = 0) { return ((const double *) (P + i * c.L.p_stride + off))[k]; }
inline double sf(const Ctx &c, const char *S, long long i, int off, int k = 0) { return ((const double *) (S + i * c.L.sph_stride + off))[k]; }

// fill_write_buffer() io.c:173 for one particle; out gets ncomp 4-byte values
inline void value_out(const Ctx &c, Block b, const char *P, const char *S, long long i, const double *nh, const double *sfr, float *out)
{
  const gl_gadget_layout &L = c.L;
  switch (b)
    {
    case B_POS:
      for(int k = 0; k < 3; k++) out[k] = (float) pf(c, P, i, L.p_Pos, k);
      break;
    case B_VEL:
      {
	// io.c:256-299, non-comoving: the kick from the middle of the particle's step to the output time
	const char *pi = P + i * L.p_stride;
	const int tb = *(const short *) (pi + L.p_TimeBin), type = *(const short *) (pi + L.p_Type);
	const int dt_step = tb ? (1 << tb) : 0;
	const double dt_kick = (c.par.Ti_Current - (*(const int *) (pi + L.p_Ti_begstep) + dt_step / 2)) * c.par.Timebase_interval;
	for(int k = 0; k < 3; k++)
	  {
	    float f = (float) (pf(c, P, i, L.p_Vel, k) + pf(c, P, i, L.p_GravAccel, k) * dt_kick);
	    if(type == 0) f = (float) (f + sf(c, S, i, L.s_HydroAccel, k) * dt_kick);	// the reference accumulates in the 4-byte buffer
	    out[k] = f;
	  }
	break;
      }
    case B_ID:
      memcpy(out, P + i * L.p_stride + c.IO.p_ID, 4);
      break;
    case B_MASS: out[0] = (float) pf(c, P, i, L.p_Mass); break;
    case B_U:
      {
	// io.c:320-324: u = A / (gamma-1) * rho^(gamma-1), floored
	const double u = sf(c, S, i, L.s_Entropy) / c.par.GammaMinus1 * pow(sf(c, S, i, L.s_Density), c.par.GammaMinus1);
	out[0] = (float) (c.par.MinEgySpec > u ? c.par.MinEgySpec : u);
	break;
      }
    case B_RHO: out[0] = (float) sf(c, S, i, L.s_Density); break;
    case B_NE: out[0] = (float) sf(c, S, i, c.IO.s_Ne); break;
    case B_NH: out[0] = nh ? (float) nh[i] : 0.0f; break;
    case B_HSML: out[0] = (float) pf(c, P, i, L.p_Hsml); break;
    case B_DUST: out[0] = (float) pf(c, P, i, c.IO.p_DustRadius); break;
    case B_SFR: out[0] = sfr ? (float) sfr[i] : 0.0f; break;
    case B_AGE: out[0] = (float) pf(c, P, i, c.IO.p_StellarAge); break;
    case B_BHMASS: out[0] = (float) pf(c, P, i, c.IO.p_BH_Mass); break;
    case B_BHMDOT: out[0] = (float) pf(c, P, i, c.IO.p_BH_Mdot); break;
    case B_ACCDISCMASS: out[0] = (float) pf(c, P, i, c.IO.p_AccDisc_Mass); break;
    case B_BHMSEED: out[0] = (float) pf(c, P, i, c.IO.p_BH_Mseed); break;
    case B_NEWDENS: out[0] = (float) pf(c, P, i, c.IO.p_NewDensity); break;
    case B_DELPHOTONMOM: out[0] = (float) pf(c, P, i, c.IO.p_DeltaPhotonMomentum); break;
    case B_OLDPHOTONMOM: out[0] = (float) pf(c, P, i, c.IO.p_OldPhotonMomentum); break;
    case B_BIPHOTONEN: out[0] = (float) pf(c, P, i, c.IO.p_BirthPhotonEnergy); break;
    case B_ACCMASS: out[0] = (float) sf(c, S, i, c.IO.s_accreted_mass); break;
    default: break;		// B_ACCMOM: see accmom_slot
    }
}

// ACMO (see the table): slot k of the record = component k % 3 of the gas particle k / 3
inline float accmom_slot(const Ctx &c, const char *S, const std::vector<long long> &gas, long long slot)
{
  return (float) sf(c, S, gas[(size_t) (slot / 3)], c.IO.s_accreted_momentum, (int) (slot % 3));
}

// empty_read_buffer() read_ic.c:238 for one particle
inline void value_in(const Ctx &c, Block b, char *P, char *S, long long i, int type, const float *in)
{
  const gl_gadget_layout &L = c.L;
  auto pset = [&](int off, int k, double v) { ((double *) (P + i * L.p_stride + off))[k] = v; };
  auto sset = [&](int off, int k, double v) { ((double *) (S + i * L.sph_stride + off))[k] = v; };
  switch (b)
    {
    case B_POS:
      for(int k = 0; k < 3; k++) pset(L.p_Pos, k, in[k]);
      *(short *) (P + i * L.p_stride + L.p_Type) = (short) type;	// read_ic.c:268-269
      break;
    case B_VEL:
      for(int k = 0; k < 3; k++) pset(L.p_Vel, k, in[k]);
      break;
    case B_ID: memcpy(P + i * L.p_stride + c.IO.p_ID, in, 4); break;
    case B_MASS: pset(L.p_Mass, 0, in[0]); break;
    case B_U: sset(L.s_Entropy, 0, in[0]); break;	// converted to entropy by init() once densities exist (read_ic.c:206)
    case B_RHO: sset(L.s_Density, 0, in[0]); break;
    case B_NE: sset(c.IO.s_Ne, 0, in[0]); break;
    case B_HSML: pset(L.p_Hsml, 0, in[0]); break;
    case B_DUST:
      if(c.par.RestartFlag == 2) pset(c.IO.p_DustRadius, 0, in[0]);	// read_ic.c:306-311
      break;
    case B_AGE: pset(c.IO.p_StellarAge, 0, in[0]); break;
    case B_BHMASS: pset(c.IO.p_BH_Mass, 0, in[0]); break;
    case B_BHMDOT: pset(c.IO.p_BH_Mdot, 0, in[0]); break;
    case B_ACCDISCMASS: pset(c.IO.p_AccDisc_Mass, 0, in[0]); break;
    case B_BHMSEED: pset(c.IO.p_BH_Mseed, 0, in[0]); break;
    case B_NEWDENS: pset(c.IO.p_NewDensity, 0, in[0]); break;
    case B_DELPHOTONMOM: pset(c.IO.p_DeltaPhotonMomentum, 0, in[0]); break;
    case B_OLDPHOTONMOM: pset(c.IO.p_OldPhotonMomentum, 0, in[0]); break;
    case B_BIPHOTONEN: pset(c.IO.p_BirthPhotonEnergy, 0, in[0]); break;
    case B_ACCMASS: sset(c.IO.s_accreted_mass, 0, in[0]); break;
    case B_ACCMOM:		// handled per slot by the caller (the record holds one value per particle, see the table)
    case B_NH: case B_SFR:	// derived records: "not needed to define the initial conditions" (read_ic.c:585-590)
    default: break;
    }
}

bool get(FILE *fd, void *p, size_t bytes) { return fread(p, 1, bytes, fd) == bytes; }

}				// namespace

extern "C" int gl_snapshot_write(const char *fname, const gl_gadget_layout *lay, const gl_io_layout *io, const gl_io_params *par,
				 const void *Pv, const void *SphPv, long long NumPart, const double *nh, const double *sfr,
				 char *err, int errlen)
{
  Ctx c{*lay, *io, *par, err, errlen};
  if(err && errlen > 0) err[0] = 0;
  if(par->format < 1 || par->format > 3) return c.fail("gl_snapshot_write: format %d is not 1, 2 or 3", par->format);
  const char *P = (const char *) Pv, *S = (const char *) SphPv;
  // format 3: the reference appends the suffix itself (io.c:2623)
  const std::string outname = par->format == 3 ? std::string(fname) + ".hdf5" : std::string(fname);
  fname = outname.c_str();

  const double T0 = gl_now();
  // particle indices per type, in P[] order (fill_write_buffer scans P[] for Type == type, io.c:212-216): every work
  // item lists its own stretch, the stretches are joined in order
  std::vector<long long> idx[6];
  {
    const long long STRETCH = 65536, nst = (NumPart + STRETCH - 1) / STRETCH;
    std::vector<std::vector<long long> > part((size_t) (nst * 6));
    std::atomic<long long> badtype{-1};
    HostPool::get().run(NumPart, STRETCH, [&](long long k, long long lo, long long hi) {
      for(long long i = lo; i < hi; i++)
	{
	  const int t = *(const short *) (P + i * c.L.p_stride + c.L.p_Type);
	  if(t < 0 || t > 5)
	    {
	      badtype.store(i);
	      return;
	    }
	  part[(size_t) (k * 6 + t)].push_back(i);
	}
    });
    if(badtype.load() >= 0) return c.fail("gl_snapshot_write: particle %lld has a Type outside 0..5", badtype.load());
    for(int t = 0; t < 6; t++)
      {
	size_t tot = 0;
	for(long long k = 0; k < nst; k++) tot += part[(size_t) (k * 6 + t)].size();
	idx[t].reserve(tot);
	for(long long k = 0; k < nst; k++) idx[t].insert(idx[t].end(), part[(size_t) (k * 6 + t)].begin(), part[(size_t) (k * 6 + t)].end());
      }
  }
  const double T1 = gl_now();
  gl_io_header h;
  memset(&h, 0, sizeof(h));
  static_assert(sizeof(gl_io_header) == 256, "struct io_header is 256 bytes (allvars.h:1733)");
  for(int t = 0; t < 6; t++)
    {
      h.npart[t] = (int) idx[t].size();	// io.c:2566-2574 (one file: totals = this file)
      h.npartTotal[t] = (unsigned int) idx[t].size();
      h.npartTotalHighWord[t] = (unsigned int) ((long long) idx[t].size() >> 32);
      h.mass[t] = par->MassTable[t];
    }
  h.time = par->Time;
  h.redshift = 0;		// non-comoving (io.c:2581-2584)
  h.flag_cooling = (par->build & GL_IO_BUILD_COOLING) ? 1 : 0;	// io.c:2589-2608
  if(par->build & GL_IO_BUILD_SFR)
    {
      h.flag_sfr = h.flag_feedback = 1;
      h.flag_stellarage = (par->build & GL_IO_BUILD_STELLARAGE) ? 1 : 0;
      h.flag_metals = (par->build & GL_IO_BUILD_METALS) ? 1 : 0;
    }
  h.num_files = 1;
  h.BoxSize = par->BoxSize;
  h.Omega0 = par->Omega0;
  h.OmegaLambda = par->OmegaLambda;
  h.HubbleParam = par->HubbleParam;
  h.flag_doubleprecision = 0;

  // Where everything goes is known before a byte is written: lay the file out first, then fill the payloads with ONE
  // sweep over the structs - a chunk of particles is converted for all its records while its cache lines are hot
  // (the reference makes one pass over P[] per record) - and write every piece at its offset.
  struct Piece
  {
    int row;
    long long off;		// file offset of this type's part of the record
  };
  std::vector<Piece> pieces[6];
  std::vector<std::pair<long long, int> > marks;	// (offset, value) of the 4-byte record markers
  std::vector<std::pair<long long, const char *> > labels;	// format 2
  std::vector<char> h5meta;	// format 3: everything in front of the raw data
  long long off = 0, header_off = -1;
  long long accmom_off = -1;
  if(par->format == 3)
    {
      // io.c:2619-2633: "/Header" with the header fields as attributes (write_header_attributes_in_hdf5, io.c:2866), one
      // group per type present, one dataset per (record, type): [npart[type]] or [npart[type], values] (io.c:2733-2748)
      std::vector<glh5::Group> groups;
      glh5::Group hg;
      hg.name = "Header";
      auto attr = [&](const char *name, glh5::Type ty, int count, const void *data) {
	glh5::Attr a{ name, ty, count, {} };
	a.data.assign((const char *) data, (const char *) data + (size_t) (count ? count : 1) * glh5::type_size(ty));
	hg.attrs.push_back(a);
      };
      attr("NumPart_ThisFile", glh5::T_I32, 6, h.npart);
      attr("NumPart_Total", glh5::T_U32, 6, h.npartTotal);
      attr("NumPart_Total_HighWord", glh5::T_U32, 6, h.npartTotalHighWord);
      attr("MassTable", glh5::T_F64, 6, h.mass);
      attr("Time", glh5::T_F64, 0, &h.time);
      attr("Redshift", glh5::T_F64, 0, &h.redshift);
      attr("BoxSize", glh5::T_F64, 0, &h.BoxSize);
      attr("NumFilesPerSnapshot", glh5::T_I32, 0, &h.num_files);
      attr("Omega0", glh5::T_F64, 0, &h.Omega0);
      attr("OmegaLambda", glh5::T_F64, 0, &h.OmegaLambda);
      attr("HubbleParam", glh5::T_F64, 0, &h.HubbleParam);
      attr("Flag_Sfr", glh5::T_I32, 0, &h.flag_sfr);
      attr("Flag_Cooling", glh5::T_I32, 0, &h.flag_cooling);
      attr("Flag_StellarAge", glh5::T_I32, 0, &h.flag_stellarage);
      attr("Flag_Metals", glh5::T_I32, 0, &h.flag_metals);
      attr("Flag_Feedback", glh5::T_I32, 0, &h.flag_feedback);
      attr("Flag_DoublePrecision", glh5::T_I32, 0, &h.flag_doubleprecision);
      groups.push_back(hg);
      int gidx[6];
      for(int t = 0; t < 6; t++)
	{
	  gidx[t] = -1;
	  if(h.npart[t] > 0)
	    {
	      glh5::Group g;
	      g.name = "PartType" + std::to_string(t);
	      gidx[t] = (int) groups.size();
	      groups.push_back(g);
	    }
	}
      struct Where { int type, row, group, dset; };
      std::vector<Where> where;
      for(int r = 0; r < B_COUNT; r++)
	{
	  const BlockRow &row = ROWS[r];
	  if(!present(c, row.b)) continue;
	  int typelist[6];
	  if(particles_in_block(c, row, h.npart, typelist) <= 0) continue;
	  for(int t = 0; t < 6; t++)
	    if(typelist[t])
	      {
		glh5::Dataset d;
		d.name = row.h5name;
		d.type = row.b == B_ID ? glh5::T_U32 : glh5::T_F32;	// get_datatype_in_block io.c:2245 (no LONGIDS, 4-byte output)
		d.rank = row.ncomp == 1 ? 1 : 2;
		d.dims[0] = (glh5::u64) h.npart[t];
		d.dims[1] = (glh5::u64) row.ncomp;
		where.push_back({ t, r, gidx[t], (int) groups[(size_t) gidx[t]].dsets.size() });
		groups[(size_t) gidx[t]].dsets.push_back(d);
	      }
	}
      glh5::Writer w(h5meta);
      off = (long long) w.layout(groups);
      if(off <= 0) return c.fail("gl_snapshot_write: more datasets in one group than the HDF5 writer lays out");
      for(const Where &q : where)
	{
	  const long long at = (long long) groups[(size_t) q.group].dsets[(size_t) q.dset].data_off;
	  if(ROWS[q.row].b == B_ACCMOM)
	    accmom_off = at;
	  else
	    pieces[q.type].push_back({ q.row, at });
	}
    }
  else
    {
  auto label_at = [&](const char *label, long long payload) {	// [8] LABEL nextblock [8] (io.c:2634-2642, :2672-2680)
    marks.push_back({off, 8});
    labels.push_back({off + 4, label});
    marks.push_back({off + 8, (int) (payload + 8)});
    marks.push_back({off + 12, 8});
    off += 16;
  };
  if(par->format == 2) label_at("HEAD", 256);
  header_off = off + 4;
  marks.push_back({off, 256});
  marks.push_back({off + 4 + 256, 256});
  off += 4 + 256 + 4;
  for(int r = 0; r < B_COUNT; r++)
    {
      const BlockRow &row = ROWS[r];
      if(!present(c, row.b)) continue;
      int typelist[6];
      const long long npart = particles_in_block(c, row, h.npart, typelist);
      if(npart <= 0) continue;
      const long long bytes = npart * row.ncomp * 4;
      if(bytes + 8 > 0x7fffffffLL)
	return c.fail("gl_snapshot_write: record %s needs %lld bytes, more than a 4-byte record marker holds", row.label, bytes);
      if(par->format == 2) label_at(row.label, bytes);
      marks.push_back({off, (int) bytes});
      off += 4;
      for(int t = 0; t < 6; t++)
	if(typelist[t])
	  {
	    if(row.b == B_ACCMOM)
	      accmom_off = off;	// gas only; written by slot, see the table
	    else
	      pieces[t].push_back({r, off});
	    off += (long long) h.npart[t] * row.ncomp * 4;
	  }
      marks.push_back({off, (int) bytes});
      off += 4;
    }
    }

  // The file is assembled in place: it is given its final length and mapped, and every work item converts its particles
  // straight into the mapping (concurrent write() calls on one file would queue on the inode lock).  Where the file
  // cannot be mapped (a pipe, an odd file system) the image is built in memory and written in one call.
  const long long total = off;
  // written under a temporary name and renamed when complete: a failure (full disk, quota) leaves the previous snapshot
  // of that name alone, where the reference would already have truncated it before endrun(777)
  const std::string tmpname = std::string(fname) + ".gl_tmp";
  const int fd = open(tmpname.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
  if(fd < 0) return c.fail("can't open file `%s' for writing snapshot.", fname);	// io.c:2626-2630 endrun(123)
  std::atomic<int> bad{0};
  char *image = nullptr;
  bool mapped = false;
  // the blocks are RESERVED before the file is mapped (posix_fallocate): on a full disk that fails here with an error
  // code instead of raising SIGBUS inside a worker when a page of the mapping cannot be backed
  if(ftruncate(fd, (off_t) total) == 0 && posix_fallocate(fd, 0, (off_t) total) == 0)
    {
      void *m = mmap(nullptr, (size_t) total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      if(m != MAP_FAILED)
	{
	  image = (char *) m;
	  mapped = true;
	}
    }
  if(!image) image = (char *) malloc((size_t) total);
  if(!image)
    {
      close(fd);
      return c.fail("gl_snapshot_write: no memory for a %lld-byte file image", total);
    }
  if(par->format == 3)
    memcpy(image, h5meta.data(), h5meta.size());
  else
    memcpy(image + header_off, &h, 256);
  for(auto &m : marks) memcpy(image + m.first, &m.second, 4);
  for(auto &l : labels) memcpy(image + l.first, l.second, 4);

  const double T2 = gl_now();
  const long long FUSE = 1024;	// particles per work item: 0.8 MB of structs, which stay in L2 across the 22 records
  for(int t = 0; t < 6; t++)
    {
      const std::vector<long long> &ix = idx[t];
      const std::vector<Piece> &pc = pieces[t];
      if(ix.empty() || pc.empty()) continue;
      HostPool::get().run((long long) ix.size(), FUSE, [&](long long, long long lo, long long hi) {
	for(const Piece &q : pc)
	  {
	    const BlockRow &row = ROWS[q.row];
	    float *out = (float *) (image + q.off) + lo * row.ncomp;
	    for(long long k = lo; k < hi; k++, out += row.ncomp) value_out(c, row.b, P, S, ix[(size_t) k], nh, sfr, out);
	  }
      });
    }
  if(accmom_off >= 0)
    {
      const std::vector<long long> &gas = idx[0];
      float *out = (float *) (image + accmom_off);
      HostPool::get().run((long long) gas.size(), FUSE, [&](long long, long long lo, long long hi) {
	for(long long k = lo; k < hi; k++) out[k] = accmom_slot(c, S, gas, k);
      });
    }
  const double T3 = gl_now();
  if(mapped)
    {
      if(msync(image, (size_t) total, MS_SYNC) != 0) bad.store(1);	// write-back errors are reported here, not lost
      if(munmap(image, (size_t) total) != 0) bad.store(1);
    }
  else
    {
      const char *q = image;
      long long left = total, where = 0;
      while(left > 0)
	{
	  const ssize_t w = pwrite(fd, q, (size_t) left, (off_t) where);
	  if(w <= 0)
	    {
	      bad.store(1);	// my_fwrite io.c:2988 endrun(777)
	      break;
	    }
	  q += w;
	  where += w;
	  left -= w;
	}
      free(image);
    }
  if(close(fd) != 0) bad.store(1);
  if(getenv("GL_DEBUG_IO")) fprintf(stderr, "snapshot_write: index %.2f ms, layout+open %.2f ms, payload %.2f ms, close %.2f ms\n", (T1 - T0) * 1e3, (T2 - T1) * 1e3, (T3 - T2) * 1e3, (gl_now() - T3) * 1e3);
  if(bad.load())
    {
      unlink(tmpname.c_str());
      return c.fail("gl_snapshot_write: I/O error on `%s'", fname);	// my_fwrite io.c:2988 endrun(777)
    }
  if(rename(tmpname.c_str(), fname) != 0)
    {
      unlink(tmpname.c_str());
      return c.fail("gl_snapshot_write: cannot rename the finished file to `%s'", fname);
    }
  return 0;
}

// the header record alone (read_file reads it first and sizes P[] / SphP[] from it: read_ic.c:688-697, :754-790)
static int read_header(Ctx &c, FILE *fd, int format, gl_io_header *h)
{
  int b1 = 0, b2 = 0, next = 0;
  char label[4];
  if(format == 2 && !(get(fd, &b1, 4) && get(fd, label, 4) && get(fd, &next, 4) && get(fd, &b2, 4)))
    return c.fail("incorrect header format");
  if(!(get(fd, &b1, 4) && get(fd, h, 256) && get(fd, &b2, 4)) || b1 != 256 || b2 != 256)
    return c.fail("incorrect header format");	// read_ic.c:699-704 endrun(890)
  return 0;
}

// format 3: struct io_header from the attributes of "/Header" (read_header_attributes_in_hdf5, read_ic.c:1333: the seven it
// reads; the flags are taken along when the file has them)
static void h5_header(const glh5::Reader &R, gl_io_header *h)
{
  memset(h, 0, sizeof(*h));
  const glh5::u64 hg = R.find("/Header");
  if(hg == glh5::UNDEF) throw glh5::Error{ "HDF5: no /Header group" };
  glh5::Value v;
  auto need = [&](const char *name, glh5::u64 count) {
    if(!R.attribute(hg, name, v)) throw glh5::Error{ std::string("HDF5: /Header has no attribute ") + name };
    if(v.count() < count) throw glh5::Error{ std::string("HDF5: /Header attribute too short: ") + name };
  };
  need("NumPart_ThisFile", 6);
  for(int t = 0; t < 6; t++) h->npart[t] = (int) v.as_u64((glh5::u64) t);
  need("NumPart_Total", 6);
  for(int t = 0; t < 6; t++) h->npartTotal[t] = (unsigned int) v.as_u64((glh5::u64) t);
  if(R.attribute(hg, "NumPart_Total_HighWord", v) && v.count() >= 6)
    for(int t = 0; t < 6; t++) h->npartTotalHighWord[t] = (unsigned int) v.as_u64((glh5::u64) t);
  need("MassTable", 6);
  for(int t = 0; t < 6; t++) h->mass[t] = v.as_double((glh5::u64) t);
  need("Time", 1);
  h->time = v.as_double(0);
  if(R.attribute(hg, "Redshift", v) && v.count()) h->redshift = v.as_double(0);
  need("BoxSize", 1);
  h->BoxSize = v.as_double(0);
  need("NumFilesPerSnapshot", 1);
  h->num_files = (int) v.as_u64(0);
  if(R.attribute(hg, "Omega0", v) && v.count()) h->Omega0 = v.as_double(0);
  if(R.attribute(hg, "OmegaLambda", v) && v.count()) h->OmegaLambda = v.as_double(0);
  if(R.attribute(hg, "HubbleParam", v) && v.count()) h->HubbleParam = v.as_double(0);
  struct { const char *name; int *dst; } flags[] = { { "Flag_Sfr", &h->flag_sfr }, { "Flag_Cooling", &h->flag_cooling }, { "Flag_StellarAge", &h->flag_stellarage },
    { "Flag_Metals", &h->flag_metals }, { "Flag_Feedback", &h->flag_feedback }, { "Flag_DoublePrecision", &h->flag_doubleprecision } };
  for(auto &f : flags)
    if(R.attribute(hg, f.name, v) && v.count()) *f.dst = (int) v.as_u64(0);
}
// a whole file in memory (mapped when possible)
struct FileImage
{
  const char *p = nullptr;
  size_t size = 0;
  void *mapping = MAP_FAILED;
  std::vector<char> copy;
  bool open(const char *fname)
  {
    FILE *fd = fopen(fname, "rb");
    if(!fd) return false;
    fseeko(fd, 0, SEEK_END);
    size = (size_t) ftello(fd);
    mapping = size ? mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fileno(fd), 0) : MAP_FAILED;
    if(mapping != MAP_FAILED)
      p = (const char *) mapping;
    else
      {
	copy.resize(size);
	fseeko(fd, 0, SEEK_SET);
	if(size && fread(copy.data(), 1, size, fd) != size)
	  {
	    fclose(fd);
	    return false;
	  }
	p = copy.data();
      }
    fclose(fd);
    return true;
  }
  ~FileImage()
  {
    if(mapping != MAP_FAILED) munmap(mapping, size);
  }
};

extern "C" int gl_snapshot_read_header(const char *fname, int format, gl_io_header *hdr, char *err, int errlen)
{
  Ctx c{};
  c.err = err;
  c.errlen = errlen;
  if(err && errlen > 0) err[0] = 0;
  if(format < 1 || format > 3) return c.fail("gl_snapshot_read_header: format %d is not 1, 2 or 3", format);
  if(format == 3)
    {
      FileImage img;
      if(!img.open(fname)) return c.fail("can't open file `%s' for reading initial conditions.", fname);
      try
	{
	  glh5::Reader R(img.p, img.size);
	  h5_header(R, hdr);
	}
      catch(const glh5::Error &e)
	{
	  return c.fail("%s (`%s')", e.what.c_str(), fname);
	}
      return 0;
    }
  FILE *fd = fopen(fname, "rb");
  if(!fd) return c.fail("can't open file `%s' for reading initial conditions.", fname);
  const int rc = read_header(c, fd, format, hdr);
  fclose(fd);
  return rc;
}

extern "C" int gl_snapshot_read(const char *fname, const gl_gadget_layout *lay, const gl_io_layout *io, gl_io_params *par, void *Pv,
				void *SphPv, long long MaxPart, long long MaxPartSph, long long *NumPart, long long *N_gas,
				gl_io_header *hdr, char *err, int errlen)
{
  Ctx c{*lay, *io, *par, err, errlen};
  if(err && errlen > 0) err[0] = 0;
  if(par->format < 1 || par->format > 3) return c.fail("gl_snapshot_read: format %d is not 1, 2 or 3", par->format);
  char *P = (char *) Pv, *S = (char *) SphPv;
  const bool h5 = par->format == 3;
  FILE *fd = h5 ? nullptr : fopen(fname, "rb");
  FileImage img;
  if(h5 ? !img.open(fname) : !fd) return c.fail("can't open file `%s' for reading initial conditions.", fname);	// read_ic.c:663-667 endrun(123)
  struct Closer
  {
    FILE *f;
    ~Closer()
    {
      if(f) fclose(f);
    }
  } closer{fd};

  int b1 = 0, b2 = 0, next = 0;
  char label[4];
  gl_io_header h;
  glh5::Reader *R = nullptr;
  struct RFree
  {
    glh5::Reader *&r;
    ~RFree() { delete r; }
  } rfree{R};
  if(h5)
    {
      try
	{
	  R = new glh5::Reader(img.p, img.size);
	  h5_header(*R, &h);
	}
      catch(const glh5::Error &e)
	{
	  return c.fail("%s (`%s')", e.what.c_str(), fname);
	}
    }
  else if(read_header(c, fd, par->format, &h))
    return 1;
  if(h.flag_doubleprecision)	// read_ic.c:744-750
    return c.fail("Problem: Code not compiled with INPUT_IN_DOUBLEPRECISION, but input files are in double precision!");
  if(h.num_files > 1) return c.fail("gl_snapshot_read: `%s' is one of %d files; multi-file snapshots are not supported", fname, h.num_files);
  for(int t = 0; t < 6; t++)
    {
      h.npartTotal[t] = (unsigned int) h.npart[t];	// read_ic.c:757-765 (num_files <= 1)
      if(par->build & GL_IO_BUILD_SFR) h.npartTotalHighWord[t] = 0;
      par->MassTable[t] = c.par.MassTable[t] = h.mass[t];	// read_ic.c:783-784
      if(h.npart[t] < 0) return c.fail("gl_snapshot_read: negative particle number for type %d", t);
    }
  long long nall = 0, first[7];
  for(int t = 0; t < 6; t++)
    {
      first[t] = nall;
      nall += h.npart[t];
    }
  first[6] = nall;
  if(h.npart[0] > MaxPartSph)	// read_ic.c:851-858 endrun(172)
    return c.fail("Not enough space for SPH particles (space for %lld, need at least %d)", MaxPartSph, h.npart[0]);
  if(nall > MaxPart)		// read_ic.c:864-869 endrun(173)
    return c.fail("Not enough space (space for %lld, need at least %lld)", MaxPart, nall);

  // Pass 1 (serial, a few reads): walk the record structure - labels, lengths; datasets in format 3 - and note where every
  // type's part of every record lies.  Pass 2: let every work item fill ALL fields of its particles in one go, so that
  // P[] / SphP[] are swept once instead of once per record; the per-particle rules that follow the records in
  // read_ic() (mass table, InitGasU, MinEgySpec) ride along.
  struct Piece
  {
    int row;
    long long off;		// formats 1 / 2: offset in the file
    const char *ptr;		// where the 4-byte values are
  };
  std::vector<Piece> pieces[6];
  long long accmom_off = -1;
  const char *accmom_ptr = nullptr;
  std::vector<std::vector<float> > converted;	// format 3: datasets stored in another type than the 4-byte one (H5Dread converts, read_ic.c:1020-1036)
  converted.reserve(6 * B_COUNT);
  for(int r = 0; r < B_COUNT; r++)
    {
      const BlockRow &row = ROWS[r];
      if(!present(c, row.b)) continue;
      if(par->RestartFlag == 0 && row.behind_u) continue;	// read_ic.c:891: "ignore all other blocks in initial conditions"
      int typelist[6];
      const long long npart = particles_in_block(c, row, h.npart, typelist);
      if(npart <= 0) continue;
      if(h5)
	{
	  for(int t = 0; t < 6; t++)
	    if(typelist[t])
	      {
		char path[96];
		snprintf(path, sizeof(path), "/PartType%d/%s", t, row.h5name);
		try
		  {
		    const glh5::u64 oh = R->find(path);
		    if(oh == glh5::UNDEF) return c.fail("gl_snapshot_read: `%s' has no dataset %s (H5Dopen fails in the reference, read_ic.c:998)", fname, path);
		    glh5::Value v;
		    R->dataset(oh, v);
		    const glh5::u64 want = (glh5::u64) h.npart[t] * (glh5::u64) row.ncomp;
		    if(v.rank < 1 || v.dims[0] != (glh5::u64) h.npart[t] || v.count() != want)
		      return c.fail("gl_snapshot_read: dataset %s does not hold %d x %d values", path, h.npart[t], row.ncomp);
		    const char *ptr;
		    const bool direct = v.size == 4 && (row.b == B_ID ? v.cls == 0 : v.cls == 1);
		    if(direct)
		      ptr = (const char *) v.data;
		    else
		      {
			// what H5Dread does with a memory type of NATIVE_FLOAT / NATIVE_UINT
			converted.emplace_back((size_t) want);
			std::vector<float> &buf = converted.back();
			HostPool::get().run((long long) want, 65536, [&](long long, long long lo, long long hi) {
			  for(long long k = lo; k < hi; k++)
			    {
			      if(row.b == B_ID)
				{
				  const unsigned int u = (unsigned int) v.as_u64((glh5::u64) k);
				  memcpy(&buf[(size_t) k], &u, 4);
				}
			      else
				buf[(size_t) k] = (float) v.as_double((glh5::u64) k);
			    }
			});
			ptr = (const char *) buf.data();
		      }
		    if(row.b == B_ACCMOM)
		      accmom_ptr = ptr;
		    else
		      pieces[t].push_back({r, 0, ptr});
		  }
		catch(const glh5::Error &e)
		  {
		    return c.fail("%s (%s in `%s')", e.what.c_str(), path, fname);
		  }
	      }
	  continue;
	}
      if(par->format == 2)
	{
	  if(!(get(fd, &b1, 4) && get(fd, label, 4) && get(fd, &next, 4) && get(fd, &b2, 4)))
	    return c.fail("gl_snapshot_read: `%s' ends before record %s", fname, row.label);
	  if(strncmp(label, row.label, 4) != 0)	// read_ic.c:936-946 endrun(1890)
	    return c.fail("incorrect block-structure! expected '%.4s' but found '%.4s'", row.label, label);
	}
      const long long bytes = npart * row.ncomp * 4;
      if(!get(fd, &b1, 4)) return c.fail("gl_snapshot_read: `%s' ends before record %s", fname, row.label);
      if(b1 != (int) bytes)
	return c.fail("gl_snapshot_read: record %s holds %d bytes, the header's particle numbers need %lld", row.label, b1, bytes);
      long long off = (long long) ftello(fd);
      for(int t = 0; t < 6; t++)
	if(typelist[t])
	  {
	    if(row.b == B_ACCMOM)
	      accmom_off = off;
	    else
	      pieces[t].push_back({r, off, nullptr});
	    off += (long long) h.npart[t] * row.ncomp * 4;
	  }
      if(fseeko(fd, (off_t) off, SEEK_SET) != 0 || !get(fd, &b2, 4) || b2 != b1)	// my_fread io.c:3005 endrun(778)
	return c.fail("gl_snapshot_read: record %s is cut short or not closed by its length (I/O error on `%s')", row.label, fname);
    }
  void *mapping = MAP_FAILED;
  long long used = 0;
  std::vector<char> copy;
  if(!h5)
    {
      used = (long long) ftello(fd);
      const char *image = nullptr;
      mapping = mmap(nullptr, (size_t) used, PROT_READ, MAP_PRIVATE, fileno(fd), 0);
      if(mapping != MAP_FAILED)
	image = (const char *) mapping;
      else
	{
	  copy.resize((size_t) used);
	  if(fseeko(fd, 0, SEEK_SET) != 0 || !get(fd, copy.data(), (size_t) used)) return c.fail("I/O error (fread) on `%s'", fname);
	  image = copy.data();
	}
      for(int t = 0; t < 6; t++)
	for(Piece &q : pieces[t]) q.ptr = image + q.off;
      if(accmom_off >= 0) accmom_ptr = image + accmom_off;
    }
  const bool ic_rules = par->RestartFlag == 0 && par->InitGasTemp > 0;	// read_ic.c:196-210
  for(int t = 0; t < 6; t++)
    {
      if(h.npart[t] == 0) continue;
      const std::vector<Piece> &pc = pieces[t];
      const double mass_t = par->MassTable[t];
      HostPool::get().run(h.npart[t], 1024, [&, t, mass_t](long long, long long lo, long long hi) {
	for(const Piece &q : pc)
	  {
	    const BlockRow &row = ROWS[q.row];
	    const float *in = (const float *) q.ptr + lo * row.ncomp;
	    for(long long k = lo; k < hi; k++, in += row.ncomp) value_in(c, row.b, P, S, first[t] + k, t, in);
	  }
	// read_ic.c:106-111: "this makes sure that masses are initialized in the case that the mass-block is completely empty"
	if(mass_t != 0)
	  for(long long k = lo; k < hi; k++) *(double *) (P + (first[t] + k) * c.L.p_stride + c.L.p_Mass) = mass_t;
	if(t == 0)
	  for(long long k = lo; k < hi; k++)
	    {
	      double *A = (double *) (S + k * c.L.sph_stride + c.L.s_Entropy);
	      if(ic_rules && *A == 0) *A = par->InitGasU;
	      if(par->MinEgySpec > *A) *A = par->MinEgySpec;	// read_ic.c:214-215
	    }
      });
    }
  if(accmom_ptr)
    {
      // the values the record really holds; the reference goes on to fill the momenta of the remaining particles from
      // whatever its I/O buffer held before (read_ic.c:555-562) - those are left untouched
      const float *in = (const float *) accmom_ptr;
      for(long long k = 0; k < h.npart[0]; k++)
	((double *) (S + (k / 3) * c.L.sph_stride + c.IO.s_accreted_momentum))[k % 3] = in[k];
    }
  if(mapping != MAP_FAILED) munmap(mapping, (size_t) used);
  if(par->RestartFlag == 0)
    {
      if(par->build & GL_IO_BUILD_BH_SWALLOWGAS) par->MassTable[5] = 0;	// read_ic.c:162-167
      if((par->build & GL_IO_BUILD_SFR) && par->MassTable[4] == 0 && par->MassTable[0] > 0)	// read_ic.c:169-178
	par->MassTable[0] = par->MassTable[4] = 0;
    }
  if(NumPart) *NumPart = nall;
  if(N_gas) *N_gas = h.npart[0];
  if(hdr) *hdr = h;
  return 0;
}

// Any numeric dataset or attribute of an HDF5 file as doubles: what the tests pin the reader with (a file the HDF5 library
// wrote), and a way to look into an initial-conditions file before P[] / SphP[] exist.  path = "/group/dataset", or
// "/group@attribute".  dims[4], *rank <- the shape; out (may be NULL) receives up to cap values, row-major.
extern "C" int gl_hdf5_read_f64(const char *fname, const char *path, double *out, long long cap, long long dims[4], int *rank, char *err,
				int errlen)
{
  Ctx c{};
  c.err = err;
  c.errlen = errlen;
  if(err && errlen > 0) err[0] = 0;
  FileImage img;
  if(!img.open(fname)) return c.fail("can't open file `%s'", fname);
  try
    {
      glh5::Reader R(img.p, img.size);
      glh5::Value v;
      std::string ps(path);
      const size_t atp = ps.find('@');
      if(atp != std::string::npos)
	{
	  const glh5::u64 oh = R.find(ps.substr(0, atp));
	  if(oh == glh5::UNDEF || !R.attribute(oh, ps.substr(atp + 1), v)) return c.fail("HDF5: no such attribute: %s", path);
	}
      else
	{
	  const glh5::u64 oh = R.find(ps);
	  if(oh == glh5::UNDEF) return c.fail("HDF5: no such object: %s", path);
	  R.dataset(oh, v);
	}
      if(rank) *rank = v.rank;
      if(dims)
	for(int k = 0; k < 4; k++) dims[k] = k < v.rank ? (long long) v.dims[k] : 0;
      const long long n = (long long) v.count();
      if(out)
	for(long long k = 0; k < n && k < cap; k++) out[k] = v.as_double((glh5::u64) k);
    }
  catch(const glh5::Error &e)
    {
      return c.fail("%s (`%s')", e.what.c_str(), fname);
    }
  return 0;
}
