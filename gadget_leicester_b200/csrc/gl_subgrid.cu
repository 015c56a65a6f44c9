// gl_subgrid.cu - the sub-grid neighbour loops that follow the force path, on the tree of the step
//
// accel.c:189-196 calls, right behind hydro_force(): blackhole_accretion() (blackhole.c:70) and, under -DDUST,
// dust_density() (dust.c:60) and dust_drag() (dust.c:263).  Each is a "search a sphere, then sum or mark" loop over
// ngb_treefind_blackhole (blackhole.c:1351) / ngb_treefind_dust_active (dust.c:1235).  Here they run on the array
// tree the force path built for this step:
//   * one warp per target (sink or dust grain); the warp marches the DFS entry array with the reference's cell test
//     (blackhole.c:1445-1463) and takes every cell of <= SG_BUCKET particles as one contiguous range of the sorted
//     particle arrays, 32 candidates per step - candidates reach the caller's code in the reference's Ngblist order
//     (tree order), which two statements of blackhole_evaluate depend on;
//   * what the reference resolves by visiting sinks one after the other is resolved by keys: "the last sink that
//     claims a dust grain or another sink gets it" (blackhole.c:991, :1030) = atomicMax over (position in the active
//     list, id); "the largest id gets the gas particle" (:1071) = atomicMax over the id;
//   * dust_evaluate_select (dust.c:889) SCATTERS onto gas neighbours and its entropy update does not commute
//     (clip at 1.5, floor MinEgySpec: dust.c:1000-1004): the (gas, dust) pairs are sorted by (gas index, position of
//     the grain in the active list) and applied per gas particle in that order - the reference's order.
// Single rank (the multi-rank build's cut cells are reported, not searched).
#include "gl_internal.h"
#include <cub/cub.cuh>

namespace
{
constexpr unsigned FULL = 0xffffffffu;
#define SG_BUCKET 96		// cells with at most this many particles are tested as one range (measured: 32 and 192 are both slower)
#define SG_GAMMA (7. / 5.)	// allvars.h:64
#define SG_GAMMA_MINUS1 (SG_GAMMA - 1)
#define SG_PI 3.14159265358979323846
#define SG_PROTONMASS 1.6726e-24	// allvars.h
#define SG_KC1 2.546479089470	// KERNEL_COEFF_1, _2, _5 (allvars.h:247-253)
#define SG_KC2 15.278874536822
#define SG_KC5 5.092958178941

struct SgTree
{
  int nentries;
  const int *e_next;
  const double4 *e_B;
  const int2 *e_rng;
  const double4 *s_pm;
  const int *s_type;
  const int *perm;
  // where a search may start instead of the root: the smallest cell around the target that holds its whole sphere
  const int *inv, *p_father, *n_parent, *n_first;
  const double4 *n_cl;
};

// The reference's neighbour search (blackhole.c:1351-1470, dust.c:1235-1330) as an ordered march: f(valid, p, type,
// mass, r2) is called by all 32 lanes for 32 consecutive tree-order particles at a time; valid = the particle is on the
// reference's Ngblist (type wanted, |d| <= h per axis, r2 <= h*h), r2 rounded as the reference's unfused
// dx*dx + dy*dy + dz*dz.
// The march covers the entries of the smallest cell around the target particle `self` (P[] index) that contains the
// whole search sphere - every neighbour lies in there, in the same order - instead of starting at the root: most of a
// root-to-leaf descent is a chain of dependent loads that finds nothing.
template <class F>
__device__ __forceinline__ void sg_march(const SgTree &T, int lane, int self, double px, double py, double pz, double h, unsigned typemask,
					 unsigned long long *counter, F f)
{
  const double h2 = __dmul_rn(h, h);
  int e = 0, e_end = T.nentries;
  {
    int no = __ldg(&T.p_father[__ldg(&T.inv[self])]);
    while(no >= 0)
      {
	const double4 cl = gl_ld4(&T.n_cl[no]);
	const double half = 0.5 * cl.w - h;
	if(fabs(px - cl.x) < half && fabs(py - cl.y) < half && fabs(pz - cl.z) < half) break;
	no = __ldg(&T.n_parent[no]);
      }
    if(no > 0)
      {
	e = __ldg(&T.n_first[no]) + no;
	e_end = __ldg(&T.e_next[e]);
      }
  }
  while(e < e_end)
    {
      // the three records of an entry are fetched together: the march is a chain of dependent L2 round trips and this
      // makes it one per entry instead of three
      const int nx = __ldg(&T.e_next[e]);
      const double4 B = gl_ld4(&T.e_B[e]);
      const int2 rg = __ldg(&T.e_rng[e]);
      int a, b;
      if(nx < 0)
	{
	  a = (-1 - nx) >> 3;
	  b = a + 1;
	  e = e + 1;
	}
      else
	{
	  double dist = h + 0.5 * B.w;
	  const double cx = fabs(B.x - px), cy = fabs(B.y - py), cz = fabs(B.z - pz);
	  bool skip = cx > dist || cy > dist || cz > dist;
	  if(!skip)
	    {
	      dist += 0.366025403785 * B.w;
	      skip = cx * cx + cy * cy + cz * cz > dist * dist;
	    }
	  if(skip)
	    {
	      e = nx;
	      continue;
	    }
	  if(rg.x < 0)
	    {
	      if(lane == 0) atomicMax(&counter[2], GL_ERR_LET_VIOLATION);
	      e = nx;
	      continue;
	    }
	  if(rg.y - rg.x > SG_BUCKET)
	    {
	      e = e + 1;
	      continue;
	    }
	  a = rg.x;
	  b = rg.y;
	  e = nx;
	}
      for(int q = a; q < b; q += 32)
	{
	  const int p = q + lane;
	  bool valid = p < b;
	  double r2 = 0, mj = 0;
	  int ty = -1;
	  if(valid)
	    {
	      ty = __ldg(&T.s_type[p]);
	      const double4 pm = gl_ld4(&T.s_pm[p]);
	      const double dx = px - pm.x, dy = py - pm.y, dz = pz - pm.z;
	      mj = pm.w;
	      r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
	      valid = ((typemask >> ty) & 1u) && !(fabs(dx) > h || fabs(dy) > h || fabs(dz) > h || r2 > h2);
	    }
	  f(valid, p, ty, mj, r2);
	}
    }
}

__device__ __forceinline__ double sg_kernel_wk(double u, double hinv3)
{
  // density.c:862-871 / blackhole.c:1046-1053 / dust.c:841-847
  if(u < 0.5) return hinv3 * (SG_KC1 + SG_KC2 * (u - 1) * u * u);
  return hinv3 * SG_KC5 * (1.0 - u) * (1.0 - u) * (1.0 - u);
}
__device__ __forceinline__ double sg_warp_sum(double v)
{
  for(int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double sg_dt(int timebin, double tbi) { return (timebin ? (double) (1 << timebin) : 0.0) * tbi; }

// ---- target lists: the active particles of one type, in active-list order ----
__global__ void k_sg_iota(long long n, int *out)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i < n) out[i] = (int) i;
}
__global__ void k_sg_flag_type(long long m, const int *__restrict__ active, const int *__restrict__ type, long long n, int want,
			       unsigned char *__restrict__ flag, int *__restrict__ bad)
{
  const long long k = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(k >= m) return;
  const int i = active[k];
  if(i < 0 || i >= n)
    {
      atomicAdd(bad, 1);
      flag[k] = 0;
      return;
    }
  flag[k] = type[i] == want;
}

// ---- sinks ----
struct SinkArr
{
  double *vel;
  const int *id;
  int *swallow;
  double *bh_mass, *bh_mdot, *accdisc, *dust_mass, *total_mass;
  const double *sinkrho;
  double *inj;
  const int *slot;		// compact mode: the arrays above hold only the particles the loops touch; slot[P index] = row
};
// row of particle j in the caller's arrays
#define SL(A, j) ((A).slot ? (A).slot[j] : (j))
#define SINK_SC 12		// per-sink scratch: dt, mdot, energy, acc_mass, acc_dust, mom[3], spare

// blackhole.c:181-214: the accretion-disc reservoir feeds the sink on the viscous time scale
__global__ void k_sink_mdot(int ns, const int *__restrict__ sinks, const int *__restrict__ timebin, const double *__restrict__ mass,
			    gl_subgrid_params par, double tbi, SinkArr A, double *__restrict__ sc)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if(s >= ns) return;
  const int i = sinks[s], is = SL(A, i);
  const double dt = sg_dt(timebin[i], tbi);
  const double mdot = A.accdisc[is] / (dt + (par.ViscousTimescale / par.UnitTime_in_s));
  const double dm = __dmul_rn(mdot, dt);
  A.accdisc[is] = __dadd_rn(A.accdisc[is], -dm);
  A.bh_mass[is] = __dadd_rn(A.bh_mass[is], dm);
  A.bh_mdot[is] = mdot;
  // blackhole.c:1139-1150: thermal feedback of a planet-mass sink (the SMBH-class sink injects nothing)
  const double m = mass[i];
  double energy = 0.;
  if(m < 0.95 * par.SMBHmass)
    energy = par.BlackHoleFeedbackFactor * pow(m * par.UnitMass_in_g, 0.6667) * 6.67e-8 * pow(4. * 3.1415 / 3. * 5., 0.3333) /
      par.UnitEnergy_in_cgs * mdot * par.UnitMass_in_g * dt;
  double *q = sc + (size_t) s * SINK_SC;
  q[0] = dt;
  q[1] = mdot;
  q[2] = energy;
  for(int k = 3; k < SINK_SC; k++) q[k] = 0;
}

// `r` of blackhole_evaluate is assigned for gas neighbours only (blackhole.c:1042) and read for dust neighbours (:1024).
// In the compiled reference it lives in a stack slot that survives from one call to the next, so a sink that meets dust
// before gas tests it with the distance of the last gas neighbour of the sink visited BEFORE it.  Reproduced: a first
// pass finds every sink's last gas distance, one thread hands them down the active list.  (Before any sink has met gas
// the reference reads whatever the stack held; the library uses 0 there.)
__global__ void __launch_bounds__(GL_BLOCK) k_sink_last_gas_r(int ns, const int *__restrict__ sinks, SgTree T, const double *__restrict__ in_pos,
							      const double *__restrict__ in_hsml, const double *__restrict__ in_mass,
							      double *__restrict__ sc, unsigned long long *__restrict__ counter)
{
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(s >= ns) return;
  const int i = sinks[s];
  double last = -1.;
  if(in_mass[i] > 0)
    {
      const double h = in_hsml[i], h2 = __dmul_rn(h, h);
      sg_march(T, lane, i, in_pos[3 * (size_t) i], in_pos[3 * (size_t) i + 1], in_pos[3 * (size_t) i + 2], h, 1u << 0, counter,
	       [&](bool valid, int, int, double mj, double r2) {
		 const bool isgas = valid && mj > 0 && r2 < h2;
		 const unsigned m = __ballot_sync(FULL, isgas);
		 if(m) last = __shfl_sync(FULL, sqrt(r2), 31 - __clz(m));
	       });
    }
  if(lane == 0) sc[(size_t) s * SINK_SC + 8] = last;
}
__global__ void k_sink_carry_r(int ns, double *__restrict__ sc)
{
  if(blockIdx.x || threadIdx.x) return;
  double run = 0.;
  for(int s = 0; s < ns; s++)
    {
      sc[(size_t) s * SINK_SC + 9] = run;
      if(sc[(size_t) s * SINK_SC + 8] >= 0) run = sc[(size_t) s * SINK_SC + 8];
    }
}

// blackhole_evaluate (blackhole.c:794-1199): one warp per active sink
__global__ void __launch_bounds__(GL_BLOCK) k_sink_evaluate(int ns, const int *__restrict__ sinks, SgTree T, const double *__restrict__ in_pos,
							    const double *__restrict__ in_hsml, const double *__restrict__ in_mass, gl_subgrid_params par,
							    SinkArr A, const double *__restrict__ sc, unsigned long long *__restrict__ claim,
							    unsigned long long *__restrict__ counter)
{
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(s >= ns) return;
  const int i = sinks[s];
  const double mass = in_mass[i];
  if(!(mass > 0)) return;	// blackhole.c:903
  const double px = in_pos[3 * (size_t) i], py = in_pos[3 * (size_t) i + 1], pz = in_pos[3 * (size_t) i + 2];
  const double h = in_hsml[i], h2 = __dmul_rn(h, h);
  const int is = SL(A, i);
  const double vx = A.vel[3 * (size_t) is], vy = A.vel[3 * (size_t) is + 1], vz = A.vel[3 * (size_t) is + 2];
  const int id = A.id[is];
  const double rho = A.sinkrho[is];
  const double energy = sc[(size_t) s * SINK_SC + 2];
  const bool smbh = mass > 0.95 * par.SMBHmass;
  const unsigned long long key = ((unsigned long long) (s + 1) << 32) | (unsigned int) id;
  const unsigned lt = (1u << lane) - 1;
  double r_stale = sc[(size_t) s * SINK_SC + 9];	// `r` of blackhole_evaluate as the previous sink left it (see k_sink_last_gas_r)
  sg_march(T, lane, i, px, py, pz, h, (1u << 0) | (1u << 2) | (1u << 5), counter, [&](bool valid, int p, int type, double mj, double r2) {
    const bool ok = valid && mj > 0 && r2 < h2;	// blackhole.c:901, :921
    const bool isgas = ok && type == 0;
    const double r = sqrt(r2);
    const unsigned gasmask = __ballot_sync(FULL, isgas);
    const unsigned below = gasmask & lt;
    const int src = below ? 31 - __clz(below) : 0;
    const double r_from = __shfl_sync(FULL, r, src);
    const double r_seen = below ? r_from : r_stale;
    if(gasmask) r_stale = __shfl_sync(FULL, r, 31 - __clz(gasmask));
    if(!ok) return;
    const int j = T.perm[p];
    if(type == 5)
      {
	// blackhole.c:956-1003 (-DBH_MERGERS_WITHIN_H): merge when closer than the accretion boundary
	if(r2 > 0)
	  {
	    const double accb = smbh ? par.InnerBoundary : par.SofteningBndry;
	    if(!(r > accb) && mj <= mass) atomicMax(&claim[j], key);
	  }
      }
    else if(type == 2)
      {
	// blackhole.c:1006-1035: dust on a bound orbit inside the accretion boundary
	if(r2 > 0)
	  {
	    const size_t js = (size_t) SL(A, j);
	    const double wx = A.vel[3 * js] - vx, wy = A.vel[3 * js + 1] - vy, wz = A.vel[3 * js + 2] - vz;
	    const double vrel = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(wx, wx), __dmul_rn(wy, wy)), __dmul_rn(wz, wz)));
	    const double accb = smbh ? par.InnerBoundary : par.SinkBoundary;
	    if(r < accb)
	      {
		const double etotal = __dadd_rn(__dmul_rn(vrel, vrel) / 2., -(mass / (r_seen + 1.e-20)));
		if(etotal <= 0.) atomicMax(&claim[j], key);
	      }
	  }
      }
    else
      {
	// blackhole.c:1038-1160
	const double hinv = 1 / h;
	const double hinv3 = hinv * hinv * hinv;
	const double wk = sg_kernel_wk(r * hinv, hinv3);
	const int js = SL(A, j);
	if(smbh && r < par.InnerBoundary) atomicMax(&A.swallow[js], id);	// :1068-1073
	if(energy != 0.) atomicAdd(&A.inj[js], energy * mj * wk / rho);	// :1154
      }
  });
}

__global__ void k_sink_apply_claims(long long m, const int *__restrict__ plist, const unsigned long long *__restrict__ claim, int *__restrict__ swallow)
{
  const long long t = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(t >= m) return;
  const unsigned long long c = claim[plist ? plist[t] : t];	// compact mode: row t holds particle plist[t]
  if(c) swallow[t] = (int) (unsigned int) (c & 0xffffffffull);
}

// blackhole_evaluate_swallow (blackhole.c:1201-1349): one warp per active sink that is not being swallowed itself
__global__ void __launch_bounds__(GL_BLOCK) k_sink_swallow(int ns, const int *__restrict__ sinks, SgTree T, const double *__restrict__ in_pos,
							   const double *__restrict__ in_hsml, double *__restrict__ in_mass, double4 *__restrict__ s_pm,
							   SinkArr A, double *__restrict__ sc, unsigned long long *__restrict__ counter)
{
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(s >= ns) return;
  const int i = sinks[s], is = SL(A, i);
  if(A.swallow[is] != 0) return;	// blackhole.c:508
  const double px = in_pos[3 * (size_t) i], py = in_pos[3 * (size_t) i + 1], pz = in_pos[3 * (size_t) i + 2];
  const double h = in_hsml[i];
  const int id = A.id[is];
  double am = 0, ad = 0, mx = 0, my = 0, mz = 0;
  int ng = 0, nb = 0, nd = 0;
  sg_march(T, lane, i, px, py, pz, h, (1u << 0) | (1u << 2) | (1u << 5), counter, [&](bool valid, int p, int type, double mj, double) {
    if(!valid) return;
    const int j = T.perm[p];
    const size_t js = (size_t) SL(A, j);
    if(A.swallow[js] != id) return;
    am += mj;
    mx += mj * A.vel[3 * js];
    my += mj * A.vel[3 * js + 1];
    mz += mj * A.vel[3 * js + 2];
    if(type == 5)
      {
	A.bh_mass[js] = 0;
	nb++;
      }
    else if(type == 2)
      {
	ad += mj;
	nd++;
      }
    else
      ng++;
    in_mass[j] = 0;
    s_pm[p].w = 0;
  });
  am = sg_warp_sum(am); ad = sg_warp_sum(ad); mx = sg_warp_sum(mx); my = sg_warp_sum(my); mz = sg_warp_sum(mz);
  ng = __reduce_add_sync(FULL, ng); nb = __reduce_add_sync(FULL, nb); nd = __reduce_add_sync(FULL, nd);
  if(lane == 0)
    {
      double *q = sc + (size_t) s * SINK_SC;
      q[3] = am; q[4] = ad; q[5] = mx; q[6] = my; q[7] = mz;
      if(ng) atomicAdd(&counter[4], (unsigned long long) ng);
      if(nb) atomicAdd(&counter[5], (unsigned long long) nb);
      if(nd) atomicAdd(&counter[6], (unsigned long long) nd);
    }
}

// blackhole.c:686-716: the sink takes mass and momentum of what it swallowed
__global__ void k_sink_finish(int ns, const int *__restrict__ sinks, const int *__restrict__ inv, double *__restrict__ in_mass,
			      double4 *__restrict__ s_pm, SinkArr A, const double *__restrict__ sc)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if(s >= ns) return;
  const int i = sinks[s];
  const double *q = sc + (size_t) s * SINK_SC;
  const double am = q[3], ad = q[4];
  if(am > 0)
    {
      const double m = in_mass[i];
      const size_t is = (size_t) SL(A, i);
      for(int k = 0; k < 3; k++)
	A.vel[3 * is + k] = __dadd_rn(__dmul_rn(A.vel[3 * is + k], m), q[5 + k]) / __dadd_rn(m, am);
      const double mnew = __dadd_rn(m, am);
      in_mass[i] = mnew;
      s_pm[inv[i]].w = mnew;
      // BH_accreted_BHMass and BH_accreted_DustMass are ONE union member (allvars.h:1261-1267); the dust mass is stored
      // last (blackhole.c:1328-1329), so both sums below receive it (blackhole.c:699-701)
      A.bh_mass[is] = __dadd_rn(A.bh_mass[is], ad);
      A.dust_mass[is] = __dadd_rn(A.dust_mass[is], ad);
      A.accdisc[is] = __dadd_rn(A.accdisc[is], am);
      A.total_mass[is] = __dadd_rn(A.total_mass[is], am);
    }
}

// ---- dust ----
// dust_evaluate_density (dust.c:748-887): sum over the Type-2 neighbours inside the grain's own Hsml
__global__ void __launch_bounds__(GL_BLOCK) k_dust_density(int nd, const int *__restrict__ dust, SgTree T, const double *__restrict__ in_pos,
							   const double *__restrict__ in_hsml, const double *__restrict__ in_mass,
							   double *__restrict__ out, bool by_target, unsigned long long *__restrict__ counter)
{
  const int d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(d >= nd) return;
  const int i = dust[d];
  const double px = in_pos[3 * (size_t) i], py = in_pos[3 * (size_t) i + 1], pz = in_pos[3 * (size_t) i + 2];
  const double h = in_hsml[i], mi = in_mass[i];
  double sum = 0;
  sg_march(T, lane, i, px, py, pz, h, 1u << 2, counter, [&](bool valid, int, int, double mj, double r2) {
    if(!valid || !(mj > 0)) return;
    const double u = sqrt(r2) / h;
    if(u < 1)
      {
	const double hinv = 1 / h;
	sum += mi * sg_kernel_wk(u, hinv * hinv * hinv);
      }
  });
  sum = sg_warp_sum(sum);
  if(lane == 0) out[by_target ? d : i] = sum;
}

struct DustArr
{
  double *vel;
  const double *gravaccel, *rho, *entr, *gasvel, *d7;
  double *d9;
  const double *radius;
  double *vcoll, *ddm, *newdrag, *dde;
  double *entropy, *dragheat;
  const int *slot;
};
#define DUST_SC 6		// per-grain scratch: DeltaDustMomentum[3] / density, DeltaDragEnergy, density, spare

// the grain's own update, dust_drag() dust.c:340-452: stopping time in the Epstein or Stokes regime, then the exact
// solution of dv/dt = -(v - v_steady)/ts + g over the grain's time step
__global__ void k_dust_own(int nd, const int *__restrict__ dust, const int *__restrict__ timebin, const double *__restrict__ in_pos,
			   const double *__restrict__ in_mass, gl_subgrid_params par, double tbi, DustArr A, double *__restrict__ sc)
{
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if(d >= nd) return;
  const int i = dust[d];
  const size_t ir = (size_t) SL(A, i), i3 = 3 * ir;
  for(int k = 0; k < 3; k++)
    {
      A.newdrag[i3 + k] = 0.;
      A.ddm[i3 + k] = 0.;
    }
  double dde = 0.;
  const double dt = sg_dt(timebin[i], tbi);
  const double rho = A.rho[ir];
  const double soundspeed = sqrt(8. / SG_PI * A.entr[ir] * pow(rho, SG_GAMMA_MINUS1));
  const double v0 = A.vel[i3], v1 = A.vel[i3 + 1], v2 = A.vel[i3 + 2];
  const double g0 = v0 - A.gasvel[i3], g1 = v1 - A.gasvel[i3 + 1], g2 = v2 - A.gasvel[i3 + 2];
  const double delta_vel = sqrt(g0 * g0 + g1 * g1 + g2 * g2);
  const double d7 = A.d7[ir];
  double vcoll = A.vcoll[ir];
  if(d7 > 0.)
    {
      double w[3] = { 0, 0, 0 };
      if(A.d9)
	for(int k = 0; k < 3; k++)
	  {
	    w[k] = A.d9[i3 + k] / d7;
	    A.d9[i3 + k] = w[k];
	  }
      const double delta_peb_vel = sqrt((v0 - w[0]) * (v0 - w[0]) + (v1 - w[1]) * (v1 - w[1]) + (v2 - w[2]) * (v2 - w[2]));
      vcoll = delta_peb_vel * par.UnitVelocity_in_cm_per_s / 1.e2 + 1.e-30;
    }
  const double rho_dust = 3.;
  const double a = A.radius[ir];
  const double mi = in_mass[i];
  double ddm[3] = { 0, 0, 0 };
  if(dt > 0)
    {
      const double lambda_h2 = par.MeanWeight * SG_PROTONMASS / (par.UnitDensity_in_cgs * rho) / 1.e-15 / par.UnitLength_in_cm;
      const double rey = 6 * delta_vel * a / par.UnitLength_in_cm / (lambda_h2 * soundspeed);
      double ts;
      if(3. / 2 * lambda_h2 * par.UnitLength_in_cm >= a)
	ts = 1. / (rho * soundspeed / (rho_dust * a) * par.UnitMass_in_g / par.UnitLength_in_cm / par.UnitLength_in_cm);	// Epstein
      else if(delta_vel > 0)
	{
	  double C_drag = 0.44;	// Stokes
	  if(rey < 800. && rey >= 1.) C_drag = 24. * pow(rey, -0.6);
	  if(rey < 1.) C_drag = 24. / rey;
	  ts = rho_dust * a / (rho * delta_vel) / par.UnitMass_in_g * par.UnitLength_in_cm * par.UnitLength_in_cm;
	  ts *= 8. / 3. / C_drag;
	}
      else
	ts = 0.66667 / (rho * soundspeed / (rho_dust * a)) / par.UnitMass_in_g * par.UnitLength_in_cm * a / lambda_h2;
      const double ex = exp(-dt / ts), ex2 = exp(-2. * dt / ts);
      double gacc2 = 0;
      for(int k = 0; k < 3; k++)
	{
	  const double vold = A.vel[i3 + k], ga = A.gravaccel[i3 + k];
	  const double vsteady = (vold * d7 + A.gasvel[i3 + k] * rho) / (d7 + rho + 1.e-30);
	  const double vnew = vsteady + (vold - vsteady) * ex + ga * ts * (1. - ex);
	  A.vel[i3 + k] = vnew;
	  const double vnew_nog = vold + ga * dt;
	  ddm[k] -= mi * (vnew_nog - vnew);
	  dde += mi * (vnew - vsteady) * (vnew - vsteady) * (1. - ex2) / 2.;
	  gacc2 += ga * ga;
	}
      vcoll = sqrt(gacc2) * ts * par.UnitVelocity_in_cm_per_s / 1.e2;
      vcoll += 0.2;
    }
  for(int k = 0; k < 3; k++) A.ddm[i3 + k] = ddm[k];
  A.dde[ir] = dde;
  A.vcoll[ir] = vcoll;
  double *q = sc + (size_t) d * DUST_SC;
  q[0] = ddm[0]; q[1] = ddm[1]; q[2] = ddm[2]; q[3] = dde; q[4] = rho; q[5] = 0;
}

// dust_evaluate_select (dust.c:889-1029), search half: the gas inside the grain's Hsml as (key, value) pairs,
// key = (gas P[] index, position of the grain in the active list), value = wk.  The pairs are sorted afterwards, so they
// may land anywhere: a warp parks its hits in shared memory and reserves room in the global list once per DP_BUF hits
// (one pass, one atomic per grain in the common case).  cursor[0] counts every pair, also those beyond cap: the caller
// grows the list and repeats when it overflowed.
#define DP_BUF 128
__global__ void __launch_bounds__(GL_BLOCK) k_dust_pairs(int nd, const int *__restrict__ dust, SgTree T, const double *__restrict__ in_pos,
							 const double *__restrict__ in_hsml, const double *__restrict__ sc, unsigned long long *__restrict__ cursor,
							 long long cap, unsigned long long *__restrict__ keys, double *__restrict__ vals,
							 unsigned long long *__restrict__ counter)
{
  __shared__ unsigned long long s_k[GL_WARPS_PER_BLOCK][DP_BUF];
  __shared__ double s_v[GL_WARPS_PER_BLOCK][DP_BUF];
  const int d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(d >= nd) return;
  const int i = dust[d];
  if(!(sc[(size_t) d * DUST_SC + 4] > 0)) return;	// dust.c:995: nothing is returned where the gas density at the grain is zero
  unsigned long long *bk = s_k[threadIdx.x >> 5];
  double *bv = s_v[threadIdx.x >> 5];
  const double px = in_pos[3 * (size_t) i], py = in_pos[3 * (size_t) i + 1], pz = in_pos[3 * (size_t) i + 2];
  const double h = in_hsml[i], hinv = 1 / h, hinv3 = hinv * hinv * hinv;
  const unsigned lt = (1u << lane) - 1;
  int nbuf = 0;
  auto flush = [&]() {
    __syncwarp();
    unsigned long long base = 0;
    if(lane == 0) base = atomicAdd(cursor, (unsigned long long) nbuf);
    base = __shfl_sync(FULL, base, 0);
    for(int k = lane; k < nbuf; k += 32)
      if((long long) (base + k) < cap)
	{
	  keys[base + k] = bk[k];
	  vals[base + k] = bv[k];
	}
    nbuf = 0;
    __syncwarp();
  };
  sg_march(T, lane, i, px, py, pz, h, 1u << 0, counter, [&](bool valid, int p, int, double mj, double r2) {
    const double u = sqrt(r2) / h;
    const bool hit = valid && mj > 0 && u < 1;
    const unsigned m = __ballot_sync(FULL, hit);
    if(!m) return;
    if(nbuf + __popc(m) > DP_BUF) flush();
    if(hit)
      {
	const int at = nbuf + __popc(m & lt);
	bk[at] = ((unsigned long long) (unsigned int) T.perm[p] << 32) | (unsigned int) d;
	bv[at] = sg_kernel_wk(u, hinv3);
      }
    nbuf += __popc(m);
  });
  if(nbuf) flush();
}

// dust.c:988-1007 per gas particle, contributions in the order the reference applies them
__global__ void k_dust_apply(long long npairs, const unsigned long long *__restrict__ keys, const double *__restrict__ vals,
			     const double *__restrict__ sc, const int *__restrict__ timebin, const double *__restrict__ in_mass,
			     gl_subgrid_params par, double tbi, DustArr A)
{
  const long long t = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(t >= npairs) return;
  const unsigned int j = (unsigned int) (keys[t] >> 32);
  if(t > 0 && (unsigned int) (keys[t - 1] >> 32) == j) return;
  const double dt = sg_dt(timebin[j], tbi);
  if(!(dt > 0.)) return;
  const double mj = in_mass[j];
  const size_t js = (size_t) SL(A, j);
  double v0 = A.vel[3 * js], v1 = A.vel[3 * js + 1], v2 = A.vel[3 * js + 2];
  double S = A.entropy[js], heat = A.dragheat[js];
  for(long long q = t; q < npairs && (unsigned int) (keys[q] >> 32) == j; q++)
    {
      const double *g = sc + (size_t) (unsigned int) (keys[q] & 0xffffffffull) * DUST_SC;
      const double wk = vals[q], dens = g[4], ret = g[3];
      v0 -= g[0] / dens * wk;
      v1 -= g[1] / dens * wk;
      v2 -= g[2] / dens * wk;
      const double u_old = fmax(par.MinEgySpec, S / SG_GAMMA_MINUS1 * pow(dens, SG_GAMMA_MINUS1));
      const double u_new = u_old + (ret * wk / dens);
      double u_inc = u_new / u_old;
      if(u_inc > 1.5) u_inc = 1.5;
      S *= u_inc;
      heat += 1.e-20 * (ret * wk / dens * mj / dt);
    }
  A.vel[3 * js] = v0;
  A.vel[3 * js + 1] = v1;
  A.vel[3 * js + 2] = v2;
  A.entropy[js] = S;
  A.dragheat[js] = heat;
}

// ---- host side ----
// a caller array on the device: itself when it already lives there, a staged copy otherwise
template <class T> struct Staged
{
  T *dev = nullptr;
  T *host = nullptr;
  size_t cnt = 0;
  bool owned = false, out = false;
  int bind(gl_ctx *c, const T *src, size_t n, int where, bool in, bool writeback)
  {
    cnt = n;
    out = writeback;
    if(!src) return 0;
    if(where)
      {
	dev = const_cast<T *>(src);
	return 0;
      }
    host = const_cast<T *>(src);
    owned = true;
    GL_CUDA(c, cudaMalloc((void **) &dev, n * sizeof(T)));
    if(in)
      GL_CUDA(c, cudaMemcpyAsync(dev, src, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    else
      GL_CUDA(c, cudaMemsetAsync(dev, 0, n * sizeof(T), c->stream));
    return 0;
  }
  int finish(gl_ctx *c)
  {
    if(owned && out && dev) GL_CUDA(c, cudaMemcpyAsync(host, dev, cnt * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
    return 0;
  }
  ~Staged()
  {
    if(owned && dev) cudaFree(dev);
  }
};

int sg_ready(gl_ctx *c, const char *who)
{
  cudaSetDevice(c->device);
  if(!c->have_par) return c->fail(std::string(who) + ": gl_set_params first");
  if(!c->have_subgrid) return c->fail(std::string(who) + ": gl_set_subgrid_params first");
  if(c->comm && c->comm->world > 1) return c->fail(std::string(who) + ": single rank only (no sub-grid loops over a locally-essential tree)");
  if(!c->tree_valid || c->nentries <= 0) return c->fail(std::string(who) + ": no tree - call gl_gravity_tree / gl_density of this step first");
  if(!c->in_timebin.p || !c->in_hsml.p) return c->fail(std::string(who) + ": time bins and Hsml must be loaded");
  GL_CUDA(c, c->d_counter.ensure(16));
  GL_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, 16 * sizeof(unsigned long long), c->stream));
  return 0;
}
SgTree sg_tree(gl_ctx *c)
{
  return SgTree{ (int) c->nentries, c->e_next.p, c->e_B.p, c->e_rng.p, c->s_pm.p, c->s_type.p, c->perm.p,
		 c->inv.p, c->p_father.p, c->n_parent.p, c->n_first.p, c->n_cl.p };
}

// the active particles of type `want`, in active-list order -> c->sg_list (device), *count
int sg_targets(gl_ctx *c, const int *active, long long nactive, int want, long long *count)
{
  const long long n = c->n;
  const long long m = active ? nactive : n;
  *count = 0;
  if(m <= 0) return 0;
  GL_CUDA(c, c->sg_active.ensure((size_t) m));
  GL_CUDA(c, c->sg_list.ensure((size_t) m + 1));
  GL_CUDA(c, c->sg_flag.ensure((size_t) m));
  if(active)
    GL_CUDA(c, cudaMemcpyAsync(c->sg_active.p, active, m * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  else
    {
      k_sg_iota<<<gl_grid(m, 256), 256, 0, c->stream>>>(m, c->sg_active.p);
      GL_LAUNCH_COUNT(c);
    }
  int *bad = (int *) (c->d_counter.p + 8), *nsel = (int *) (c->d_counter.p + 9);
  k_sg_flag_type<<<gl_grid(m, 256), 256, 0, c->stream>>>(m, c->sg_active.p, c->in_type.p, n, want, c->sg_flag.p, bad);
  GL_LAUNCH_COUNT(c);
  size_t tmp = 0;
  cub::DeviceSelect::Flagged(nullptr, tmp, c->sg_active.p, c->sg_flag.p, c->sg_list.p, nsel, (int) m, c->stream);
  GL_CUDA(c, c->cub_tmp.ensure(tmp));
  GL_CUDA(c, cub::DeviceSelect::Flagged(c->cub_tmp.p, tmp, c->sg_active.p, c->sg_flag.p, c->sg_list.p, nsel, (int) m, c->stream));
  GL_LAUNCH_COUNT(c);
  unsigned long long hc[2];
  GL_CUDA(c, cudaMemcpyAsync(hc, c->d_counter.p + 8, sizeof(hc), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  if((int) hc[0] != 0) return c->fail("sub-grid loop: active list holds an index outside [0, n)");
  *count = (int) hc[1];
  return 0;
}

// ---- compact mode (the AoS drop-in): only the particles the loops touch travel ----
// mark the targets and everything on their neighbour lists
__global__ void __launch_bounds__(GL_BLOCK) k_sg_mark(int nt, const int *__restrict__ targets, SgTree T, const double *__restrict__ in_pos,
						      const double *__restrict__ in_hsml, unsigned typemask, int *__restrict__ flag,
						      unsigned long long *__restrict__ counter)
{
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(t >= nt) return;
  const int i = targets[t];
  if(lane == 0) flag[i] = 1;
  if(!typemask) return;
  const double px = in_pos[3 * (size_t) i], py = in_pos[3 * (size_t) i + 1], pz = in_pos[3 * (size_t) i + 2];
  sg_march(T, lane, i, px, py, pz, in_hsml[i], typemask, counter, [&](bool valid, int p, int, double, double) {
    if(valid) flag[T.perm[p]] = 1;
  });
}
__global__ void k_sg_list(long long n, const int *__restrict__ flag, const int *__restrict__ slot, int *__restrict__ plist)
{
  const long long j = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(j < n && flag[j]) plist[slot[j]] = (int) j;
}
__global__ void k_sg_gather(long long m, const int *__restrict__ plist, const double *__restrict__ src, double *__restrict__ dst, int ncomp)
{
  const long long t = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(t >= m * ncomp) return;
  dst[t] = src[(size_t) plist[t / ncomp] * ncomp + t % ncomp];
}

// targets (c->sg_list, nt of them) + their neighbours of the given types -> c->sg_plist (ascending P[] indices, *m of
// them, also copied to `host`) and c->sg_slot (P index -> row)
int sg_compact(gl_ctx *c, long long nt, unsigned typemask, long long *m, std::vector<int> &host)
{
  const long long n = c->n;
  GL_CUDA(c, c->sg_pflag.ensure((size_t) n + 1));
  GL_CUDA(c, c->sg_slot.ensure((size_t) n + 1));
  GL_CUDA(c, cudaMemsetAsync(c->sg_pflag.p, 0, ((size_t) n + 1) * sizeof(int), c->stream));
  k_sg_mark<<<gl_grid(nt * 32, GL_BLOCK), GL_BLOCK, 0, c->stream>>>((int) nt, c->sg_list.p, sg_tree(c), c->in_pos.p, c->in_hsml.p, typemask, c->sg_pflag.p,
								     c->d_counter.p);
  GL_LAUNCH_COUNT(c);
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->sg_pflag.p, c->sg_slot.p, (int) n + 1, c->stream);
  GL_CUDA(c, c->cub_tmp.ensure(tmp));
  GL_CUDA(c, cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, c->sg_pflag.p, c->sg_slot.p, (int) n + 1, c->stream));
  GL_LAUNCH_COUNT(c);
  int mm = 0;
  GL_CUDA(c, cudaMemcpyAsync(&mm, c->sg_slot.p + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  *m = mm;
  GL_CUDA(c, c->sg_plist.ensure((size_t) mm + 1));
  k_sg_list<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, c->sg_pflag.p, c->sg_slot.p, c->sg_plist.p);
  GL_LAUNCH_COUNT(c);
  host.resize((size_t) mm);
  GL_CUDA(c, cudaMemcpyAsync(host.data(), c->sg_plist.p, (size_t) mm * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// the five kernels of blackhole_accretion on ns sinks (c->sg_list); rows = A's arrays: all n particles (plist null) or
// the m listed ones
int sg_run_sinks(gl_ctx *c, long long ns, const SinkArr &A, const int *plist, long long m)
{
  if(ns <= 0) return 0;
  const size_t n = (size_t) c->n;
  GL_CUDA(c, c->sg_scratch.ensure((size_t) ns * SINK_SC));
  GL_CUDA(c, c->sg_claim.ensure(n));
  GL_CUDA(c, cudaMemsetAsync(c->sg_claim.p, 0, n * sizeof(unsigned long long), c->stream));
  const SgTree T = sg_tree(c);
  const double tbi = c->par.Timebase_interval;
  const int wgrid = gl_grid(ns * 32, GL_BLOCK);
  k_sink_mdot<<<gl_grid(ns, 128), 128, 0, c->stream>>>((int) ns, c->sg_list.p, c->in_timebin.p, c->in_mass.p, c->subgrid, tbi, A, c->sg_scratch.p);
  k_sink_last_gas_r<<<wgrid, GL_BLOCK, 0, c->stream>>>((int) ns, c->sg_list.p, T, c->in_pos.p, c->in_hsml.p, c->in_mass.p, c->sg_scratch.p, c->d_counter.p);
  k_sink_carry_r<<<1, 32, 0, c->stream>>>((int) ns, c->sg_scratch.p);
  k_sink_evaluate<<<wgrid, GL_BLOCK, 0, c->stream>>>((int) ns, c->sg_list.p, T, c->in_pos.p, c->in_hsml.p, c->in_mass.p, c->subgrid, A, c->sg_scratch.p,
						      c->sg_claim.p, c->d_counter.p);
  k_sink_apply_claims<<<gl_grid(m, 256), 256, 0, c->stream>>>(m, plist, c->sg_claim.p, A.swallow);
  k_sink_swallow<<<wgrid, GL_BLOCK, 0, c->stream>>>((int) ns, c->sg_list.p, T, c->in_pos.p, c->in_hsml.p, c->in_mass.p, c->s_pm.p, A, c->sg_scratch.p,
						     c->d_counter.p);
  k_sink_finish<<<gl_grid(ns, 128), 128, 0, c->stream>>>((int) ns, c->sg_list.p, c->inv.p, c->in_mass.p, c->s_pm.p, A, c->sg_scratch.p);
  c->kernel_launches += 7;
  GL_CUDA(c, cudaGetLastError());
  return 0;
}

// dust_drag on nd grains (c->sg_list)
int sg_run_dust(gl_ctx *c, long long nd, const DustArr &A)
{
  if(nd <= 0) return 0;
  const size_t n = (size_t) c->n;
  const SgTree T = sg_tree(c);
  const double tbi = c->par.Timebase_interval;
  GL_CUDA(c, c->sg_scratch.ensure((size_t) nd * DUST_SC));
  const int wgrid = gl_grid(nd * 32, GL_BLOCK);
  k_dust_own<<<gl_grid(nd, 128), 128, 0, c->stream>>>((int) nd, c->sg_list.p, c->in_timebin.p, c->in_pos.p, c->in_mass.p, c->subgrid, tbi, A, c->sg_scratch.p);
  c->kernel_launches += 1;
  unsigned long long *cursor = c->d_counter.p + 10;
  long long npairs = 0;
  unsigned long long *k0 = nullptr, *k1 = nullptr;
  double *v0 = nullptr, *v1 = nullptr;
  long long cap = std::max<long long>(64 * nd, c->sg_pairs_seen + c->sg_pairs_seen / 4);
  for(int attempt = 0; attempt < 2; attempt++)
    {
      GL_CUDA(c, c->sg_keys.ensure(2 * (size_t) cap));
      GL_CUDA(c, c->sg_vals.ensure(2 * (size_t) cap));
      k0 = c->sg_keys.p; k1 = c->sg_keys.p + cap;
      v0 = c->sg_vals.p; v1 = c->sg_vals.p + cap;
      GL_CUDA(c, cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), c->stream));
      k_dust_pairs<<<wgrid, GL_BLOCK, 0, c->stream>>>((int) nd, c->sg_list.p, T, c->in_pos.p, c->in_hsml.p, c->sg_scratch.p, cursor, cap, k0, v0, c->d_counter.p);
      GL_LAUNCH_COUNT(c);
      unsigned long long hc = 0;
      GL_CUDA(c, cudaMemcpyAsync(&hc, cursor, sizeof(hc), cudaMemcpyDeviceToHost, c->stream));
      GL_CUDA(c, cudaStreamSynchronize(c->stream));
      npairs = (long long) hc;
      c->sg_pairs_seen = npairs;
      if(npairs <= cap) break;
      if(attempt == 1) return c->fail("gl_dust_drag: pair list overflowed twice");
      cap = npairs;
    }
  if(npairs >= (1ll << 31)) return c->fail("gl_dust_drag: more than 2^31 (grain, gas) pairs");
  if(npairs > 0)
    {
      int top = 33;	// bits of the key that vary: 32 of the grain's position + those of the gas index
      for(long long q = (long long) n; q > 1; q >>= 1) top++;
      if(top > 64) top = 64;
      size_t tmp = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, tmp, k0, k1, v0, v1, (int) npairs, 0, top, c->stream);
      GL_CUDA(c, c->cub_tmp.ensure(tmp));
      GL_CUDA(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, k0, k1, v0, v1, (int) npairs, 0, top, c->stream));
      GL_LAUNCH_COUNT(c);
      k_dust_apply<<<gl_grid(npairs, 256), 256, 0, c->stream>>>(npairs, k1, v1, c->sg_scratch.p, c->in_timebin.p, c->in_mass.p, c->subgrid, tbi, A);
      GL_LAUNCH_COUNT(c);
    }
  GL_CUDA(c, cudaGetLastError());
  return 0;
}
int sg_counts(gl_ctx *c, const char *who, long long n_swallowed[3])
{
  unsigned long long hc[8];
  GL_CUDA(c, cudaMemcpyAsync(hc, c->d_counter.p, sizeof(hc), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  if(hc[2]) return c->fail(std::string(who) + ": the search reached a cut cell of a multi-rank tree");
  if(n_swallowed)
    {
      n_swallowed[0] = (long long) hc[4];
      n_swallowed[1] = (long long) hc[5];
      n_swallowed[2] = (long long) hc[6];
    }
  return 0;
}

// compact device block of the AoS entry points: ncols columns of m doubles, staged through one host vector
struct Block
{
  std::vector<double> h;
  double *d = nullptr;
  long long m = 0;
  int ncols = 0;
  int alloc(gl_ctx *c, long long rows, int cols)
  {
    m = rows;
    ncols = cols;
    h.assign((size_t) rows * cols, 0.0);
    GL_CUDA(c, cudaMalloc((void **) &d, (size_t) (rows * cols + 1) * sizeof(double)));
    return 0;
  }
  double *hcol(int k) { return h.data() + (size_t) k * m; }
  double *dcol(int k) { return d + (size_t) k * m; }
  int up(gl_ctx *c)
  {
    GL_CUDA(c, cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return 0;
  }
  int down(gl_ctx *c)
  {
    GL_CUDA(c, cudaMemcpyAsync(h.data(), d, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    GL_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
  }
  ~Block()
  {
    if(d) cudaFree(d);
  }
};
int sg_gadget_ready(gl_ctx *c, const char *who, const gl_gadget_subgrid_layout *L2, const void *P)
{
  if(!c->have_lay) return c->fail(std::string(who) + ": gl_gadget_bind first");
  if(!L2 || !P) return c->fail(std::string(who) + ": layout and P[] are required");
  if(c->g_numpart != c->n || c->g_off != 0) return c->fail(std::string(who) + ": the context does not hold this P[] (gl_gadget_pull / gl_gadget_compute_accelerations first)");
  return 0;
}
// the active particles of one type from the reference's own arrays, in active-list order -> c->sg_list
int sg_gadget_targets(gl_ctx *c, const void *P, const int *active, long long nactive, int want, long long *count)
{
  const gl_gadget_layout &L = c->lay;
  const long long n = c->n, m = active ? nactive : n;
  std::vector<int> t;
  for(long long k = 0; k < m; k++)
    {
      const long long i = active ? active[k] : k;
      if(i < 0 || i >= n) return c->fail("sub-grid loop: active list holds an index outside [0, NumPart)");
      if(*(const short *) ((const char *) P + i * L.p_stride + L.p_Type) == want) t.push_back((int) i);
    }
  *count = (long long) t.size();
  if(t.empty()) return 0;
  GL_CUDA(c, c->sg_list.ensure(t.size() + 1));
  GL_CUDA(c, cudaMemcpyAsync(c->sg_list.p, t.data(), t.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}
}	// namespace

extern "C" int gl_set_subgrid_params(gl_ctx *c, const gl_subgrid_params *p)
{
  if(!p) return c->fail("gl_set_subgrid_params: null");
  if(!(p->UnitTime_in_s > 0) || !(p->UnitMass_in_g > 0) || !(p->UnitLength_in_cm > 0) || !(p->UnitEnergy_in_cgs > 0))
    return c->fail("gl_set_subgrid_params: the unit system must be set (UnitTime_in_s, UnitMass_in_g, UnitLength_in_cm, UnitEnergy_in_cgs > 0)");
  c->subgrid = *p;
  c->have_subgrid = true;
  return 0;
}

extern "C" int gl_blackhole_accretion(gl_ctx *c, const int *active, long long nactive, const gl_sink_state *st, int where,
				      long long n_swallowed[3])
{
  if(sg_ready(c, "gl_blackhole_accretion")) return 1;
  if(!st || !st->vel || !st->id || !st->swallow_id || !st->bh_mass || !st->bh_mdot || !st->accdisc_mass || !st->dust_mass || !st->total_mass ||
     !st->injected_bh_energy)
    return c->fail("gl_blackhole_accretion: vel, id, swallow_id, bh_mass, bh_mdot, accdisc_mass, dust_mass, total_mass, injected_bh_energy are required");
  if(!st->sink_density && !c->o_sinkrho.p) return c->fail("gl_blackhole_accretion: no sink_density given and gl_density has not run");
  const size_t n = (size_t) c->n;
  long long ns = 0;
  if(sg_targets(c, active, nactive, 5, &ns)) return 1;
  if(n_swallowed) n_swallowed[0] = n_swallowed[1] = n_swallowed[2] = 0;
  if(ns == 0)
    {
      // no active sink: nothing is read or written (the caller's arrays are not even copied to the device)
      if(st->mass)
	GL_CUDA(c, cudaMemcpyAsync(st->mass, c->in_mass.p, n * sizeof(double), where ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
      GL_CUDA(c, cudaStreamSynchronize(c->stream));
      return 0;
    }
  Staged<double> vel, bhm, bhd, acd, dum, tom, rho, inj;
  Staged<int> id, swl;
  if(vel.bind(c, st->vel, 3 * n, where, true, true) || id.bind(c, st->id, n, where, true, false) || swl.bind(c, st->swallow_id, n, where, true, true) ||
     bhm.bind(c, st->bh_mass, n, where, true, true) || bhd.bind(c, st->bh_mdot, n, where, true, true) || acd.bind(c, st->accdisc_mass, n, where, true, true) ||
     dum.bind(c, st->dust_mass, n, where, true, true) || tom.bind(c, st->total_mass, n, where, true, true) ||
     rho.bind(c, st->sink_density, n, where, true, false) || inj.bind(c, st->injected_bh_energy, n, where, true, true))
    return 1;
  SinkArr A{ vel.dev, id.dev, swl.dev, bhm.dev, bhd.dev, acd.dev, dum.dev, tom.dev, st->sink_density ? rho.dev : c->o_sinkrho.p, inj.dev, nullptr };
  if(sg_run_sinks(c, ns, A, nullptr, (long long) n)) return 1;
  if(vel.finish(c) || swl.finish(c) || bhm.finish(c) || bhd.finish(c) || acd.finish(c) || dum.finish(c) || tom.finish(c) || inj.finish(c)) return 1;
  if(st->mass)
    GL_CUDA(c, cudaMemcpyAsync(st->mass, c->in_mass.p, n * sizeof(double), where ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  return sg_counts(c, "gl_blackhole_accretion", n_swallowed);
}

extern "C" int gl_dust_density(gl_ctx *c, const int *active, long long nactive, double *dust_particle_density, int where)
{
  if(sg_ready(c, "gl_dust_density")) return 1;
  if(!dust_particle_density) return c->fail("gl_dust_density: output array is required");
  const size_t n = (size_t) c->n;
  long long nd = 0;
  if(sg_targets(c, active, nactive, 2, &nd)) return 1;
  if(nd == 0) return 0;
  Staged<double> out;
  if(out.bind(c, dust_particle_density, n, where, true, true)) return 1;
  k_dust_density<<<gl_grid(nd * 32, GL_BLOCK), GL_BLOCK, 0, c->stream>>>((int) nd, c->sg_list.p, sg_tree(c), c->in_pos.p, c->in_hsml.p, c->in_mass.p,
									  out.dev, false, c->d_counter.p);
  GL_LAUNCH_COUNT(c);
  GL_CUDA(c, cudaGetLastError());
  if(out.finish(c)) return 1;
  return sg_counts(c, "gl_dust_density", nullptr);
}

extern "C" int gl_dust_drag(gl_ctx *c, const int *active, long long nactive, const gl_dust_state *st, int where)
{
  if(sg_ready(c, "gl_dust_drag")) return 1;
  if(!st || !st->vel || !st->dust_particle_density || !st->dust_radius || !st->dust_vcoll || !st->delta_dust_momentum || !st->new_drag_acc ||
     !st->delta_drag_energy || !st->drag_heating)
    return c->fail("gl_dust_drag: vel, dust_particle_density, dust_radius, dust_vcoll, delta_dust_momentum, new_drag_acc, delta_drag_energy, "
		   "drag_heating are required");
  if(!st->gravaccel && !c->o_acc.p) return c->fail("gl_dust_drag: no gravaccel given and gl_gravity_tree has not run");
  if((!st->gas_density || !st->gas_entropy || !st->gas_vel) && !c->o_sinkrho.p)
    return c->fail("gl_dust_drag: no gas_density / gas_entropy / gas_vel given and gl_density has not run");
  if(!st->entropy && !c->in_entropy.p) return c->fail("gl_dust_drag: no entropy given and none loaded");
  const size_t n = (size_t) c->n;
  long long nd = 0;
  if(sg_targets(c, active, nactive, 2, &nd)) return 1;
  if(nd == 0) return 0;
  Staged<double> vel, ga, rho, entr, gv, d7, d9, rad, vc, ddm, nda, dde, S, heat;
  if(vel.bind(c, st->vel, 3 * n, where, true, true) || ga.bind(c, st->gravaccel, 3 * n, where, true, false) || rho.bind(c, st->gas_density, n, where, true, false) ||
     entr.bind(c, st->gas_entropy, n, where, true, false) || gv.bind(c, st->gas_vel, 3 * n, where, true, false) ||
     d7.bind(c, st->dust_particle_density, n, where, true, false) || d9.bind(c, st->dust_particle_velocity, 3 * n, where, true, true) ||
     rad.bind(c, st->dust_radius, n, where, true, false) || vc.bind(c, st->dust_vcoll, n, where, true, true) ||
     ddm.bind(c, st->delta_dust_momentum, 3 * n, where, true, true) || nda.bind(c, st->new_drag_acc, 3 * n, where, true, true) ||
     dde.bind(c, st->delta_drag_energy, n, where, true, true) || S.bind(c, st->entropy, n, where, true, true) ||
     heat.bind(c, st->drag_heating, n, where, true, true))
    return 1;
  DustArr A{ vel.dev, st->gravaccel ? ga.dev : c->o_acc.p, st->gas_density ? rho.dev : c->o_sinkrho.p, st->gas_entropy ? entr.dev : c->o_sinkentr.p,
	     st->gas_vel ? gv.dev : c->o_sinkvel.p, d7.dev, d9.dev, rad.dev, vc.dev, ddm.dev, nda.dev, dde.dev, st->entropy ? S.dev : c->in_entropy.p, heat.dev,
	     nullptr };
  if(sg_run_dust(c, nd, A)) return 1;
  if(vel.finish(c) || d9.finish(c) || vc.finish(c) || ddm.finish(c) || nda.finish(c) || dde.finish(c) || S.finish(c) || heat.finish(c)) return 1;
  return sg_counts(c, "gl_dust_drag", nullptr);
}

// ---- the same three on the reference's own P[] / SphP[] (drop-in layer) -------------------------------------------
// Only the particles the loops touch cross PCIe: the search runs first and returns their indices, the host gathers their
// fields from the structs, the kernels run on those rows, the host scatters what changed.
#define PFIELD(i, off) ((char *) P + (i) * L.p_stride + (off))
#define SFIELD(i, off) ((char *) SphP + (i) * L.sph_stride + (off))

extern "C" int gl_gadget_blackhole_accretion(gl_ctx *c, const gl_gadget_subgrid_layout *L2, void *P, void *SphP, const int *active, long long nactive,
					     long long n_swallowed[3])
{
  if(sg_ready(c, "gl_gadget_blackhole_accretion") || sg_gadget_ready(c, "gl_gadget_blackhole_accretion", L2, P)) return 1;
  const gl_gadget_layout &L = c->lay;
  const long long ngas = c->g_ngas;
  if(n_swallowed) n_swallowed[0] = n_swallowed[1] = n_swallowed[2] = 0;
  long long ns = 0, m = 0;
  if(sg_gadget_targets(c, P, active, nactive, 5, &ns)) return 1;
  if(ns == 0) return 0;
  std::vector<int> rows;
  if(sg_compact(c, ns, (1u << 0) | (1u << 2) | (1u << 5), &m, rows)) return 1;
  enum { VX, VY, VZ, BHM, BHD, ACD, DUM, TOM, RHO, INJ, MASS, NCOL };
  Block B;
  if(B.alloc(c, m, NCOL)) return 1;
  std::vector<int> ids((size_t) 2 * m);
  // the kernels address vel as 3 doubles per row: columns VX..VZ are laid out row-major in one 3*m block
  double *hv = B.hcol(VX);
  for(long long t = 0; t < m; t++)
    {
      const long long i = rows[t];
      const double *v = (const double *) PFIELD(i, L.p_Vel);
      hv[3 * t] = v[0]; hv[3 * t + 1] = v[1]; hv[3 * t + 2] = v[2];
      ids[t] = *(const int *) PFIELD(i, L2->p_ID);
      ids[m + t] = *(const int *) PFIELD(i, L.p_SwallowID);
      const int type = *(const short *) PFIELD(i, L.p_Type);
      if(type == 5)
	{
	  B.hcol(BHM)[t] = *(const double *) PFIELD(i, L2->p_BH_Mass);
	  B.hcol(BHD)[t] = *(const double *) PFIELD(i, L2->p_BH_Mdot);
	  B.hcol(ACD)[t] = *(const double *) PFIELD(i, L2->p_AccDisc_Mass);
	  B.hcol(DUM)[t] = *(const double *) PFIELD(i, L2->p_Dust_Mass);
	  B.hcol(TOM)[t] = *(const double *) PFIELD(i, L2->p_Total_Mass);
	  B.hcol(RHO)[t] = *(const double *) PFIELD(i, L.p_SinkDensity);
	}
      // injected energy: the device accumulates from zero, the host adds the sum to SphP (same sum, one rounding more)
    }
  int *dids = nullptr;
  GL_CUDA(c, cudaMalloc((void **) &dids, (size_t) (2 * m + 1) * sizeof(int)));
  struct Free { int *p; ~Free() { cudaFree(p); } } fr{ dids };
  GL_CUDA(c, cudaMemcpyAsync(dids, ids.data(), (size_t) 2 * m * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  if(B.up(c)) return 1;
  SinkArr A{ B.dcol(VX), dids, dids + m, B.dcol(BHM), B.dcol(BHD), B.dcol(ACD), B.dcol(DUM), B.dcol(TOM), B.dcol(RHO), B.dcol(INJ), c->sg_slot.p };
  if(sg_run_sinks(c, ns, A, c->sg_plist.p, m)) return 1;
  k_sg_gather<<<gl_grid(m, 256), 256, 0, c->stream>>>(m, c->sg_plist.p, c->in_mass.p, B.dcol(MASS), 1);
  GL_LAUNCH_COUNT(c);
  GL_CUDA(c, cudaMemcpyAsync(ids.data(), dids, (size_t) 2 * m * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if(B.down(c)) return 1;
  for(long long t = 0; t < m; t++)
    {
      const long long i = rows[t];
      const int type = *(const short *) PFIELD(i, L.p_Type);
      *(int *) PFIELD(i, L.p_SwallowID) = ids[m + t];
      *(double *) PFIELD(i, L.p_Mass) = B.hcol(MASS)[t];
      if(type == 5)
	{
	  double *v = (double *) PFIELD(i, L.p_Vel);
	  v[0] = hv[3 * t]; v[1] = hv[3 * t + 1]; v[2] = hv[3 * t + 2];
	  *(double *) PFIELD(i, L2->p_BH_Mass) = B.hcol(BHM)[t];
	  *(double *) PFIELD(i, L2->p_BH_Mdot) = B.hcol(BHD)[t];
	  *(double *) PFIELD(i, L2->p_AccDisc_Mass) = B.hcol(ACD)[t];
	  *(double *) PFIELD(i, L2->p_Dust_Mass) = B.hcol(DUM)[t];
	  *(double *) PFIELD(i, L2->p_Total_Mass) = B.hcol(TOM)[t];
	}
      else if(type == 0 && i < ngas && SphP && B.hcol(INJ)[t] != 0)
	*(double *) SFIELD(i, L2->s_Injected_BH_Energy) += B.hcol(INJ)[t];
    }
  return sg_counts(c, "gl_gadget_blackhole_accretion", n_swallowed);
}

extern "C" int gl_gadget_dust_density(gl_ctx *c, const gl_gadget_subgrid_layout *L2, void *P, const int *active, long long nactive)
{
  if(sg_ready(c, "gl_gadget_dust_density") || sg_gadget_ready(c, "gl_gadget_dust_density", L2, P)) return 1;
  const gl_gadget_layout &L = c->lay;
  long long nd = 0;
  if(sg_gadget_targets(c, P, active, nactive, 2, &nd)) return 1;
  if(nd == 0) return 0;
  Block B;
  if(B.alloc(c, nd, 1)) return 1;
  k_dust_density<<<gl_grid(nd * 32, GL_BLOCK), GL_BLOCK, 0, c->stream>>>((int) nd, c->sg_list.p, sg_tree(c), c->in_pos.p, c->in_hsml.p, c->in_mass.p,
									  B.d, true, c->d_counter.p);
  GL_LAUNCH_COUNT(c);
  GL_CUDA(c, cudaGetLastError());
  std::vector<int> t((size_t) nd);
  GL_CUDA(c, cudaMemcpyAsync(t.data(), c->sg_list.p, (size_t) nd * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if(B.down(c)) return 1;
  for(long long d = 0; d < nd; d++) *(double *) PFIELD((long long) t[d], L2->p_DustParticleDensity) = B.h[d];
  return sg_counts(c, "gl_gadget_dust_density", nullptr);
}

extern "C" int gl_gadget_dust_drag(gl_ctx *c, const gl_gadget_subgrid_layout *L2, void *P, void *SphP, const int *active, long long nactive)
{
  if(sg_ready(c, "gl_gadget_dust_drag") || sg_gadget_ready(c, "gl_gadget_dust_drag", L2, P)) return 1;
  if(!SphP) return c->fail("gl_gadget_dust_drag: SphP[] is required");
  const gl_gadget_layout &L = c->lay;
  const long long ngas = c->g_ngas;
  long long nd = 0, m = 0;
  if(sg_gadget_targets(c, P, active, nactive, 2, &nd)) return 1;
  if(nd == 0) return 0;
  std::vector<int> rows;
  if(sg_compact(c, nd, 1u << 0, &m, rows)) return 1;
  // 3-vectors first (row-major blocks of 3*m), then the scalars
  enum { VEL = 0, GACC = 3, GVEL = 6, D9 = 9, DDM = 12, NDA = 15, RHO = 18, ENTR, D7, RAD, VC, DDE, S, HEAT, NCOL };
  Block B;
  if(B.alloc(c, m, NCOL)) return 1;
  for(long long t = 0; t < m; t++)
    {
      const long long i = rows[t];
      const int type = *(const short *) PFIELD(i, L.p_Type);
      const double *v = (const double *) PFIELD(i, L.p_Vel);
      for(int k = 0; k < 3; k++) B.hcol(VEL)[3 * t + k] = v[k];
      if(type == 2)
	{
	  const double *g = (const double *) PFIELD(i, L.p_GravAccel), *gv = (const double *) PFIELD(i, L.p_DustGasVel),
	    *w = (const double *) PFIELD(i, L2->p_DustParticleVelocity);
	  for(int k = 0; k < 3; k++)
	    {
	      B.hcol(GACC)[3 * t + k] = g[k];
	      B.hcol(GVEL)[3 * t + k] = gv[k];
	      B.hcol(D9)[3 * t + k] = w[k];
	    }
	  B.hcol(RHO)[t] = *(const double *) PFIELD(i, L.p_DustDensity);
	  B.hcol(ENTR)[t] = *(const double *) PFIELD(i, L.p_DustEntropy);
	  B.hcol(D7)[t] = *(const double *) PFIELD(i, L2->p_DustParticleDensity);
	  B.hcol(RAD)[t] = *(const double *) PFIELD(i, L2->p_DustRadius);
	  B.hcol(VC)[t] = *(const double *) PFIELD(i, L2->p_DustVcoll);
	}
      else if(type == 0 && i < ngas)
	{
	  B.hcol(S)[t] = *(const double *) SFIELD(i, L.s_Entropy);
	  B.hcol(HEAT)[t] = *(const double *) SFIELD(i, L2->s_DragHeating);
	}
    }
  if(B.up(c)) return 1;
  DustArr A{ B.dcol(VEL), B.dcol(GACC), B.dcol(RHO), B.dcol(ENTR), B.dcol(GVEL), B.dcol(D7), B.dcol(D9), B.dcol(RAD), B.dcol(VC), B.dcol(DDM), B.dcol(NDA),
	     B.dcol(DDE), B.dcol(S), B.dcol(HEAT), c->sg_slot.p };
  if(sg_run_dust(c, nd, A)) return 1;
  if(B.down(c)) return 1;
  for(long long t = 0; t < m; t++)
    {
      const long long i = rows[t];
      const int type = *(const short *) PFIELD(i, L.p_Type);
      double *v = (double *) PFIELD(i, L.p_Vel);
      for(int k = 0; k < 3; k++) v[k] = B.hcol(VEL)[3 * t + k];
      if(type == 2)
	{
	  double *w = (double *) PFIELD(i, L2->p_DustParticleVelocity), *dm = (double *) PFIELD(i, L2->p_DeltaDustMomentum),
	    *na = (double *) PFIELD(i, L2->p_NewDragAcc);
	  for(int k = 0; k < 3; k++)
	    {
	      w[k] = B.hcol(D9)[3 * t + k];
	      dm[k] = B.hcol(DDM)[3 * t + k];
	      na[k] = B.hcol(NDA)[3 * t + k];
	    }
	  *(double *) PFIELD(i, L2->p_DustVcoll) = B.hcol(VC)[t];
	  *(double *) PFIELD(i, L2->p_DeltaDragEnergy) = B.hcol(DDE)[t];
	}
      else if(type == 0 && i < ngas)
	{
	  *(double *) SFIELD(i, L.s_Entropy) = B.hcol(S)[t];
	  *(double *) SFIELD(i, L2->s_DragHeating) = B.hcol(HEAT)[t];
	}
    }
  return sg_counts(c, "gl_gadget_dust_drag", nullptr);
}
