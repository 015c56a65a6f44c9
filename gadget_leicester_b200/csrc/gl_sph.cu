// gl_sph.cu - K7/K8 SPH density with in-kernel smoothing-length iteration, K10 hydro force.
//
// Reference: density() (density.c:89-704) loops "evaluate all unconverged targets, post-process,
// adjust Hsml" until every target holds DesNumNgb +- MaxNumNgbDeviation kernel-weighted neighbours;
// density_evaluate (density.c:711-1029) gathers the neighbours with ngb_treefind_variable
// (ngb.c:169-297) and forms 11 sums; hydro_force/hydro_evaluate (hydra.c:145-813, :822-1995) do the
// same with the symmetric search ngb_treefind_pairs (ngb.c:32-160).
//
// B200 design.  One warp owns 32 consecutive SPH targets of the Peano-Hilbert order (a compact
// clump).  The Hsml sequence of a target depends on that target alone (the one-sided search only
// uses h_i), so the whole iteration runs inside ONE kernel launch, round by round, until every
// lane of the warp has converged - no host round trips, no relaunch, no MPI_Allreduce of
// "particles left" (density.c:656).  Each round:
//   1. collect  - the warp keeps a stack of tree cells; every step the lanes pop up to 32 of them
//                 with independent loads, park the (FP32, slack-padded) cubes in shared memory, and
//                 every lane tests ITS OWN search sphere against each cube (one far-out lane does
//                 not inflate everybody's search); a touched cell with <= 32 particles is swallowed
//                 with one coalesced load, a larger one pushes its children.
//   2. stage    - the gas candidates of the swallowed cells go to SHARED MEMORY as FP32 offsets from
//                 the group centre (192 per chunk), with the candidate's own threshold for the
//                 symmetric search.
//   3. filter   - lane = target: every lane scans the chunk (broadcast reads, 3 FMAs per test,
//                 conservative margins) and appends what may be a neighbour of ITS target to a
//                 private ring in shared memory.
//   4. heavy    - the reference's FP64 predicate (r2 < h^2, same operation order, no FMA) and the
//                 kernel sums are taken from the rings whenever one of them is nearly full; every
//                 lane adds into its own accumulators.  With few targets left in a warp a
//                 target-major variant (ballot compaction, WARP-SHUFFLE sums) takes over.
// The neighbour SETS (hence the integer pair counts) are exactly the reference's; the FP64 sums are
// formed in tree order per lane (or as 32 partial sums + butterfly in the target-major variant),
// i.e. they agree with the reference to ~1e-15 relative.  Targets with thousands of neighbours (a
// sink in the disc hole) go to a thread-block-cluster path (k_fat_scan / k_fat_rounds).
#include "gl_internal.h"
#include <climits>
#include <cooperative_groups.h>
#include <algorithm>

gl_dparams gl_make_dparams(const gl_ctx *c);

#define KERNEL_COEFF_1 2.546479089470	// allvars.h:247-253
#define KERNEL_COEFF_2 15.278874536822
#define KERNEL_COEFF_3 45.836623610466
#define KERNEL_COEFF_4 30.557749073644
#define KERNEL_COEFF_5 5.092958178941
#define KERNEL_COEFF_6 (-15.278874536822)
#define NORM_COEFF 4.188790204786
#define GL_GAMMA (7. / 5.)	// allvars.h:64
#define GL_GAMMA_MINUS1 (GL_GAMMA - 1)
#define BOX_SLACK 1.0000000001
#ifndef CAND_CAP
#define CAND_CAP 192		// staged candidates per warp
#endif
#ifndef QCAP
#define QCAP 32			// lane = target mode: entries of every lane's private hit queue (power of two)
#endif
#define SPH_WARPS 2		// warps per block: 64-thread blocks (3.48 / 3.70 ms density / hydro; 128 threads 3.57 / 3.75, 32 threads 3.52 / 3.74)
#ifndef STK_CAP
#define STK_CAP 512
#endif
#define LT_MIN_ACTIVE 12	// with at least this many active targets a warp runs lane = target (see density_chunk_lt)
#define FAT_THRESH 4096	// neighbours above which a target leaves the warp path
#define FAT_MAX 64
#define STK_RESERVE 160	// 7 pushes per level x 21 levels: pure depth-first continuation never overflows
#define SPH_BLOCK (32 * SPH_WARPS)
#ifndef SPH_MIN_BLOCKS
#define SPH_MIN_BLOCKS 6
#endif
#define FULL 0xffffffffu

__device__ __forceinline__ double warp_sum_d(double v)
{
  for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_max_d(double v)
{
  for(int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ double warp_min_d(double v)
{
  for(int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// squared distance from point p to the cube (centre B.xyz, half side `half`)
__device__ __forceinline__ double dist2_to_cube(double px, double py, double pz, const double4 &B, double half)
{
  const double dx = fmax(0.0, fabs(B.x - px) - half);
  const double dy = fmax(0.0, fabs(B.y - py) - half);
  const double dz = fmax(0.0, fabs(B.z - pz) - half);
  return dx * dx + dy * dy + dz * dz;
}

// Per-warp shared-memory stage.  Candidates are kept as FP32 coordinates RELATIVE TO THE GROUP CENTRE: they only feed
// a conservative pre-filter (see cand_threshold); every accept/reject decision that counts is re-taken in FP64 from
// the FP64 particle records.
#ifdef GL_PHASE_PROF
__device__ unsigned long long g_phase[8];
#define PH_T(v) const long long v = clock64()
#define PH_ADD(S, idx, v) do { if((threadIdx.x & 31) == 0) (S).ph[idx] += (unsigned long long) (clock64() - (v)); } while(0)
#else
#define PH_T(v)
#define PH_ADD(S, idx, v)
#endif

struct WarpStage
{
#ifdef GL_PHASE_PROF
  unsigned long long ph[8];
#endif
  float4 cand[CAND_CAP];	// candidate position - group centre (x,y,z); w = the candidate's own FP32 threshold (symmetric search)
  int j[CAND_CAP];		// tree index of the candidate
  int q[QCAP * 32];		// lane = target mode: q[slot * 32 + lane] = ring of tree indices whose FP32 pre-test passed for that lane
  int hit[64];			// target-major mode: compacted "maybe" list of the current target
  int stk[STK_CAP];		// tree cells still to be examined (DFS entry index, or -(p+2) for a lone particle)
  int2 rng[32];			// particle ranges swallowed in the current step
  float4 cell[32];		// cubes popped in the current step: centre - group centre, padded half side (< 0: not a cube)
  float cellh[32];		// their padded hmax (symmetric search only)
};

// FP32 view of the warp's targets for the two conservative filters
struct GroupF
{
  double gcx, gcy, gcz;		// group centre (FP64)
  float ax, ay, az;		// this lane's target position - group centre (idle lanes: 1e18, out of everybody's reach)
  float r;			// this lane's search radius (Hsml), 0 when the lane is idle
  float rpad, r2pad;		// radius padded for the FP32 cube tests of the collect phase, and its square
  float amax;			// max |a| component over the active lanes of the warp
};

// Conservative FP32 bound for "r2 < h*h".  With a, b the FP64 offsets of target and candidate from the group centre
// and L >= every |component| involved, each component of (af - bf) is off by at most 3*2^-24*L, so
// |r2f - r2| <= 2*sqrt(3)*|d|*1.8e-7*L + 3*(1.8e-7*L)^2 + 2e-7*r2.  Near the boundary |d| ~ h; the threshold below
// exceeds that bound more than tenfold, so a true neighbour can never be filtered out (the pair counts stay exact).
__device__ __forceinline__ float cand_threshold(float h, float L)
{
  return h * h * 1.00002f + (8.0e-6f * h + 1.0e-11f * L) * L;
}

// append the gas particles of the tree range [a,b) (b-a <= 32) to the staged list; returns the new count.
// The three loads are issued together (one memory latency per range, not three).  With WANT_H the candidate's own
// FP32 threshold for "r2 < h_j^2" is formed here, once per candidate, with L = max(|its offset|, amax) - the bound
// of cand_threshold only needs L >= the components of THIS pair.
template <bool WANT_H, class Stage>
__device__ __forceinline__ int stage_range(Stage &S, int ncand, int a, int b, int lane, const GroupF &G, float &lmax,
					    const double4 *__restrict__ s_pm, const double4 *__restrict__ s_vh,
					    const int *__restrict__ s_type)
{
  const int j = a + lane;
  const bool in = j < b;
  int ty = -1;
  double hj = 0;
  double4 pm = make_double4(0, 0, 0, 0);
  if(in)
    {
      ty = __ldg(&s_type[j]);
      pm = gl_ld4(&s_pm[j]);
      if(WANT_H) hj = __ldg(&s_vh[j].w);
    }
  // ngb.c:54 / :191: only gas particles are neighbours; density.c:832 / hydra.c:1236: skip Mass == 0
  const bool ok = in && ty == 0 && pm.w != 0;
  const unsigned ball = __ballot_sync(FULL, ok);
  if(ok)
    {
      const int pos = ncand + __popc(ball & ((1u << lane) - 1));
      const float bx = (float) (pm.x - G.gcx), by = (float) (pm.y - G.gcy), bz = (float) (pm.z - G.gcz);
      const float lj = fmaxf(fabsf(bx), fmaxf(fabsf(by), fabsf(bz)));
      S.cand[pos] = make_float4(bx, by, bz, WANT_H ? cand_threshold((float) hj, fmaxf(lj, G.amax)) : 0.f);
      S.j[pos] = j;
      lmax = fmaxf(lmax, lj);
    }
  return ncand + __popc(ball);
}

// after a batch of `take` hits has been consumed, move the remaining hit-list entries to the front
__device__ __forceinline__ void hits_shift(WarpStage &S, int nh, int lane)
{
  __syncwarp();
  if(nh > 32)
    {
      const int mv = S.hit[32 + lane];
      __syncwarp();
      if(lane < nh - 32) S.hit[lane] = mv;
      __syncwarp();
    }
}

// Phase 1, warp-parallel.  The warp keeps a stack of tree cells still to be examined.  Every step the lanes
// pop up to 32 cells and fetch their cubes with INDEPENDENT loads (one memory latency per step instead of one
// per cell); the cubes are parked in shared memory and every lane tests ITS OWN search sphere against each of them
// (ngb.c:270-289 / :136-152, here in FP32 with a conservative margin - this test only prunes).  A touched cell
// is swallowed (<= 32 particles: its tree range is staged with one coalesced load) or its children are pushed
// (child table written by k_moments).  flush(ncand, lmax) runs phases 2+3.
template <bool USE_HMAX, class Stage, class Flush>
__device__ __forceinline__ void sph_collect(Stage &S, int lane, int nentries, const int *__restrict__ e_next,
					    const double4 *__restrict__ e_B, const double *__restrict__ e_hmax,
					    const int2 *__restrict__ e_rng, const int *__restrict__ n_child,
					    const double4 *__restrict__ s_pm, const double4 *__restrict__ s_vh,
					    const int *__restrict__ s_type, unsigned active, const GroupF &G,
					    unsigned &visits, unsigned &staged, unsigned long long *__restrict__ counter, Flush flush)
{
  if(nentries <= 0 || active == 0) return;
  int ncand = 0;
  float lmax = 0.f;
  if(lane == 0) S.stk[0] = (__ldg(&e_next[0]) >= 0) ? 0 : -2;
  __syncwarp();
  int sp = 1;
  const unsigned lt = (1u << lane) - 1;
  while(sp > 0)
    {
      int k = sp < 32 ? sp : 32;
      int room = (STK_CAP - STK_RESERVE - sp) / 7;
      if(room < 1) room = 1;
      if(k > room) k = room;
      PH_T(tc0);
      const int my = (lane < k) ? S.stk[sp - 1 - lane] : -1;
      sp -= k;
      visits += k;
      const bool isnode = my >= 0;
      const bool ispart = my < -1;
      int2 rg = make_int2(0, 0);
      float4 cf = make_float4(0.f, 0.f, 0.f, -1.0e30f);	// "half side" -1e30: the distance test below can never pass
      float hm = 0.f;
      if(isnode)
	{
	  const double4 B = gl_ld4(&e_B[my]);
	  rg = __ldg(&e_rng[my]);
	  if(USE_HMAX) hm = (float) __ldg(&e_hmax[my]) * 1.001002f + 1.0e-6f * G.amax;
	  const float cx = (float) (B.x - G.gcx), cy = (float) (B.y - G.gcy), cz = (float) (B.z - G.gcz), hf = (float) (0.5 * B.w) * 1.000001f;
	  // FP32 slack for the coordinates that enter the subtractions, folded into the cube: growing the cube by s
	  // shortens the distance to it by at least min(distance, s), so this is at least as conservative as adding s
	  // to the search radius
	  cf = make_float4(cx, cy, cz, hf + 1.0e-6f * (fabsf(cx) + fabsf(cy) + fabsf(cz) + hf));
	}
      S.cell[lane] = cf;
      if(USE_HMAX) S.cellh[lane] = hm;
      __syncwarp();
      // lane = target: which of the k cubes does my sphere touch?  (idle lanes sit at 1e18: they touch nothing)
      unsigned mybits = 0;
      for(int i = 0; i < k; i++)
	{
	  const float4 c = S.cell[i];
	  const float dx = fmaxf(0.f, fabsf(c.x - G.ax) - c.w), dy = fmaxf(0.f, fabsf(c.y - G.ay) - c.w),
	    dz = fmaxf(0.f, fabsf(c.z - G.az) - c.w);
	  const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
	  float R2 = G.r2pad;
	  if(USE_HMAX)
	    {
	      const float R = fmaxf(G.rpad, S.cellh[i]);
	      R2 = R * R;
	    }
	  mybits |= (d2 < R2) ? (1u << i) : 0u;
	}
      const unsigned touched = __reduce_or_sync(FULL, mybits);
      bool acc = isnode && ((touched >> lane) & 1);
      if(acc && rg.x < 0)
	{
	  // a CUT CELL of the locally-essential tree (multi-rank mode) within reach of a search sphere: the gas below it
	  // stayed with its owner, whose export test promised this cannot happen.  Report instead of missing neighbours.
	  atomicMax(&counter[2], GL_ERR_LET_VIOLATION);
	  acc = false;
	}
      const bool small = acc && (rg.y - rg.x <= 32);
      bool big = acc && !small;
      bool bucket = false;
      // push the children of the large cells
      int4 c0 = make_int4(-1, -1, -1, -1), c1 = c0;
      if(big)
	{
	  const int4 *ct = reinterpret_cast<const int4 *>(n_child + 8ll * (my - rg.x));
	  c0 = __ldg(ct);
	  c1 = __ldg(ct + 1);
	  if(c0.x == INT_MIN)
	    {
	      bucket = true;	// finest-level cell holding coincident keys: no child table, staged piecewise below
	      big = false;
	    }
	}
      const int cc[8] = { c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w };
#pragma unroll
      for(int i = 0; i < 8; i++)
	{
	  const bool has = big && cc[i] != -1;
	  const unsigned ball = __ballot_sync(FULL, has);
	  if(has) S.stk[sp + __popc(ball & lt)] = cc[i];
	  sp += __popc(ball);
	}
      PH_ADD(S, 0, tc0);
      // swallow the small cells and the lone particles
      const bool isrange = small || ispart;
      if(ispart) rg = make_int2(-(my + 2), -(my + 2) + 1);
      const unsigned rball = __ballot_sync(FULL, isrange);
      if(isrange) S.rng[__popc(rball & lt)] = rg;
      const int nr = __popc(rball);
      __syncwarp();
      for(int i = 0; i < nr; i++)
	{
	  const int2 r = S.rng[i];
	  if(ncand + 32 > CAND_CAP)
	    {
	      flush(ncand, lmax);
	      ncand = 0;
	      lmax = 0.f;
	    }
	  PH_T(ts0);
	  ncand = stage_range<USE_HMAX>(S, ncand, r.x, r.y, lane, G, lmax, s_pm, s_vh, s_type);
	  PH_ADD(S, 1, ts0);
	  staged += r.y - r.x;
	}
      for(unsigned bm = __ballot_sync(FULL, bucket); bm; bm &= bm - 1)
	{
	  const int b = __ffs(bm) - 1;
	  const int ra = __shfl_sync(FULL, rg.x, b), rb = __shfl_sync(FULL, rg.y, b);
	  for(int a = ra; a < rb; a += 32)
	    {
	      if(ncand + 32 > CAND_CAP)
		{
		  flush(ncand, lmax);
		  ncand = 0;
		  lmax = 0.f;
		}
	      ncand = stage_range<USE_HMAX>(S, ncand, a, (a + 32 < rb) ? a + 32 : rb, lane, G, lmax, s_pm, s_vh, s_type);
	    }
	  staged += rb - ra;
	}
      __syncwarp();
    }
  if(ncand > 0) flush(ncand, lmax);
}

// group centre + FP32 target offsets for one round
__device__ __forceinline__ GroupF make_group(bool act, double px, double py, double pz, double radius)
{
  GroupF G;
  const double xlo = warp_min_d(act ? px : 1e300), xhi = warp_max_d(act ? px : -1e300);
  const double ylo = warp_min_d(act ? py : 1e300), yhi = warp_max_d(act ? py : -1e300);
  const double zlo = warp_min_d(act ? pz : 1e300), zhi = warp_max_d(act ? pz : -1e300);
  G.gcx = 0.5 * (xlo + xhi);
  G.gcy = 0.5 * (ylo + yhi);
  G.gcz = 0.5 * (zlo + zhi);
  G.ax = act ? (float) (px - G.gcx) : 0.f;
  G.ay = act ? (float) (py - G.gcy) : 0.f;
  G.az = act ? (float) (pz - G.gcz) : 0.f;
  G.r = act ? (float) radius * 1.000001f : 0.f;
  float am = fmaxf(fabsf(G.ax), fmaxf(fabsf(G.ay), fabsf(G.az)));
  for(int o = 16; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(FULL, am, o));
  G.amax = am;
  // relative slack on the radius, absolute slack for the target coordinates that entered the subtractions
  G.rpad = G.r * 1.001f + 1.0e-6f * am;
  G.r2pad = act ? G.rpad * G.rpad : -1.f;
  if(!act) G.ax = G.ay = G.az = 1.0e18f;
  return G;
}

struct DensOut
{
  double *hsml_in, *numngb, *density, *dhsml, *divvel, *curlvel, *pressure, *sinkrho, *sinkentr, *sinkvel;
  int *pairs;
  float4 *e_F1;			// FP32 walk record of the particle's DFS entry (its softening length follows Hsml)
  const double *grav_scale;
};

struct DensAcc
{
  double rho, wngb, dhsmlrho, divv, rot0, rot1, rot2, sentr, gv0, gv1, gv2;
  int numngb;
  __device__ void zero() { rho = wngb = dhsmlrho = divv = rot0 = rot1 = rot2 = sentr = gv0 = gv1 = gv2 = 0; numngb = 0; }
};

// density.c:548-646 for one target: is the neighbour number inside the window?  If not, move the bracket and
// Hsml exactly as the reference does.  Returns true when the target has to be evaluated again.
__device__ __forceinline__ bool hsml_step(double wngb, double rho, double dhsmlrho, int ptype, double desnumngb,
					  const gl_dparams &par, double &h, double &left, double &right,
					  unsigned long long *counter)
{
  bool redo = (wngb < (desnumngb - par.maxdev)) || (wngb > (desnumngb + par.maxdev) && h > (1.01 * par.mingashsml));
  if(redo && left > 0 && right > 0 && (right - left) < 1.0e-3 * left) redo = false;	// density.c:566-573
  if(!redo) return false;
  double dfac = 1.0;
  if(ptype == 0 && rho > 0)
    {
      dfac = dhsmlrho * (h / (3 * rho));	// density.c:442
      dfac = (dfac > -0.9) ? 1 / (1 + dfac) : 1;
    }
  if(wngb < (desnumngb - par.maxdev))
    left = fmax(h, left);
  else
    {
      if(right != 0)
	{
	  if(h < right) right = h;
	}
      else
	right = h;
    }
  if(right > 0 && left > 0)
    h = pow(0.5 * (pow(left, 3.0) + pow(right, 3.0)), 1.0 / 3);
  else
    {
      if(right == 0 && left == 0) atomicMax(&counter[2], 8188ull);	// endrun(8188), density.c:608
      if(right == 0 && left > 0)
	{
	  if(ptype == 0 && fabs(wngb - desnumngb) < 0.5 * desnumngb)
	    {
	      const double fac = 1 - (wngb - desnumngb) / (3 * wngb) * dfac;
	      h *= (fac < 1.26) ? fac : 1.26;
	    }
	  else
	    h *= 1.26;
	}
      if(right > 0 && left == 0)
	{
	  if(ptype == 0 && fabs(wngb - desnumngb) < 0.5 * desnumngb)
	    {
	      const double fac = 1 - (wngb - desnumngb) / (3 * wngb) * dfac;
	      if(fac > 1 / 1.26)
		h *= fac;
	      else
		h /= 1.26;
	    }
	  else
	    h /= 1.26;
	}
    }
  if(h < par.mingashsml) h = par.mingashsml;
  return true;
}

struct DensIn
{
  const int *noff;
  const double *in_entropy, *in_dtentropy;
  const int *in_timebin, *in_tibegstep;
};

// final store of one converged target: density.c:434-546 post-processing + the P[]/SphP[] fields
__device__ __forceinline__ void density_store(const DensAcc &A, double h, int p, int pin, int ptype, const DensOut &out,
					      const DensIn &in, double4 *__restrict__ s_vh, double4 *__restrict__ e_B,
					      const gl_dparams &par)
{
  const double rho = A.rho;
  out.hsml_in[pin] = h;
  s_vh[p].w = h;
  out.numngb[pin] = A.wngb;
  out.pairs[pin] = A.numngb;
  if(ptype == 0)
    {
      e_B[(long long) p + in.noff[p + 1]].x = h;	// live Hsml for the gravity softening, forcetree.c:2046-2054
      out.e_F1[(long long) p + in.noff[p + 1]].w = (float) (h * out.grav_scale[0]);
      double dfac = A.dhsmlrho, curl = 0, dv = A.divv;
      if(rho > 0)
	{
	  dfac *= h / (3 * rho);
	  dfac = (dfac > -0.9) ? 1 / (1 + dfac) : 1;
	  curl = sqrt(A.rot0 * A.rot0 + A.rot1 * A.rot1 + A.rot2 * A.rot2) / rho;
	  dv = A.divv / rho;
	}
      const int tb = in.in_timebin[pin];
      const int dt_step = tb ? (1 << tb) : 0;
      const double dt_entr = (par.ti_current - (in.in_tibegstep[pin] + dt_step / 2)) * par.timebase_interval;
      out.density[pin] = rho;
      out.dhsml[pin] = dfac;
      out.divvel[pin] = dv;
      out.curlvel[pin] = curl;
      out.pressure[pin] = (in.in_entropy[pin] + in.in_dtentropy[pin] * dt_entr) * pow(rho, GL_GAMMA);
    }
  else
    {				// density.c:520-546 sink (Type 5) / dust (Type 2) post-processing
      double se = A.sentr, g0 = A.gv0, g1 = A.gv1, g2 = A.gv2;
      if(rho > 0)
	{
	  se /= rho; g0 /= rho; g1 /= rho; g2 /= rho;
	}
      out.sinkrho[pin] = rho;
      out.sinkentr[pin] = se;
      out.sinkvel[3ll * pin] = g0;
      out.sinkvel[3ll * pin + 1] = g1;
      out.sinkvel[3ll * pin + 2] = g2;
    }
}

// kernel value + derivative and the density.c:862-913 sums for one neighbour
__device__ __forceinline__ void density_pair(DensAcc &P, int ttype, double dx, double dy, double dz, double r2, double mass_j,
					     double hinv, double hinv3, double hinv4, double tvx, double tvy, double tvz,
					     const double4 &vj, double entr_j)
{
  P.numngb++;
  // r and 1/r from one reciprocal square root (r2 == 0 is the target itself: r = 0, no velocity terms)
  const double rinv = (r2 > 0) ? rsqrt(r2) : 0.0;
  const double r = r2 * rinv;
  const double u = r * hinv;
  double wk, dwk;
  if(u < 0.5)
    {
      wk = hinv3 * (KERNEL_COEFF_1 + KERNEL_COEFF_2 * (u - 1) * u * u);
      dwk = hinv4 * u * (KERNEL_COEFF_3 * u - KERNEL_COEFF_4);
    }
  else
    {
      wk = hinv3 * KERNEL_COEFF_5 * (1.0 - u) * (1.0 - u) * (1.0 - u);
      dwk = hinv4 * KERNEL_COEFF_6 * (1.0 - u) * (1.0 - u);
    }
  P.rho += mass_j * wk;
  P.wngb += NORM_COEFF * wk / hinv3;
  P.dhsmlrho += -mass_j * (3 * hinv * wk + u * dwk);
  if(ttype != 0)
    {				// only sink / dust targets keep these (density.c:944-988)
      P.gv0 += mass_j * wk * vj.x;
      P.gv1 += mass_j * wk * vj.y;
      P.gv2 += mass_j * wk * vj.z;
      P.sentr += mass_j * wk * entr_j;
    }
  else if(r > 0)
    {
      const double fac = mass_j * dwk * rinv;
      const double dvx = tvx - vj.x, dvy = tvy - vj.y, dvz = tvz - vj.z;
      P.divv += -fac * (dx * dvx + dy * dvy + dz * dvz);
      P.rot0 += fac * (dz * dvy - dy * dvz);
      P.rot1 += fac * (dx * dvz - dz * dvx);
      P.rot2 += fac * (dy * dvx - dx * dvy);
    }
}

// phases 2+3 for one staged chunk: every active target against all staged candidates
__device__ __forceinline__ void density_chunk(WarpStage &S, int ncand, float lmax, unsigned todo_mask, int lane, const GroupF &G,
					      double px, double py, double pz, double vx, double vy, double vz, double h,
					      int ptype, DensAcc &A, const double4 *__restrict__ s_pm,
					      const double4 *__restrict__ s_vh, const double *__restrict__ s_entropy)
{
  __syncwarp();
  float L = lmax;
  for(int o = 16; o > 0; o >>= 1) L = fmaxf(L, __shfl_xor_sync(FULL, L, o));
  L = fmaxf(L, G.amax);
  const float mythr = cand_threshold(G.r, L);
  for(unsigned rest = todo_mask; rest; rest &= rest - 1)
    {
      const int t = __ffs(rest) - 1;
      const float tax = __shfl_sync(FULL, G.ax, t), tay = __shfl_sync(FULL, G.ay, t), taz = __shfl_sync(FULL, G.az, t);
      const float thr = __shfl_sync(FULL, mythr, t);
      // phase 2 (FP32, conservative): compact the candidates that MAY be neighbours of target t
      int nh = 0;
      DensAcc P;
      P.zero();
      bool loaded = false;
      double tx = 0, ty = 0, tz = 0, tvx = 0, tvy = 0, tvz = 0, h2 = 0, hinv = 0, hinv3 = 0, hinv4 = 0;
      int ttype = 0;
      for(int base = 0;; base += 32)
	{
	  if(base < ncand)
	    {
	      const int c = base + lane;
	      bool hit = false;
	      if(c < ncand)
		{
		  const float4 cd = S.cand[c];
		  const float dx = tax - cd.x, dy = tay - cd.y, dz = taz - cd.z;
		  hit = (dx * dx + dy * dy + dz * dz) <= thr;
		}
	      const unsigned ball = __ballot_sync(FULL, hit);
	      if(hit) S.hit[nh + __popc(ball & ((1u << lane) - 1))] = c;
	      nh += __popc(ball);
	      __syncwarp();
	      if(nh < 32 && base + 32 < ncand) continue;	// keep filling until a full warp of hits is available
	    }
	  if(nh == 0) break;
	  if(!loaded)
	    {
	      tx = __shfl_sync(FULL, px, t); ty = __shfl_sync(FULL, py, t); tz = __shfl_sync(FULL, pz, t);
	      tvx = __shfl_sync(FULL, vx, t); tvy = __shfl_sync(FULL, vy, t); tvz = __shfl_sync(FULL, vz, t);
	      const double th = __shfl_sync(FULL, h, t);
	      ttype = __shfl_sync(FULL, ptype, t);
	      h2 = th * th;
	      hinv = 1.0 / th;
	      hinv3 = hinv * hinv * hinv;
	      hinv4 = hinv3 * hinv;
	      loaded = true;
	    }
	  // phase 3: one candidate per lane, the reference's FP64 test and sums
	  const int take = nh < 32 ? nh : 32;
	  if(lane < take)
	    {
	      const int j = S.j[S.hit[lane]];
	      const double4 pj = gl_ld4(&s_pm[j]);
	      // density.c:835-856: dx = pos_i - pos_j, r2 = dx*dx + dy*dy + dz*dz, r2 < h2
	      const double dx = tx - pj.x, dy = ty - pj.y, dz = tz - pj.z;
	      const double r2 = dx * dx + dy * dy + dz * dz;
	      if(r2 < h2)
		{
		  const double4 vj = gl_ld4(&s_vh[j]);
		  density_pair(P, ttype, dx, dy, dz, r2, pj.w, hinv, hinv3, hinv4, tvx, tvy, tvz, vj,
			       (ttype != 0) ? __ldg(&s_entropy[j]) : 0.0);
		}
	    }
	  hits_shift(S, nh, lane);
	  nh -= take;
	  if(base + 32 >= ncand && nh == 0) break;
	}
      // warp-shuffle reduction of the partial sums into the owner lane
      const int tot = __reduce_add_sync(FULL, P.numngb);
      if(tot == 0) continue;
      const double r_rho = warp_sum_d(P.rho), r_w = warp_sum_d(P.wngb), r_d = warp_sum_d(P.dhsmlrho);
      double r4, r5, r6, r7;
      if(ttype != 0)
	{
	  r4 = warp_sum_d(P.gv0); r5 = warp_sum_d(P.gv1); r6 = warp_sum_d(P.gv2); r7 = warp_sum_d(P.sentr);
	}
      else
	{
	  r4 = warp_sum_d(P.divv); r5 = warp_sum_d(P.rot0); r6 = warp_sum_d(P.rot1); r7 = warp_sum_d(P.rot2);
	}
      if(lane == t)
	{
	  A.rho += r_rho; A.wngb += r_w; A.dhsmlrho += r_d; A.numngb += tot;
	  if(ttype != 0)
	    {
	      A.gv0 += r4; A.gv1 += r5; A.gv2 += r6; A.sentr += r7;
	    }
	  else
	    {
	      A.divv += r4; A.rot0 += r5; A.rot1 += r6; A.rot2 += r7;
	    }
	}
    }
  __syncwarp();
}


// Phases 2+3, lane = target.  Used while most of the warp is still active: every lane scans the staged candidates
// (broadcast shared-memory reads, FP32 conservative filter, 3 FMAs per test) and appends the tree indices its own
// target may touch to a PRIVATE RING in shared memory; the reference's FP64 test and the sums ("heavy" step) are
// taken from that ring - no broadcasts, no cross-lane reduction, and every lane adds its neighbours straight into
// its own accumulators.  The ring outlives the staged chunk, so the heavy steps are not run chunk by chunk (where
// most lanes had nothing to do in most steps: 12 of 32 lanes busy in the ncu source view) but only when some lane's
// ring is nearly full, and then every lane with work takes part.  (When few targets are left, density_chunk's
// target-major scheme is cheaper.)
template <class Heavy>
__device__ __forceinline__ void q_drain(WarpStage &S, int lane, int &qh, const int qt, const int low, Heavy heavy)
{
  while(__any_sync(FULL, qt - qh > low))
    if(qt > qh)
      {
	const int j = S.q[(qh & (QCAP - 1)) * 32 + lane];
	qh++;
	heavy(j);
      }
}

template <bool SYM, class Heavy>
__device__ __forceinline__ void lt_chunk(WarpStage &S, int ncand, float lmax, int lane, const GroupF &G, int &qh, int &qt, Heavy heavy)
{
  __syncwarp();
  float L = lmax;
  for(int o = 16; o > 0; o >>= 1) L = fmaxf(L, __shfl_xor_sync(FULL, L, o));
  L = fmaxf(L, G.amax);
  const float thr = cand_threshold(G.r, L);
  char *const qbase = reinterpret_cast<char *>(S.q);
  // one candidate against MY target.  The append is branch-free: the slot at the ring's tail is written always and
  // only kept (tail advanced) when the candidate passed - so the ring must never be full when this runs.
  auto test = [&](int c, unsigned &qoff) {
    const float4 cd = S.cand[c];
    const int jj = S.j[c];
    const float dx = G.ax - cd.x, dy = G.ay - cd.y, dz = G.az - cd.z;
    const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
    *reinterpret_cast<int *>(qbase + qoff) = jj;
    // conservative FP32 version of r2 < h_i^2 (density.c:856), or of r2 < h_i^2 || r2 < h_j^2 (hydra.c:1266)
    const bool hit = r2 <= (SYM ? fmaxf(thr, cd.w) : thr);
    qt += hit ? 1 : 0;
    qoff = hit ? ((qoff + 128u) & (QCAP * 128u - 1u)) : qoff;
  };
  int c0 = 0;
  while(c0 < ncand)
    {
      // room for eight more entries in every ring?
      if(__any_sync(FULL, qt - qh > QCAP - 8)) q_drain(S, lane, qh, qt, QCAP / 2, heavy);
      unsigned qoff = ((unsigned) (qt & (QCAP - 1)) << 7) | ((unsigned) lane << 2);	// byte offset of the tail slot
      if(c0 + 8 <= ncand)
	{
#pragma unroll
	  for(int u = 0; u < 8; u++) test(c0 + u, qoff);
	}
      else
	for(int c = c0; c < ncand; c++) test(c, qoff);
      c0 += 8;
    }
  __syncwarp();
}

__global__ void __launch_bounds__(SPH_BLOCK, SPH_MIN_BLOCKS)
k_density(const int *__restrict__ tgt, long long nt, int nentries, const int *__restrict__ e_next,
	  double4 *__restrict__ e_B, const int2 *__restrict__ e_rng, const int *__restrict__ n_child,
	  const double4 *__restrict__ s_pm, double4 *__restrict__ s_vh, const int *__restrict__ s_type, const double *__restrict__ s_entropy,
	  const int *__restrict__ perm, DensIn in, DensOut out, int *__restrict__ fat_p, double4 *__restrict__ fat_state,
	  unsigned long long *__restrict__ counter, gl_dparams par)
{
  extern __shared__ __align__(16) unsigned char gl_smem_raw[];
  WarpStage *stage = reinterpret_cast<WarpStage *>(gl_smem_raw);
  WarpStage &S = stage[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long) blockDim.x + threadIdx.x) >> 5;
  const long long t = warp * 32 + lane;
  const bool live = t < nt;
  int p = 0, ptype = 0, pin = 0;
  double px = 0, py = 0, pz = 0, vx = 0, vy = 0, vz = 0, h = 1;
  if(live)
    {
      p = tgt[t];
      const double4 pm = s_pm[p];
      const double4 vh = s_vh[p];
      px = pm.x; py = pm.y; pz = pm.z;
      ptype = s_type[p];
      pin = perm[p];
      h = vh.w;
      if(ptype == 0)
	{			// density.c:772-780: gas targets use VelPred, sink/dust targets a zero vector
	  vx = vh.x; vy = vh.y; vz = vh.z;
	}
    }
  const double desnumngb = (ptype == 5) ? par.desnumngb * par.bhfac : par.desnumngb;	// density.c:548-557
  const long long clk0 = clock64();
  unsigned visits = 0, staged = 0;
  double left = 0, right = 0;
  bool todo = live;
  int iter = 0;
  DensAcc A;
  A.zero();
  // the pair count of the previous call tells which targets are "fat" before any work is spent on them
  bool defer = live && (out.pairs[pin] > FAT_THRESH), deferred = false;
  if(__any_sync(FULL, defer))
    {
      if(defer)
	{
	  const int slot = (int) atomicAdd(&counter[8], 1ull);
	  if(slot < FAT_MAX)
	    {
	      fat_p[slot] = p;
	      fat_state[slot] = make_double4(h, 0.0, 0.0, 0.0);
	      todo = false;
	      deferred = true;
	    }
	}
      defer = false;
    }

  int qh = 0, qt = 0, myrounds = 0;
  while(true)
    {
      const unsigned todo_mask = __ballot_sync(FULL, todo);
      if(todo_mask == 0) break;
      if(todo) A.zero();	// lanes that already converged keep the sums of their final round
      const GroupF G = make_group(todo, px, py, pz, h);
      const bool ltmode = __popc(todo_mask) >= LT_MIN_ACTIVE;
      const double h2 = h * h, hinv = 1.0 / h, hinv3 = hinv * hinv * hinv, hinv4 = hinv3 * hinv;
      // lane = target "heavy" step: the reference's FP64 test and sums for one queued candidate of MY target
      auto heavy = [&](int j) {
	const double4 pj = gl_ld4(&s_pm[j]);
	// density.c:835-856: dx = pos_i - pos_j, r2 = dx*dx + dy*dy + dz*dz, r2 < h2
	const double dx = px - pj.x, dy = py - pj.y, dz = pz - pj.z;
	const double r2 = dx * dx + dy * dy + dz * dz;
	if(r2 < h2)
	  {
	    const double4 vj = gl_ld4(&s_vh[j]);
	    density_pair(A, ptype, dx, dy, dz, r2, pj.w, hinv, hinv3, hinv4, vx, vy, vz, vj, (ptype != 0) ? __ldg(&s_entropy[j]) : 0.0);
	  }
      };
      sph_collect<false>(S, lane, nentries, e_next, e_B, nullptr, e_rng, n_child, s_pm, s_vh, s_type, todo_mask, G, visits, staged, counter,
			 [&](int ncand, float lmax) {
			   if(ltmode)
			     lt_chunk<false>(S, ncand, lmax, lane, G, qh, qt, heavy);
			   else
			     density_chunk(S, ncand, lmax, todo_mask, lane, G, px, py, pz, vx, vy, vz, h, ptype, A, s_pm, s_vh, s_entropy);
			 });
      if(ltmode) q_drain(S, lane, qh, qt, 0, heavy);

      // ---- post-processing and Hsml update, density.c:434-652 (one lane = one target) ----
      if(todo)
	{
	  myrounds++;
	  todo = hsml_step(A.wngb, A.rho, A.dhsmlrho, ptype, desnumngb, par, h, left, right, counter);
	  if(todo && A.numngb > FAT_THRESH) defer = true;
	}
      iter++;
      if(defer && todo)
	{
	  // a target with thousands of neighbours (a sink in a disc hole) would keep this warp busy for the rest of
	  // the launch: hand it to the block-per-target path, which continues its iteration where it stands
	  const int slot = (int) atomicAdd(&counter[8], 1ull);
	  if(slot < FAT_MAX)
	    {
	      fat_p[slot] = p;
	      fat_state[slot] = make_double4(h, left, right, (double) iter);
	      todo = false;
	      deferred = true;
	    }
	  defer = false;
	}
      if(iter > GL_MAXITER)
	{			// density.c:669-674 endrun(1155)
	  if(todo) atomicMax(&counter[2], 1155ull);
	  break;
	}
    }

  if(live && !deferred) density_store(A, h, p, pin, ptype, out, in, s_vh, e_B, par);
  const unsigned maxit = __reduce_max_sync(FULL, (unsigned) iter);
  const unsigned wrounds = __reduce_add_sync(FULL, (unsigned) myrounds);
  // which target needed the most passes (diagnostics: the analogue of the reference's "ngb iteration" lines, density.c:661)
  const unsigned mostr = __reduce_max_sync(FULL, (live && !deferred) ? (unsigned) myrounds : 0u);
  if(live && !deferred && mostr > 1 && (unsigned) myrounds == mostr)
    atomicMax(&counter[12], ((unsigned long long) mostr << 32) | (unsigned long long) (unsigned) pin);
  const unsigned wpairs = __reduce_add_sync(FULL, (live && !deferred) ? (unsigned) A.numngb : 0u);
  if(lane == 0)
    {
      atomicAdd(&counter[9], (unsigned long long) wpairs);	// pairs of the final round (= the integer neighbour counts)
      atomicAdd(&counter[11], (unsigned long long) wrounds);	// target-passes: every evaluation of every target
      // load-balance diagnostics (the analogue of the reference's work-load balance line, gravtree.c:812)
      const unsigned long long cyc = (unsigned long long) (clock64() - clk0);
      atomicMax(&counter[1], (unsigned long long) maxit);
      atomicMax(&counter[3], cyc);
      atomicAdd(&counter[4], cyc);
      atomicMax(&counter[5], (unsigned long long) staged);
      atomicAdd(&counter[6], (unsigned long long) staged);
      atomicMax(&counter[7], (unsigned long long) visits);
    }
}

__global__ void k_gather_d(long long nit, long long n, const int *__restrict__ perm, const double *__restrict__ in, double *__restrict__ out)
{
  long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= nit) return;
  const int j = perm[i];
  out[i] = (j < n) ? in[j] : 0.0;	// (items >= n: cut cells of the multi-rank mode)
}

// ---------------------------------------------------------------- fat targets ----
// A target with thousands of neighbours (e.g. the central sink, whose 160 kernel-weighted neighbours
// sit at the rim of a ~22k-particle sphere) is a long serial tail for one warp.  It is handed to this
// path instead: (1) ALL SMs brute-force the gas particles once and append those inside a guard radius
// Rc to a compact candidate list, (2) one 1024-thread block per target runs the reference's Hsml
// iteration over that list.  If Hsml leaves the guard radius the host enlarges Rc and repeats.
struct FatTarget
{
  double x, y, z, rc;
};

__global__ void k_fat_scan(long long n, const double4 *__restrict__ s_pm, const double4 *__restrict__ s_vh,
			   const int *__restrict__ s_type, const double *__restrict__ s_entropy, int nf,
			   const FatTarget *__restrict__ ft, int *__restrict__ cnt, const int *__restrict__ off,
			   double4 *__restrict__ rec, const double4 *__restrict__ e_B, const int *__restrict__ noff,
			   unsigned long long *__restrict__ counter)
{
  __shared__ FatTarget sh[8];
  if(threadIdx.x < nf) sh[threadIdx.x] = ft[threadIdx.x];
  __syncthreads();
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= n) return;
  if(s_type[i] == -1)
    {
      // cut cell (multi-rank mode): the gas below it is not here - a guard sphere that reaches it is a violation
      const double4 B = e_B[i + noff[i + 1]];
      for(int f = 0; f < nf; f++)
	if(sh[f].rc > 0 && dist2_to_cube(sh[f].x, sh[f].y, sh[f].z, B, 0.5 * B.w) < sh[f].rc * sh[f].rc * 1.0001)
	  atomicMax(&counter[2], GL_ERR_LET_VIOLATION);
      return;
    }
  if(s_type[i] != 0) return;
  const double4 pm = s_pm[i];
  if(pm.w == 0) return;
  for(int f = 0; f < nf; f++)
    {
      const double dx = sh[f].x - pm.x, dy = sh[f].y - pm.y, dz = sh[f].z - pm.z;
      if(dx * dx + dy * dy + dz * dz < sh[f].rc * sh[f].rc)
	{
	  const int k = atomicAdd(&cnt[f], 1);
	  if(rec)
	    {
	      const double4 vh = s_vh[i];
	      rec[2ll * (off[f] + k)] = pm;
	      rec[2ll * (off[f] + k) + 1] = make_double4(vh.x, vh.y, vh.z, s_entropy[i]);
	    }
	}
    }
}

#define FAT_THREADS 1024
#define FAT_CLUSTER 8		// CTAs (= SMs) that share one fat target
// One THREAD-BLOCK CLUSTER per fat target: the candidate list is dealt out over 8 CTAs x 1024 threads, the partial
// sums of a round meet in the shared memory of cluster rank 0 (distributed shared memory writes), rank 0 takes the
// reference's Hsml step (density.c:548-652) and posts the new Hsml back into every CTA's shared memory; two
// cluster barriers per round.  (One CTA per target made the central sink of the 1.6M disc - 11 rounds over a
// ~50 000-entry list - a 0.9 ms serial tail on a single SM.)
__global__ void __cluster_dims__(FAT_CLUSTER, 1, 1) __launch_bounds__(FAT_THREADS)
k_fat_rounds(int nf, const int *__restrict__ fat_p, double4 *__restrict__ fat_state, int *__restrict__ fat_flag,
	     const FatTarget *__restrict__ ft, const int *__restrict__ cnt, const int *__restrict__ off,
	     const double4 *__restrict__ rec, const double4 *__restrict__ s_pm, double4 *__restrict__ s_vh,
	     const int *__restrict__ s_type, const int *__restrict__ perm, double4 *__restrict__ e_B, DensIn in, DensOut out,
	     unsigned long long *__restrict__ counter, gl_dparams par)
{
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double red[FAT_THREADS / 32][8];
  __shared__ int redn[FAT_THREADS / 32];
  __shared__ double part[FAT_CLUSTER][8];	// meaningful in rank 0: one row of partial sums per CTA of the cluster
  __shared__ int partn[FAT_CLUSTER];
  __shared__ double sh_h;
  __shared__ int sh_go;
  const int f = blockIdx.x / FAT_CLUSTER;	// every CTA of a cluster takes the same branch below
  const unsigned rank = cluster.block_rank();
  if(f >= nf) return;
  if(fat_flag[f] == 1) return;	// already converged in an earlier sweep
  const int p = fat_p[f];
  const int ptype = s_type[p], pin = perm[p];
  const double4 pm = s_pm[p];
  const double4 vh = s_vh[p];
  const double px = pm.x, py = pm.y, pz = pm.z;
  const double vx = (ptype == 0) ? vh.x : 0, vy = (ptype == 0) ? vh.y : 0, vz = (ptype == 0) ? vh.z : 0;
  const double desnumngb = (ptype == 5) ? par.desnumngb * par.bhfac : par.desnumngb;
  const double rc = ft[f].rc;
  const int nrec = cnt[f];
  const double4 *my = rec + 2ll * off[f];
  double4 st = fat_state[f];
  double h = st.x, left = st.y, right = st.z;
  int iter = (int) st.w;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double *part0 = cluster.map_shared_rank(&part[0][0], 0);
  int *partn0 = cluster.map_shared_rank(&partn[0], 0);
  while(true)
    {
      const double h2 = h * h, hinv = 1.0 / h, hinv3 = hinv * hinv * hinv, hinv4 = hinv3 * hinv;
      DensAcc P;
      P.zero();
      for(int k = rank * FAT_THREADS + threadIdx.x; k < nrec; k += FAT_CLUSTER * FAT_THREADS)
	{
	  const double4 a = my[2ll * k];
	  const double dx = px - a.x, dy = py - a.y, dz = pz - a.z;
	  const double r2 = dx * dx + dy * dy + dz * dz;
	  if(r2 < h2)
	    {
	      const double4 b = my[2ll * k + 1];
	      density_pair(P, ptype, dx, dy, dz, r2, a.w, hinv, hinv3, hinv4, vx, vy, vz, b, b.w);
	    }
	}
      double v[8] = { P.rho, P.wngb, P.dhsmlrho, (ptype != 0) ? P.gv0 : P.divv, (ptype != 0) ? P.gv1 : P.rot0,
	(ptype != 0) ? P.gv2 : P.rot1, (ptype != 0) ? P.sentr : P.rot2, 0.0 };
      for(int q = 0; q < 7; q++) v[q] = warp_sum_d(v[q]);
      const int cn = __reduce_add_sync(FULL, P.numngb);
      if(lane == 0)
	{
	  for(int q = 0; q < 7; q++) red[w][q] = v[q];
	  redn[w] = cn;
	}
      __syncthreads();
      if(threadIdx.x < 8)
	{
	  // this CTA's partial sums -> row `rank` of cluster rank 0's table
	  double sum = 0;
	  int sn = 0;
	  for(int ww = 0; ww < FAT_THREADS / 32; ww++)
	    {
	      if(threadIdx.x < 7) sum += red[ww][threadIdx.x];
	      else sn += redn[ww];
	    }
	  if(threadIdx.x < 7) part0[rank * 8 + threadIdx.x] = sum;
	  else partn0[rank] = sn;
	}
      cluster.sync();
      if(rank == 0 && threadIdx.x == 0)
	{
	  DensAcc A;
	  A.zero();
	  for(int r = 0; r < FAT_CLUSTER; r++)
	    {
	      A.rho += part[r][0]; A.wngb += part[r][1]; A.dhsmlrho += part[r][2];
	      if(ptype != 0)
		{
		  A.gv0 += part[r][3]; A.gv1 += part[r][4]; A.gv2 += part[r][5]; A.sentr += part[r][6];
		}
	      else
		{
		  A.divv += part[r][3]; A.rot0 += part[r][4]; A.rot1 += part[r][5]; A.rot2 += part[r][6];
		}
	      A.numngb += partn[r];
	    }
	  const double h_eval = h;
	  const bool redo = hsml_step(A.wngb, A.rho, A.dhsmlrho, ptype, desnumngb, par, h, left, right, counter);
	  iter++;
	  int go = 1;
	  if(!redo)
	    {
	      density_store(A, h_eval, p, pin, ptype, out, in, s_vh, e_B, par);
	      fat_flag[f] = 1;
	      go = 0;
	    }
	  else if(iter > GL_MAXITER)
	    {
	      atomicMax(&counter[2], 1155ull);
	      fat_flag[f] = 1;
	      go = 0;
	    }
	  else if(h * BOX_SLACK >= rc)
	    {
	      fat_flag[f] = 2;	// Hsml left the guard radius: the host rebuilds the candidate list
	      go = 0;
	    }
	  fat_state[f] = make_double4(h, left, right, (double) iter);
	  atomicMax(&counter[1], (unsigned long long) iter);
	  atomicMax(&counter[12], ((unsigned long long) (unsigned) iter << 32) | (unsigned long long) (unsigned) pin);
	  atomicAdd(&counter[11], 1ull);
	  if(!redo) atomicAdd(&counter[9], (unsigned long long) A.numngb);
	  // post the decision into every CTA of the cluster
	  for(int r = 0; r < FAT_CLUSTER; r++)
	    {
	      *cluster.map_shared_rank(&sh_h, r) = h;
	      *cluster.map_shared_rank(&sh_go, r) = go;
	    }
	}
      cluster.sync();
      h = sh_h;
      if(!sh_go) break;
    }
  cluster.sync();		// nobody leaves while a peer may still touch its shared memory
}

static int density_fat_path(gl_ctx *c, int nfat, const DensIn &in, const DensOut &out, const double *s_entropy,
			    const gl_dparams &dp)
{
  long long n = c->nit;	// tree-ordered items
  std::vector<int> hp(nfat), hflag(nfat, 0);
  std::vector<double4> hst(nfat);
  GL_CUDA(c, cudaMemcpyAsync(hp.data(), c->fat_p.p, nfat * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaMemcpyAsync(hst.data(), c->fat_state.p, nfat * sizeof(double4), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  std::vector<double4> hpos(nfat);
  for(int f = 0; f < nfat; f++)
    GL_CUDA(c, cudaMemcpyAsync(&hpos[f], c->s_pm.p + hp[f], sizeof(double4), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  GL_CUDA(c, c->fat_tgt.ensure(8 * 4));
  GL_CUDA(c, c->fat_i.ensure(64));
  for(int b0 = 0; b0 < nfat; b0 += 8)
    {
      const int nb = std::min(8, nfat - b0);
      std::vector<double> guard(nb, 1.3);
      for(int sweep = 0; sweep < 64; sweep++)
	{
	  FatTarget ht[8];
	  bool any = false;
	  for(int f = 0; f < nb; f++)
	    {
	      const double hh = std::max(hst[b0 + f].x, std::max(hst[b0 + f].y, hst[b0 + f].z));
	      ht[f] = FatTarget{hpos[b0 + f].x, hpos[b0 + f].y, hpos[b0 + f].z, hflag[b0 + f] == 1 ? 0.0 : guard[f] * hh};
	      any = any || hflag[b0 + f] != 1;
	    }
	  if(!any) break;
	  int hz[24] = { 0 };
	  GL_CUDA(c, cudaMemcpyAsync(c->fat_tgt.p, ht, sizeof(FatTarget) * nb, cudaMemcpyHostToDevice, c->stream));
	  GL_CUDA(c, cudaMemcpyAsync(c->fat_i.p, hz, sizeof(hz), cudaMemcpyHostToDevice, c->stream));
	  int *cnt = c->fat_i.p, *off = c->fat_i.p + 8, *flag = c->fat_i.p + 16;
	  k_fat_scan<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, c->s_pm.p, c->s_vh.p, c->s_type.p, s_entropy, nb,
							    (const FatTarget *) c->fat_tgt.p, cnt, off, nullptr, c->e_B.p, c->noff.p, c->d_counter.p);
	  GL_LAUNCH_COUNT(c);
	  int hc[8];
	  GL_CUDA(c, cudaMemcpyAsync(hc, cnt, sizeof(int) * nb, cudaMemcpyDeviceToHost, c->stream));
	  GL_CUDA(c, cudaStreamSynchronize(c->stream));
	  int ho[8], tot = 0;
	  for(int f = 0; f < nb; f++)
	    {
	      ho[f] = tot;
	      tot += hc[f];
	    }
	  GL_CUDA(c, c->fat_rec.ensure(2 * (size_t) tot + 2));
	  int hfl[8];
	  for(int f = 0; f < nb; f++) hfl[f] = hflag[b0 + f] == 1 ? 1 : 0;
	  GL_CUDA(c, cudaMemsetAsync(cnt, 0, 8 * sizeof(int), c->stream));
	  GL_CUDA(c, cudaMemcpyAsync(off, ho, sizeof(int) * nb, cudaMemcpyHostToDevice, c->stream));
	  GL_CUDA(c, cudaMemcpyAsync(flag, hfl, sizeof(int) * nb, cudaMemcpyHostToDevice, c->stream));
	  k_fat_scan<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, c->s_pm.p, c->s_vh.p, c->s_type.p, s_entropy, nb,
							    (const FatTarget *) c->fat_tgt.p, cnt, off, c->fat_rec.p, c->e_B.p, c->noff.p, c->d_counter.p);
	  GL_LAUNCH_COUNT(c);
	  k_fat_rounds<<<nb * FAT_CLUSTER, FAT_THREADS, 0, c->stream>>>(nb, c->fat_p.p + b0, c->fat_state.p + b0, flag,
							  (const FatTarget *) c->fat_tgt.p, cnt, off, c->fat_rec.p, c->s_pm.p,
							  c->s_vh.p, c->s_type.p, c->perm.p, c->e_B.p, in, out, c->d_counter.p, dp);
	  GL_LAUNCH_COUNT(c);
	  GL_CUDA(c, cudaMemcpyAsync(hfl, flag, sizeof(int) * nb, cudaMemcpyDeviceToHost, c->stream));
	  GL_CUDA(c, cudaMemcpyAsync(&hst[b0], c->fat_state.p + b0, sizeof(double4) * nb, cudaMemcpyDeviceToHost, c->stream));
	  GL_CUDA(c, cudaStreamSynchronize(c->stream));
	  GL_CUDA(c, cudaGetLastError());
	  for(int f = 0; f < nb; f++)
	    {
	      if(hfl[f] == 2) guard[f] = 1.6;	// Hsml ran out of the guard sphere: take a wider one around the new value
	      hflag[b0 + f] = hfl[f] == 1 ? 1 : 0;
	    }
	}
      for(int f = 0; f < nb; f++)
	if(hflag[b0 + f] != 1) return c->fail("density: fat-target path did not converge");
    }
  return 0;
}

int gls_density(gl_ctx *c)
{
  if(!c->tree_valid)		// the reference builds it beforehand (ngb_treebuild, init.c:529); a fresh load has none yet
    if(glt_build(c)) return 1;
  if(c->par.ComovingIntegrationOn) return c->fail("density: ComovingIntegrationOn != 0 is not implemented");
  if(gl_prepare_targets(c)) return 1;
  long long n = c->n, nt = c->n_tgt_sph;
  GL_CUDA(c, c->o_numngb.ensure_keep(n));
  GL_CUDA(c, c->o_density.ensure_keep(n));
  GL_CUDA(c, c->o_dhsml.ensure_keep(n));
  GL_CUDA(c, c->o_divvel.ensure_keep(n));
  GL_CUDA(c, c->o_curlvel.ensure_keep(n));
  GL_CUDA(c, c->o_pressure.ensure_keep(n));
  GL_CUDA(c, c->o_sinkrho.ensure_keep(n));
  GL_CUDA(c, c->o_sinkentr.ensure_keep(n));
  GL_CUDA(c, c->o_sinkvel.ensure_keep(3 * n));
  if(c->o_denspairs.cap < (size_t) n)
    {
      GL_CUDA(c, c->o_denspairs.ensure(n));
      GL_CUDA(c, cudaMemsetAsync(c->o_denspairs.p, 0, c->o_denspairs.cap * sizeof(int), c->stream));	// no history: nobody is fat yet
    }
  GL_CUDA(c, c->d_counter.ensure(16));
  GL_CUDA(c, c->h_q.ensure(c->nit + 1));
  GL_CUDA(c, c->h_r.ensure(c->nit + 1));	// scratch: the tree-ordered entropy lives here until hydro overwrites it
  double *s_entropy = (double *) c->h_r.p;
  GL_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, 16 * sizeof(unsigned long long), c->stream));
  cudaEventRecord(c->ev[0], c->stream);
  k_gather_d<<<gl_grid(c->nit, 256), 256, 0, c->stream>>>(c->nit, n, c->perm.p, c->in_entropy.p, s_entropy);
  GL_LAUNCH_COUNT(c);
  DensOut out{c->in_hsml.p, c->o_numngb.p, c->o_density.p, c->o_dhsml.p, c->o_divvel.p, c->o_curlvel.p, c->o_pressure.p,
	      c->o_sinkrho.p, c->o_sinkentr.p, c->o_sinkvel.p, c->o_denspairs.p, c->e_F1.p, c->grav_scale.p};
  DensIn in{c->noff.p, c->in_entropy.p, c->in_dtentropy.p, c->in_timebin.p, c->in_tibegstep.p};
  gl_dparams dp = gl_make_dparams(c);
  GL_CUDA(c, c->fat_p.ensure(FAT_MAX));
  GL_CUDA(c, c->fat_state.ensure(FAT_MAX));
  if(nt > 0)
    {
      if(!c->attr_density)	// per context (= per device), not per process
	{
	  GL_CUDA(c, cudaFuncSetAttribute(k_density, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (SPH_WARPS * sizeof(WarpStage))));
	  c->attr_density = true;
	}
      k_density<<<gl_grid(nt, SPH_BLOCK), SPH_BLOCK, SPH_WARPS * sizeof(WarpStage), c->stream>>>(
	c->tgt_sph.p, nt, (int) c->nentries, c->e_next.p, c->e_B.p, c->e_rng.p, c->n_child.p, c->s_pm.p, c->s_vh.p, c->s_type.p,
	s_entropy, c->perm.p, in, out, c->fat_p.p, c->fat_state.p, c->d_counter.p, dp);
      GL_LAUNCH_COUNT(c);
      c->fire_post_launch();
      unsigned long long nfat = 0;
      GL_CUDA(c, cudaMemcpyAsync(&nfat, c->d_counter.p + 8, sizeof(nfat), cudaMemcpyDeviceToHost, c->stream));
      GL_CUDA(c, c->wait_stream());
      GL_CUDA(c, cudaGetLastError());
      if(nfat > 0)
	if(density_fat_path(c, (int) std::min<unsigned long long>(nfat, FAT_MAX), in, out, s_entropy, dp)) return 1;
      c->st.density_fat_targets = (long long) nfat;
    }
  cudaEventRecord(c->ev[1], c->stream);
  unsigned long long cnt[13];
  GL_CUDA(c, cudaMemcpyAsync(cnt, c->d_counter.p, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  GL_CUDA(c, cudaGetLastError());
  float ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->st.ms_density = ms;
  c->st.density_targets = nt;
  c->st.density_passes_max = (long long) cnt[1];
  c->st.density_pairs = (long long) cnt[9];
  c->st.density_target_passes = (long long) cnt[11];
  c->st.density_slowest_index = (cnt[12] >> 32) ? (long long) (cnt[12] & 0xffffffffull) : -1;
  c->st.density_slowest_passes = (long long) (cnt[12] >> 32);
  if(getenv("GL_DEBUG_SPH"))
    fprintf(stderr, "[gl] density: %lld warps, cycles max %.3g avg %.3g | staged max %llu avg %.1f | entry visits max %llu | rounds max %llu | %.2f ms\n",
	    (nt + 31) / 32, (double) cnt[3], (double) cnt[4] / ((nt + 31) / 32), cnt[5], (double) cnt[6] / ((nt + 31) / 32), cnt[7], cnt[1], ms);
  if(cnt[2] == GL_ERR_LET_VIOLATION) return c->fail("density: GL_ERR_LET_VIOLATION");
  if(cnt[2] == 1155) return c->fail("density: failed to converge in neighbour iteration (endrun(1155), density.c:669)");
  if(cnt[2] == 8188) return c->fail("density: Left == Right == 0 (endrun(8188), density.c:608)");
  return 0;
}

// ---------------------------------------------------------------- hydro ----
__global__ void k_hydro_prep(long long n, const int *__restrict__ perm, const int *__restrict__ s_type,
			     const double4 *__restrict__ s_vh, const double *__restrict__ density,
			     const double *__restrict__ pressure, const double *__restrict__ dhsml,
			     const double *__restrict__ divvel, const double *__restrict__ curlvel,
			     const int *__restrict__ in_timebin, double4 *__restrict__ h_q, double4 *__restrict__ h_r)
{
  long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= n) return;
  if(s_type[i] != 0)
    {
      h_q[i] = make_double4(0, 0, 0, 0);
      h_r[i] = make_double4(0, 0, 0, 0);
      return;
    }
  const int j = perm[i];
  const double rho = density[j], P = pressure[j];
  // hydra.c:1271-1273: p_over_rho2_j and soundspeed_j are per-source quantities
  const double porho2 = P / (rho * rho);
  const double cs = sqrt(GL_GAMMA * porho2 * rho);
  const double adv = fabs(divvel[j]);
  // hydra.c:1531-1534 (fac_mu = 1): f2 = |divv| / (|divv| + curl + 0.0001 c / fac_mu / h)
  const double f = adv / (adv + curlvel[j] + 0.0001 * cs / 1.0 / s_vh[i].w);
  const int tb = in_timebin[j];
  h_q[i] = make_double4(rho, porho2, porho2 * dhsml[j], cs);	// hydra.c:1600: p_over_rho2_j *= DhsmlDensityFactor
  // 1/h_j and 1/h_j^4 exactly as hydra.c:1311-1312 forms them, once per source instead of once per pair
  const double hinv = 1.0 / s_vh[i].w;
  h_r[i] = make_double4(f, (double) (tb ? (1 << tb) : 0), hinv, hinv * hinv * hinv * hinv);
}

struct HydroAcc
{
  double a0, a1, a2, dtE, maxsig;
  int npairs;
  __device__ void zero() { a0 = a1 = a2 = dtE = maxsig = 0; npairs = 0; }
};

struct HydroTarget
{
  double px, py, pz, vx, vy, vz, h_i, mass, rho, f1, cs_i, p_over_rho2_i;
  int timestep;
};

// hydra.c:958-979 target load (per-target quantities formed once: sound speed, Balsara factor, P/rho^2 * Omega)
__device__ __forceinline__ HydroTarget hydro_load_target(bool live, int p, int pin, const double4 *__restrict__ s_pm,
							 const double4 *__restrict__ s_vh, const double4 *__restrict__ h_q,
							 const double4 *__restrict__ h_r, const double *__restrict__ pressure,
							 const double *__restrict__ dhsml, const double *__restrict__ divvel,
							 const double *__restrict__ curlvel)
{
  HydroTarget T;
  T.px = T.py = T.pz = T.vx = T.vy = T.vz = T.h_i = T.mass = T.f1 = T.cs_i = T.p_over_rho2_i = 0;
  T.rho = 1;
  T.timestep = 0;
  if(live)
    {
      const double4 pm = s_pm[p], vh = s_vh[p];
      T.px = pm.x; T.py = pm.y; T.pz = pm.z; T.mass = pm.w;
      T.vx = vh.x; T.vy = vh.y; T.vz = vh.z; T.h_i = vh.w;
      T.rho = h_q[p].x;
      const double P = pressure[pin];
      T.timestep = (int) h_r[p].y;
      T.cs_i = sqrt(GL_GAMMA * P / T.rho);
      const double adv = fabs(divvel[pin]);
      T.f1 = adv / (adv + curlvel[pin] + 0.0001 * T.cs_i / T.h_i / 1.0);
      T.p_over_rho2_i = P / (T.rho * T.rho);	// hydra.c:1154-1155
      T.p_over_rho2_i *= dhsml[pin];
    }
  return T;
}

struct HydroConst
{
  double h_i2, hinv_i, hinv4_i, artvisc, timebase_interval;
};

// hydra.c:1236-1621 for one candidate j of target T: the reference's FP64 predicate, then the pair force
__device__ __forceinline__ void hydro_pair(const HydroTarget &T, const HydroConst &K, HydroAcc &A, int j, const double4 *__restrict__ s_pm,
					   const double4 *__restrict__ s_vh, const double4 *__restrict__ h_q,
					   const double4 *__restrict__ h_r)
{
  const double4 pj = gl_ld4(&s_pm[j]);
  const double4 vj = gl_ld4(&s_vh[j]);
  const double dx = T.px - pj.x, dy = T.py - pj.y, dz = T.pz - pj.z;
  const double r2 = dx * dx + dy * dy + dz * dz;
  const double h_j = vj.w, mass_j = pj.w, h_j2 = h_j * h_j;
  if((r2 < K.h_i2 || r2 < h_j2) && r2 > 0)	// hydra.c:1266, :1270
    {
      const double4 q = gl_ld4(&h_q[j]);
      const double4 rr = gl_ld4(&h_r[j]);
      A.npairs++;
      // r and 1/r from one reciprocal square root; the reference divides by r three times
      // (hydra.c:1511, 1602, 1606), here these are multiplications by rinv (1 ulp apart)
      const double rinv = rsqrt(r2);
      const double r = r2 * rinv;
      const double rho_j = q.x, soundspeed_j = q.w;
      const double dvx = T.vx - vj.x, dvy = T.vy - vj.y, dvz = T.vz - vj.z;
      const double vdotr2 = dx * dvx + dy * dvy + dz * dvz;	// non-comoving, hydra.c:1279-1282
      double dwk_i = 0, dwk_j = 0;
      if(r2 < K.h_i2)
	{
	  const double u = r * K.hinv_i;
	  dwk_i = (u < 0.5) ? K.hinv4_i * u * (KERNEL_COEFF_3 * u - KERNEL_COEFF_4) : K.hinv4_i * KERNEL_COEFF_6 * (1.0 - u) * (1.0 - u);
	}
      if(r2 < h_j2)
	{
	  // 1/h_j and 1/h_j^4 come with the source record (k_hydro_prep forms them as hydra.c:1311-1312 does)
	  const double u = r * rr.z;
	  dwk_j = (u < 0.5) ? rr.w * u * (KERNEL_COEFF_3 * u - KERNEL_COEFF_4) : rr.w * KERNEL_COEFF_6 * (1.0 - u) * (1.0 - u);
	}
      double vsig = T.cs_i + soundspeed_j;	// hydra.c:1502
      if(vsig > A.maxsig) A.maxsig = vsig;
      double visc = 0;
      if(vdotr2 < 0)
	{
	  const double mu_ij = vdotr2 * rinv;	// fac_mu = 1
	  vsig -= 3 * mu_ij;
	  if(vsig > A.maxsig) A.maxsig = vsig;
	  const double rho_ij = 0.5 * (T.rho + rho_j);
	  visc = 0.25 * K.artvisc * vsig * (-mu_ij) / rho_ij * (T.f1 + rr.x);
	  // hydra.c:1583-1597 viscosity limiter
	  const int ts_j = (int) rr.y;
	  const double dt = 2 * ((T.timestep > ts_j) ? T.timestep : ts_j) * K.timebase_interval;
	  if(dt > 0 && (dwk_i + dwk_j) < 0)
	    {
	      const double lim = 0.5 * 1.0 * vdotr2 / (0.5 * (T.mass + mass_j) * (dwk_i + dwk_j) * r * dt);
	      visc = (visc < lim) ? visc : lim;
	    }
	}
      const double hfc_visc = 0.5 * mass_j * visc * (dwk_i + dwk_j) * rinv;
      const double hfc = hfc_visc + mass_j * (T.p_over_rho2_i * dwk_i + q.z * dwk_j) * rinv;	// q.z = P_j/rho_j^2 x Omega_j
      A.a0 += (-hfc * dx);
      A.a1 += (-hfc * dy);
      A.a2 += (-hfc * dz);
      A.dtE += (0.5 * hfc_visc * vdotr2);
    }
}

__device__ __forceinline__ void hydro_store(const HydroTarget &T, const HydroAcc &A, int pin, double *__restrict__ o_hacc,
					    double *__restrict__ o_dtentr, double *__restrict__ o_maxsig, int *__restrict__ o_pairs)
{
  o_hacc[3ll * pin] = A.a0;
  o_hacc[3ll * pin + 1] = A.a1;
  o_hacc[3ll * pin + 2] = A.a2;
  // hydra.c:583: DtEntropy *= GAMMA_MINUS1 / (hubble_a2 * rho^GAMMA_MINUS1)
  o_dtentr[pin] = A.dtE * (GL_GAMMA_MINUS1 / (1.0 * pow(T.rho, GL_GAMMA_MINUS1)));
  o_maxsig[pin] = A.maxsig;
  o_pairs[pin] = A.npairs;
}

#ifndef HYDRO_MIN_BLOCKS
#define HYDRO_MIN_BLOCKS 8
#endif
// Search and pair arithmetic live in ONE kernel on purpose.  A split variant was built and measured (search kernel at 61
// registers / 32 warps per SM writing per-target candidate lists, pair kernel reading them: bit-identical results): the
// search alone took 1.8 ms, but the pair kernel alone 3.8 ms against ~1 ms worth of the same work inside this kernel -
// the pair step is bound by the L1 gather rate (32 distinct lines per load instruction), and only pays off while other
// warps of the SM are in the ALU-bound search phase.  5.7 ms split vs 3.8 ms fused.
__global__ void __launch_bounds__(SPH_BLOCK, HYDRO_MIN_BLOCKS)
k_hydro(const int *__restrict__ tgt, long long nt, int nentries, const int *__restrict__ e_next,
	const double4 *__restrict__ e_B, const double *__restrict__ e_hmax, const int2 *__restrict__ e_rng,
	const int *__restrict__ n_child, const double4 *__restrict__ s_pm, const double4 *__restrict__ s_vh, const int *__restrict__ s_type,
	const double4 *__restrict__ h_q, const double4 *__restrict__ h_r, const int *__restrict__ perm,
	const double *__restrict__ pressure, const double *__restrict__ dhsml, const double *__restrict__ divvel,
	const double *__restrict__ curlvel, double *__restrict__ o_hacc, double *__restrict__ o_dtentr,
	double *__restrict__ o_maxsig, int *__restrict__ o_pairs, unsigned long long *__restrict__ counter, gl_dparams par)
{
  extern __shared__ __align__(16) unsigned char gl_smem_raw[];
  WarpStage *stage = reinterpret_cast<WarpStage *>(gl_smem_raw);
  WarpStage &S = stage[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long) blockDim.x + threadIdx.x) >> 5;
  const long long t = warp * 32 + lane;
  const bool live = t < nt;
  const int p = live ? tgt[t] : 0, pin = live ? perm[p] : 0;
  const HydroTarget T = hydro_load_target(live, p, pin, s_pm, s_vh, h_q, h_r, pressure, dhsml, divvel, curlvel);
  const unsigned live_mask = __ballot_sync(FULL, live);
  HydroAcc A;
  A.zero();
  unsigned visits = 0, staged = 0;
#ifdef GL_PHASE_PROF
  if(lane == 0) for(int q = 0; q < 8; q++) S.ph[q] = 0;
  __syncwarp();
#endif
  PH_T(tk0);
  const GroupF G = make_group(live, T.px, T.py, T.pz, T.h_i);
  const double hinv_i = 1.0 / T.h_i;
  const HydroConst K = { T.h_i * T.h_i, hinv_i, hinv_i * hinv_i * hinv_i * hinv_i, par.artvisc, par.timebase_interval };
  auto heavy = [&](int j) { hydro_pair(T, K, A, j, s_pm, s_vh, h_q, h_r); };
  int qh = 0, qt = 0;
  // hydro has no iteration: every warp but the last one is full, so lane = target always
  sph_collect<true>(S, lane, nentries, e_next, e_B, e_hmax, e_rng, n_child, s_pm, s_vh, s_type, live_mask, G, visits, staged, counter,
		    [&](int ncand, float lmax) { lt_chunk<true>(S, ncand, lmax, lane, G, qh, qt, heavy); });
  q_drain(S, lane, qh, qt, 0, heavy);
#ifdef GL_PHASE_PROF
  PH_ADD(S, 5, tk0);
  __syncwarp();
  if(lane < 6) atomicAdd(&g_phase[lane], S.ph[lane]);
#endif
  if(live) hydro_store(T, A, pin, o_hacc, o_dtentr, o_maxsig, o_pairs);
  const unsigned wpairs = __reduce_add_sync(FULL, live ? (unsigned) A.npairs : 0u);
  if(lane == 0 && wpairs) atomicAdd(&counter[10], (unsigned long long) wpairs);
}

int gls_hydro_begin(gl_ctx *c)
{
  if(!c->tree_valid)
    if(glt_build(c)) return 1;
  if(c->par.ComovingIntegrationOn) return c->fail("hydro_force: ComovingIntegrationOn != 0 is not implemented");
  if(c->o_density.cap < (size_t) c->n) return c->fail("hydro_force: density() has not been run (SphP.Density unknown)");
  if(gl_prepare_targets(c)) return 1;
  long long n = c->n;
  GL_CUDA(c, c->o_hacc.ensure_keep(3 * n));
  GL_CUDA(c, c->o_dtentr.ensure_keep(n));
  GL_CUDA(c, c->o_maxsig.ensure_keep(n));
  GL_CUDA(c, c->o_hydropairs.ensure_keep(n));
  GL_CUDA(c, c->h_q.ensure(c->nit + 1));
  GL_CUDA(c, c->h_r.ensure(c->nit + 1));
  GL_CUDA(c, c->d_counter.ensure(16));
  GL_CUDA(c, cudaMemsetAsync(c->d_counter.p + 10, 0, sizeof(unsigned long long), c->stream));
  GL_CUDA(c, cudaMemsetAsync(c->d_counter.p + 2, 0, sizeof(unsigned long long), c->stream));
  cudaEventRecord(c->ev[0], c->stream);
  k_hydro_prep<<<gl_grid(c->nit, 256), 256, 0, c->stream>>>(c->nit, c->perm.p, c->s_type.p, c->s_vh.p, c->o_density.p, c->o_pressure.p,
							c->o_dhsml.p, c->o_divvel.p, c->o_curlvel.p, c->in_timebin.p, c->h_q.p,
							c->h_r.p);
  GL_LAUNCH_COUNT(c);
  return 0;
}

// one launch over a list of gas targets (tree indices, Peano-Hilbert ordered)
int gls_hydro_launch(gl_ctx *c, const int *tgt, long long nt)
{
  if(nt <= 0) return 0;
  if(!c->attr_hydro)
    {
      GL_CUDA(c, cudaFuncSetAttribute(k_hydro, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (SPH_WARPS * sizeof(WarpStage))));
      c->attr_hydro = true;
    }
  gl_dparams dp = gl_make_dparams(c);
  k_hydro<<<gl_grid(nt, SPH_BLOCK), SPH_BLOCK, SPH_WARPS * sizeof(WarpStage), c->stream>>>(tgt, nt, (int) c->nentries, c->e_next.p, c->e_B.p,
							  c->e_hmax.p, c->e_rng.p, c->n_child.p, c->s_pm.p, c->s_vh.p, c->s_type.p,
							  c->h_q.p, c->h_r.p, c->perm.p, c->o_pressure.p, c->o_dhsml.p,
							  c->o_divvel.p, c->o_curlvel.p, c->o_hacc.p, c->o_dtentr.p,
							  c->o_maxsig.p, c->o_hydropairs.p, c->d_counter.p, dp);
  GL_LAUNCH_COUNT(c);
  return 0;
}

int gls_hydro_end(gl_ctx *c)
{
  cudaEventRecord(c->ev[1], c->stream);
  c->fire_post_launch();
  unsigned long long hp = 0, viol = 0;
  GL_CUDA(c, cudaMemcpyAsync(&hp, c->d_counter.p + 10, sizeof(hp), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaMemcpyAsync(&viol, c->d_counter.p + 2, sizeof(viol), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, c->wait_stream());
  GL_CUDA(c, cudaGetLastError());
  if(viol == GL_ERR_LET_VIOLATION) return c->fail("hydro_force: GL_ERR_LET_VIOLATION");
  float ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->st.ms_hydro = ms;
  c->st.hydro_pairs = (long long) hp;
#ifdef GL_PHASE_PROF
  {
    unsigned long long ph[8];
    cudaMemcpyFromSymbol(ph, g_phase, sizeof(ph));
    fprintf(stderr, "[gl] hydro phase cycles (sum over warps, cumulative): collect %.3g stage %.3g test %.3g heavy %.3g reduce %.3g total %.3g\n",
	    (double) ph[0], (double) ph[1], (double) ph[2], (double) ph[3], (double) ph[4], (double) ph[5]);
  }
#endif
  c->st.hydro_targets = c->n_tgt_gas;
  return 0;
}

int gls_hydro(gl_ctx *c)
{
  if(gls_hydro_begin(c)) return 1;
  if(gls_hydro_launch(c, c->tgt_gas.p, c->n_tgt_gas)) return 1;
  return gls_hydro_end(c);
}
