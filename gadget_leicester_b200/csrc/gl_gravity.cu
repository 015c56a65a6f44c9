// gl_gravity.cu - K6 warp-cooperative tree walk + K11 epilogue.
//
// Reference: force_treeevaluate (forcetree.c:1797-2317) walks, for ONE target, the threaded tree
// (nextnode = open, sibling = skip) and applies a per-target opening test; gravity_tree
// (gravtree.c:381-403) then sets OldAcc = |acc| and multiplies by G.
//
// B200 design: one warp owns 32 consecutive targets of the Morton-sorted order (neighbours in
// space).  The warp marches ONCE through the DFS entry array; the entry record is fetched with a
// warp-uniform vector load (one L1/L2 transaction for 32 targets) and every lane applies ITS OWN
// opening test - the decisions, and therefore the interaction counts, are those of the reference,
// lane by lane.  A lane that accepted a node idles until the march leaves that subtree
// (skip_until); when all lanes have accepted, the warp jumps.  All accept/open predicates are
// evaluated in FP64 with the reference's operation order and without FMA contraction, so they are
// bit-identical to the gcc -ffp-contract=off oracle; only the force magnitude m/r^3 (or the spline)
// is evaluated in FP32, and accumulated in FP64.
#include "gl_internal.h"

gl_dparams gl_make_dparams(const gl_ctx *c);

#ifndef GRAV_MIN_BLOCKS
#define GRAV_MIN_BLOCKS 10	// resident 128-thread blocks per SM the walk is compiled for (48 registers; 8: 3.37 ms, 10: 3.18, 12: 3.18 on the 1.6M disc)
#endif

// ---- margins of the FP32 pre-test (units: domain cube [1,2) wide, total mass [1,2); u = 2^-24) --------------------
// A coordinate difference d_f = (hi_s - hi_t) + (lo_s - lo_t) obeys |d_f - d| <= u|d| + 2^-43 (gl_split), hence
// |r2_f - r2| <= 2^-20.8 r2 + 2^-62: the interval [r2_f - err, r2_f + err] with err = 2^-20 r2_f + 2^-58 holds the
// reference's FP64 r2.  Every product below is rounded a few more times (<= 13 u in total); the factors (1 +- 2^-20)
// on the node side of each comparison cover that.  A lane whose interval straddles a threshold repeats the decision
// with the reference's own FP64 statements (node_open_exact): decisions - and interaction counts - are the reference's.
#define GW_REL 9.5367431640625e-07f	// 2^-20
#define GW_ABS 3.4694469519536142e-18f	// 2^-58
#define GW_ONE_LO (1.0f - GW_REL)
#define GW_ONE_HI (1.0f + GW_REL)
#define GW_TWO_LO (1.0f - 2.0f * GW_REL)
#define GW_TWO_HI (1.0f + 2.0f * GW_REL)

// The reference's opening decision for node entry e and ONE target, in FP64, statement by statement
// (forcetree.c:2009-2011, :2074-2137).  Returns -1 = open, else the softening length to use for the accepted node.
template <bool BH>
__device__ __noinline__ double node_open_exact(int e, const double4 *__restrict__ e_A, const double4 *__restrict__ e_B,
					     const double2 *__restrict__ e_C, double px, double py, double pz, double aold,
					     double theta, double hi)
{
  const double4 A = gl_ld4(&e_A[e]);
  const double4 B = gl_ld4(&e_B[e]);
  const double2 C = __ldg(&e_C[e]);
  const double dx = __dsub_rn(A.x, px), dy = __dsub_rn(A.y, py), dz = __dsub_rn(A.z, pz);
  const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  double h = hi;
  bool open;
  if(BH)
    open = __dmul_rn(B.w, B.w) > __dmul_rn(__dmul_rn(r2, theta), theta);	// forcetree.c:2076
  else
    {
      open = C.y > __dmul_rn(__dmul_rn(r2, r2), aold);	// forcetree.c:2085
      if(!open)
	{
	  const double lim = __dmul_rn(0.60, B.w);	// forcetree.c:2094-2105
	  if(fabs(__dsub_rn(B.x, px)) < lim && fabs(__dsub_rn(B.y, py)) < lim && fabs(__dsub_rn(B.z, pz)) < lim) open = true;
	}
    }
  if(!open && h < C.x)
    {				// forcetree.c:2124-2137
      h = C.x;
      if(r2 < __dmul_rn(h, h)) open = true;
    }
  return open ? -1.0 : h;
}

// K6.  One warp = 32 consecutive targets, marching once through the DFS entry array (warp-uniform record loads).
// Per entry every lane takes ITS OWN opening decision - in FP32 on the split coordinates, inside error bounds that
// make the decision certain, and in FP64 (node_open_exact) in the rare uncertain shell.  Force magnitude in FP32,
// FP64 accumulation.
template <bool BH, bool POT, int MINB>
__global__ void __launch_bounds__(GL_BLOCK, MINB)
k_gravity_walk(const int *__restrict__ tgt, long long nt, int nentries, const int *__restrict__ e_next,
	       const float4 *__restrict__ e_F0, const float4 *__restrict__ e_F1, const float4 *__restrict__ e_F2,
	       const double4 *__restrict__ e_A, const double4 *__restrict__ e_B, const double2 *__restrict__ e_C,
	       const double4 *__restrict__ s_pm, const double4 *__restrict__ s_vh, const int *__restrict__ s_type,
	       const int *__restrict__ perm, const double *__restrict__ in_oldacc, const double *__restrict__ scale,
	       double *__restrict__ o_acc, double *__restrict__ o_oldacc, float *__restrict__ o_gravcost,
	       double *__restrict__ o_pot, unsigned long long *__restrict__ counter, gl_dparams par, gl_domain dom, int last_first)
{
  const int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long) blockDim.x + threadIdx.x) >> 5;
  // A caller-ordered list (a Peano-Hilbert segment) ends with the few non-gas particles (the gas block comes
  // first, peano.c:37-88): the warp that mixes them with far-away gas particles marches through two unrelated
  // parts of the tree and takes about twice as long as the others.  Started first it disappears behind the rest.
  if(last_first) warp = ((nt + 31) >> 5) - 1 - warp;
  const long long t = warp * 32 + lane;
  const bool live = t >= 0 && t < nt;
  // (p, its type, its input index and the two scales are only needed in the rare FP64 branch and in the epilogue: they are
  //  fetched again there instead of living in registers through the march - at 48 registers the compiler otherwise spills
  //  four of the six split coordinates and re-loads them from local memory in EVERY iteration)
  float thx = 0, thy = 0, thz = 0, tlx = 0, tly = 0, tlz = 0;	// split target position
  float hif = 0, af = 0;
  if(live)
    {
      const double S = scale[0], Ms = scale[1];
      const int p = tgt[t];
      const double4 pm = s_pm[p];
      const int ptype = s_type[p];
      const int pin = perm[p];
      gl_split(pm.x, dom.center[0], S, thx, tlx);
      gl_split(pm.y, dom.center[1], S, thy, tly);
      gl_split(pm.z, dom.center[2], S, thz, tlz);
      const double hi = (ptype == 0) ? s_vh[p].w : par.soft[ptype];	// forcetree.c:2039-2042 / :2117-2120
      hif = (float) (hi * S);
      // aold = ErrTolForceAcc * OldAcc (forcetree.c:1850), in the scaled units: x Ms / S^2
      const double aold = __dmul_rn(par.errtolforceacc, in_oldacc[pin]) * (Ms / (S * S));
      af = (float) aold;
    }
  const float thetaf = (float) par.theta;
  double ax = 0, ay = 0, az = 0;
  int ninter = 0;
  int skip_until = live ? 0 : nentries;
  int e = 0;
  while(e < nentries)
    {
      const int nx = __ldg(&e_next[e]);
      const float4 F0 = __ldg(&e_F0[e]);
      const float4 F1 = __ldg(&e_F1[e]);
      const bool active = e >= skip_until;
      const float dx = (F0.x - thx) + (F1.x - tlx), dy = (F0.y - thy) + (F1.y - tly), dz = (F0.z - thz) + (F1.z - tlz);
      const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
      float hf = hif;
      bool interact = false;
      if(nx < 0)
	{
	  // particle leaf (forcetree.c:1906-1918, :2035-2070): no decision to take
	  hf = fmaxf(hf, F1.w);
	  interact = active;
	  e = e + 1;
	}
      else
	{
	  const float4 F2 = __ldg(&e_F2[e]);
	  if(active)
	    {
	      // all sub-tests are evaluated unconditionally (branch-free): with ~14 of 32 lanes active nearly every path
	      // is taken by some lane anyway, and the predicated form drops the divergence bookkeeping
	      const float err = fmaf(r2, GW_REL, GW_ABS);
	      const float r2lo = fmaxf(r2 - err, 0.f), r2hi = r2 + err;
	      float lhs, t_hi, t_lo;
	      if(BH)
		{
		  // forcetree.c:2076  len*len > r2*theta*theta
		  lhs = F1.w * F1.w;
		  t_hi = (r2hi * thetaf) * thetaf;
		  t_lo = (r2lo * thetaf) * thetaf;
		}
	      else
		{
		  // forcetree.c:2085  mass*len*len > r2*r2*aold
		  lhs = (F0.w * F1.w) * F1.w;
		  t_hi = (r2hi * r2hi) * af;
		  t_lo = (r2lo * r2lo) * af;
		}
	      // both margin factors sit on the node side: (1 - 2^-20) / (1 + 2^-20) > 1 - 2^-19 (one register for aold, not two)
	      const bool mac_open = lhs * GW_TWO_LO > t_hi, mac_closed = lhs * GW_TWO_HI < t_lo;
	      // forcetree.c:2094-2105  inside the 0.6*len box around the geometric centre (relative criterion only)
	      const float cmax = fmaxf(fmaxf(fabsf(dx + F2.x), fabsf(dy + F2.y)), fabsf(dz + F2.z));
	      const bool box_in = !BH && cmax < F1.w * (0.6f - GW_REL), box_out = BH || cmax > F1.w * (0.6f + GW_REL);
	      // forcetree.c:2124-2137  softening guard: h_i < maxsoft -> h = maxsoft, opened if r2 < h*h
	      // (monotone rounding: hif < F2.w implies h_i < maxsoft in FP64 as well; equal floats leave the order open,
	      //  which only matters when r2 < maxsoft^2 is possible)
	      const float ms2 = F2.w * F2.w;
	      const bool g_lt = hif < F2.w, g_gt = hif > F2.w;
	      const bool g_open = r2hi < ms2 * GW_ONE_LO, g_closed = r2lo > ms2 * GW_ONE_HI;
	      const bool closedpath = mac_closed && box_out;
	      bool open = mac_open || (mac_closed && box_in) || (closedpath && g_lt && g_open);
	      const bool sure = open || (closedpath && (g_gt || g_closed));
	      hf = g_lt ? F2.w : hif;
	      if(!sure)
		{
		  // the uncertain shell: the reference's own FP64 statements
		  const int p = tgt[t];
		  const int ptype = s_type[p], pin = perm[p];
		  const double S = scale[0];
		  const double4 pm = s_pm[p];
		  const double hi = (ptype == 0) ? s_vh[p].w : par.soft[ptype];
		  const double aold = __dmul_rn(par.errtolforceacc, in_oldacc[pin]);
		  const double hn = node_open_exact<BH>(e, e_A, e_B, e_C, pm.x, pm.y, pm.z, aold, par.theta, hi);
		  open = hn < 0;
		  hf = (float) (hn * S);
		}
	      if(open && nx == e + 1)
		{
		  // a CUT CELL of the locally-essential tree (multi-rank mode): its descendants stayed with their owner,
		  // whose export test promised that no target of this rank opens it.  Report instead of guessing.
		  atomicMax(&counter[2], GL_ERR_LET_VIOLATION);
		  open = false;
		}
	      if(!open)
		{
		  interact = true;
		  skip_until = nx;
		}
	    }
	  const int m = __reduce_min_sync(0xffffffffu, skip_until);
	  e = max(e + 1, m);
	}
      if(interact)
	{
	  // forcetree.c:2143-2198, magnitude in FP32
	  const float mf = F0.w;
	  // one MUFU.RSQ gives r and 1/r^3.  r2 == 0 is the target's own leaf: rinv = inf, r = 0 -> spline branch with u = 0
	  const float rinv = rsqrtf(r2);
	  const float r = (r2 > 0.f) ? r2 * rinv : 0.f;
	  float fac;
	  if(POT)
	    {
	      // forcetree.c:3497-3511: -m/r outside the softening length, spline potential wp inside
	      if(r >= hf)
		fac = -mf * rinv;
	      else
		{
		  const float hinv = __fdividef(1.0f, hf);	// MUFU.RCP (2 ulp): the magnitude is FP32 anyway
		  const float u = r * hinv;
		  float wp;
		  if(u < 0.5f)
		    wp = -2.8f + u * u * (5.333333333333f + u * u * (6.4f * u - 9.6f));
		  else
		    wp = -3.2f + 0.066666666667f * __fdividef(1.0f, u) + u * u * (10.666666666667f + u * (-16.0f + u * (9.6f - 2.133333333333f * u)));
		  fac = mf * hinv * wp;
		}
	      ax += (double) fac;
	    }
	  else if(r >= hf)
	    fac = mf * rinv * rinv * rinv;
	  else
	    {
	      const float hinv = __fdividef(1.0f, hf);	// MUFU.RCP (2 ulp): the magnitude is FP32 anyway
	      const float h3inv = hinv * hinv * hinv;
	      const float u = r * hinv;
	      if(u < 0.5f)
		fac = mf * h3inv * (10.666666666667f + u * u * (32.0f * u - 38.4f));
	      else
		{
		  const float uinv = __fdividef(1.0f, u);
		  fac = mf * h3inv * (21.333333333333f - 48.0f * u + 38.4f * u * u - 10.666666666667f * u * u * u -
				      0.066666666667f * (uinv * uinv * uinv));
		}
	    }
	  if(!POT)
	    {
	      const double fd = (double) fac;
	      ax = fma((double) dx, fd, ax);
	      ay = fma((double) dy, fd, ay);
	      az = fma((double) dz, fd, az);
	    }
	  if(mf > 0.f) ninter++;	// forcetree.c:2214
	}
    }
  if(live && POT)
    {
      // potential.c:169-184: remove the self term with All.SofteningTable, then x G (the Lambda term of :235 is 0 here)
      const double S = scale[0], Ms = scale[1];
      const int p = tgt[t];
      const int ptype = s_type[p], pin = perm[p];
      const double m_i = s_pm[p].w;
      ax *= S / Ms;		// back from the scaled units: m'/r' = m/r * Ms / S
      o_pot[pin] = __dmul_rn(__dadd_rn(ax, __ddiv_rn(m_i, par.softtab[ptype])), par.G);
    }
  if(live && !POT)
    {
      // gravtree.c:388-403: OldAcc from the un-scaled acceleration, then x G (+ the Lambda term, :405-411)
      const double S = scale[0], Ms = scale[1];
      const int p = tgt[t];
      const int pin = perm[p];
      const double un = (S * S) / Ms;	// d' m'/r'^3 = d m/r^3 * Ms / S^2
      ax *= un; ay *= un; az *= un;
      const double4 pm = s_pm[p];
      const double old = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az)));
      o_oldacc[pin] = old;
      o_acc[3ll * pin + 0] = __dadd_rn(__dmul_rn(ax, par.G), __dmul_rn(par.lambda_fac, pm.x));
      o_acc[3ll * pin + 1] = __dadd_rn(__dmul_rn(ay, par.G), __dmul_rn(par.lambda_fac, pm.y));
      o_acc[3ll * pin + 2] = __dadd_rn(__dmul_rn(az, par.G), __dmul_rn(par.lambda_fac, pm.z));
      o_gravcost[pin] = (float) ninter;
    }
  // interaction total (timings.txt "ia/part", gravtree.c:822)
  unsigned long long tot = (unsigned long long) __reduce_add_sync(0xffffffffu, live ? ninter : 0);
  if(lane == 0 && tot && !POT) atomicAdd(&counter[0], tot);
}

#define GRAV_ARGS                                                                                                          \
  c->tgt_sorted.p, nt, (int) c->nentries, c->e_next.p, c->e_F0.p, c->e_F1.p, c->e_F2.p, c->e_A.p, c->e_B.p, c->e_C.p, c->s_pm.p, \
    c->s_vh.p, c->s_type.p, c->perm.p, c->in_oldacc.p, c->grav_scale.p

int glg_gravity(gl_ctx *c)
{
  if(!c->have_par) return c->fail("gravity_tree: gl_set_params not called");
  if(c->n <= 0) return c->fail("gravity_tree: no particles loaded");
  if(c->tree_flag || !c->tree_valid)	// gravtree.c:60-70: rebuild when TreeReconstructFlag is set
    if(glt_build(c)) return 1;
  if(gl_prepare_targets(c)) return 1;
  long long n = c->n, nt = c->n_tgt;
  // (ensure_keep: with an active SUBSET the results of the other particles are state - the resident loop integrates
  //  them - and must survive a build that imported more particles than the last one)
  GL_CUDA(c, c->o_acc.ensure_keep(3 * n));
  GL_CUDA(c, c->o_oldacc.ensure_keep(n));
  GL_CUDA(c, c->o_gravcost.ensure_keep(n));
  GL_CUDA(c, c->d_counter.ensure(8));
  GL_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, 8 * sizeof(unsigned long long), c->stream));
  // non-targets keep their OldAcc (only the active list is updated, gravtree.c:381-392)
  GL_CUDA(c, cudaMemcpyAsync(c->o_oldacc.p, c->in_oldacc.p, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  gl_dparams dp = gl_make_dparams(c);
  cudaEventRecord(c->ev[0], c->stream);
  if(nt > 0)
    {
      int grid = gl_grid(nt, GL_BLOCK);
      const int lf = (c->n_tgt_in >= 0 && c->tgt_ordered) ? 1 : 0;
      if(c->par.ErrTolTheta != 0)
	k_gravity_walk<true, false, GRAV_MIN_BLOCKS><<<grid, GL_BLOCK, 0, c->stream>>>(GRAV_ARGS, c->o_acc.p, c->o_oldacc.p, c->o_gravcost.p, nullptr,
							      c->d_counter.p, dp, c->dom, lf);
      else
	{
	  // tuning switch (GL_GRAV_MINB = 12 / 10 / 8 resident blocks per SM = 40 / 48 / 64 registers); default GRAV_MIN_BLOCKS
	  static const int minb = getenv("GL_GRAV_MINB") ? atoi(getenv("GL_GRAV_MINB")) : GRAV_MIN_BLOCKS;
	  if(minb == 12)
	    k_gravity_walk<false, false, 12><<<grid, GL_BLOCK, 0, c->stream>>>(GRAV_ARGS, c->o_acc.p, c->o_oldacc.p, c->o_gravcost.p, nullptr,
									       c->d_counter.p, dp, c->dom, lf);
	  else if(minb == 10)
	    k_gravity_walk<false, false, 10><<<grid, GL_BLOCK, 0, c->stream>>>(GRAV_ARGS, c->o_acc.p, c->o_oldacc.p, c->o_gravcost.p, nullptr,
									       c->d_counter.p, dp, c->dom, lf);
	  else
	    k_gravity_walk<false, false, 8><<<grid, GL_BLOCK, 0, c->stream>>>(GRAV_ARGS, c->o_acc.p, c->o_oldacc.p, c->o_gravcost.p, nullptr,
									      c->d_counter.p, dp, c->dom, lf);
	}
      GL_LAUNCH_COUNT(c);
    }
  cudaEventRecord(c->ev[1], c->stream);
  c->fire_post_launch();
  unsigned long long cnt3[3] = { 0, 0, 0 };
  GL_CUDA(c, cudaMemcpyAsync(cnt3, c->d_counter.p, sizeof(cnt3), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, c->wait_stream());
  GL_CUDA(c, cudaGetLastError());
  const unsigned long long tot = cnt3[0];
  float ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->st.ms_gravity_walk = ms;
  c->st.interactions = (long long) tot;
  if(cnt3[2] == GL_ERR_LET_VIOLATION) return c->fail("gravity_tree: GL_ERR_LET_VIOLATION");
  // gravtree.c:392: the new OldAcc feeds the next call's relative criterion (accel.c:61-67 calls twice at Ti=0)
  GL_CUDA(c, cudaMemcpyAsync(c->in_oldacc.p, c->o_oldacc.p, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  // gravtree.c:396-397: switch to the relative criterion for all following force computations
  if(c->par.TypeOfOpeningCriterion == 1) c->par.ErrTolTheta = 0;
  return 0;
}

// compute_potential (potential.c:22): the same walk and opening decisions, accumulating -m/r (or the spline
// potential) for ALL particles; the self term is removed with All.SofteningTable as potential.c:176 does.
int glg_potential(gl_ctx *c)
{
  if(!c->have_par) return c->fail("compute_potential: gl_set_params not called");
  if(c->n <= 0) return c->fail("compute_potential: no particles loaded");
  if(c->tree_flag || !c->tree_valid)	// potential.c:40-52
    if(glt_build(c)) return 1;
  // potential.c:54 loops over all particles, not the active list
  const long long keep = c->n_tgt_in;
  c->n_tgt_in = -1;
  c->tgt_dirty = true;
  if(gl_prepare_targets(c)) return 1;
  long long n = c->n, nt = c->n_tgt;
  GL_CUDA(c, c->o_pot.ensure(n));
  GL_CUDA(c, c->d_counter.ensure(16));
  gl_dparams dp = gl_make_dparams(c);
  cudaEventRecord(c->ev[0], c->stream);
  int grid = gl_grid(nt, GL_BLOCK);
  if(c->par.ErrTolTheta != 0)
    k_gravity_walk<true, true, GRAV_MIN_BLOCKS><<<grid, GL_BLOCK, 0, c->stream>>>(GRAV_ARGS, nullptr, nullptr, nullptr, c->o_pot.p, c->d_counter.p, dp, c->dom, 0);
  else
    k_gravity_walk<false, true, GRAV_MIN_BLOCKS><<<grid, GL_BLOCK, 0, c->stream>>>(GRAV_ARGS, nullptr, nullptr, nullptr, c->o_pot.p, c->d_counter.p, dp, c->dom, 0);
  GL_LAUNCH_COUNT(c);
  cudaEventRecord(c->ev[1], c->stream);
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  GL_CUDA(c, cudaGetLastError());
  float ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->st.ms_potential = ms;
  c->n_tgt_in = keep;
  c->tgt_dirty = true;
  return 0;
}
