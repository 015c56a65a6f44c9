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

struct GravTarget
{
  double x, y, z, aold, soft;
};

#ifndef GRAV_MIN_BLOCKS
#define GRAV_MIN_BLOCKS 12	// 40 registers, 48 warps per SM: 3.15 ms for the 1.6M disc (10 blocks: 3.24, 16 blocks with spills: 3.39)
#endif
template <bool BH, bool POT>
__global__ void __launch_bounds__(GL_BLOCK, GRAV_MIN_BLOCKS)
k_gravity_walk(const int *__restrict__ tgt, long long nt, int nentries, const int *__restrict__ e_next,
	       const double4 *__restrict__ e_A, const double4 *__restrict__ e_B, const double2 *__restrict__ e_C,
	       const double4 *__restrict__ s_pm, const double4 *__restrict__ s_vh, const int *__restrict__ s_type,
	       const int *__restrict__ perm, const double *__restrict__ in_oldacc, double *__restrict__ o_acc,
	       double *__restrict__ o_oldacc, float *__restrict__ o_gravcost, double *__restrict__ o_pot,
	       unsigned long long *__restrict__ counter, gl_dparams par, int last_first)
{
  const int lane = threadIdx.x & 31;
  long long warp = (blockIdx.x * (long long) blockDim.x + threadIdx.x) >> 5;
  // A caller-ordered list (a Peano-Hilbert segment) ends with the few non-gas particles (the gas block comes
  // first, peano.c:37-88): the warp that mixes them with far-away gas particles marches through two unrelated
  // parts of the tree and takes about twice as long as the others.  Started last it was a 0.7 ms tail of the
  // 1.6M disc's walk; started first it disappears behind the rest.
  if(last_first) warp = ((nt + 31) >> 5) - 1 - warp;
  const long long t = warp * 32 + lane;
  const bool live = t >= 0 && t < nt;
  int p = 0, ptype = 0, pin = 0;
  double px = 0, py = 0, pz = 0, aold = 0, hi = 0;
  if(live)
    {
      p = tgt[t];
      double4 pm = s_pm[p];
      px = pm.x; py = pm.y; pz = pm.z;
      ptype = s_type[p];
      pin = perm[p];
      aold = __dmul_rn(par.errtolforceacc, in_oldacc[pin]);	// forcetree.c:1850
      hi = (ptype == 0) ? s_vh[p].w : par.soft[ptype];	// forcetree.c:2039-2042 / :2117-2120
    }
  const double theta2f = par.theta;
  double ax = 0, ay = 0, az = 0;
  int ninter = 0;
  int skip_until = live ? 0 : nentries;
  int e = 0;
  while(e < nentries)
    {
      const int nx = __ldg(&e_next[e]);
      const double4 A = gl_ld4(&e_A[e]);
      const bool active = e >= skip_until;
      double dx = __dsub_rn(A.x, px), dy = __dsub_rn(A.y, py), dz = __dsub_rn(A.z, pz);
      double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      double h = hi;
      bool interact = false;
      if(nx < 0)
	{
	  // particle leaf (forcetree.c:1906-1918, :2035-2070)
	  const double hj = __ldg(&e_B[e].x);
	  if(h < hj) h = hj;
	  interact = active;
	  e = e + 1;
	}
      else
	{
	  const double4 B = gl_ld4(&e_B[e]);
	  const double2 C = __ldg(&e_C[e]);
	  if(active)
	    {
	      bool open;
	      if(BH)
		{
		  // forcetree.c:2076  len*len > r2*theta*theta
		  open = __dmul_rn(B.w, B.w) > __dmul_rn(__dmul_rn(r2, theta2f), theta2f);
		}
	      else
		{
		  // forcetree.c:2085  mass*len*len > r2*r2*aold
		  open = C.y > __dmul_rn(__dmul_rn(r2, r2), aold);
		  if(!open)
		    {
		      // forcetree.c:2094-2105  inside the 0.6*len box around the geometric centre
		      const double lim = __dmul_rn(0.60, B.w);
		      if(fabs(__dsub_rn(B.x, px)) < lim && fabs(__dsub_rn(B.y, py)) < lim && fabs(__dsub_rn(B.z, pz)) < lim)
			open = true;
		    }
		}
	      if(!open && h < C.x)
		{
		  // forcetree.c:2124-2137  softening guard
		  h = C.x;
		  if(r2 < __dmul_rn(h, h)) open = true;
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
	  const float r2f = (float) r2, mf = (float) A.w, hf = (float) h;
	  // one MUFU.RSQ gives r and 1/r^3 (2 ulp, FP32): the IEEE sqrtf + divide it replaces cost ~4x the instructions.
	  // r2f == 0 is the target's own leaf: rinv = inf, r = NaN -> the comparison is false -> spline branch with r = 0
	  const float rinv = rsqrtf(r2f);
	  const float r = (r2f > 0.f) ? r2f * rinv : 0.f;
	  float fac;
	  if(POT)
	    {
	      // forcetree.c:3497-3511: -m/r outside the softening length, spline potential wp inside
	      if(r >= hf)
		fac = -mf * rinv;
	      else
		{
		  const float hinv = 1.0f / hf;
		  const float u = r * hinv;
		  float wp;
		  if(u < 0.5f)
		    wp = -2.8f + u * u * (5.333333333333f + u * u * (6.4f * u - 9.6f));
		  else
		    wp = -3.2f + 0.066666666667f / u + u * u * (10.666666666667f + u * (-16.0f + u * (9.6f - 2.133333333333f * u)));
		  fac = mf * hinv * wp;
		}
	      ax += (double) fac;
	    }
	  else if(r >= hf)
	    fac = mf * rinv * rinv * rinv;
	  else
	    {
	      const float hinv = 1.0f / hf;
	      const float h3inv = hinv * hinv * hinv;
	      const float u = r * hinv;
	      if(u < 0.5f)
		fac = mf * h3inv * (10.666666666667f + u * u * (32.0f * u - 38.4f));
	      else
		fac = mf * h3inv * (21.333333333333f - 48.0f * u + 38.4f * u * u - 10.666666666667f * u * u * u -
				    0.066666666667f / (u * u * u));
	    }
	  if(!POT)
	    {
	      const double fd = (double) fac;
	      ax = fma(dx, fd, ax);
	      ay = fma(dy, fd, ay);
	      az = fma(dz, fd, az);
	    }
	  if(A.w > 0) ninter++;	// forcetree.c:2214
	}
    }
  if(live && POT)
    {
      // potential.c:169-184: remove the self term with All.SofteningTable, then x G (the Lambda term of :235 is 0 here)
      const double m_i = s_pm[p].w;
      o_pot[pin] = __dmul_rn(__dadd_rn(ax, __ddiv_rn(m_i, par.softtab[ptype])), par.G);
    }
  if(live && !POT)
    {
      // gravtree.c:388-403: OldAcc from the un-scaled acceleration, then x G (+ the Lambda term, :405-411)
      const double old = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az)));
      o_oldacc[pin] = old;
      o_acc[3ll * pin + 0] = __dadd_rn(__dmul_rn(ax, par.G), __dmul_rn(par.lambda_fac, px));
      o_acc[3ll * pin + 1] = __dadd_rn(__dmul_rn(ay, par.G), __dmul_rn(par.lambda_fac, py));
      o_acc[3ll * pin + 2] = __dadd_rn(__dmul_rn(az, par.G), __dmul_rn(par.lambda_fac, pz));
      o_gravcost[pin] = (float) ninter;
    }
  // interaction total (timings.txt "ia/part", gravtree.c:822)
  unsigned long long tot = (unsigned long long) __reduce_add_sync(0xffffffffu, live ? ninter : 0);
  if(lane == 0 && tot && !POT) atomicAdd(&counter[0], tot);
}

int glg_gravity(gl_ctx *c)
{
  if(!c->have_par) return c->fail("gravity_tree: gl_set_params not called");
  if(c->n <= 0) return c->fail("gravity_tree: no particles loaded");
  if(c->tree_flag || !c->tree_valid)	// gravtree.c:60-70: rebuild when TreeReconstructFlag is set
    if(glt_build(c)) return 1;
  if(gl_prepare_targets(c)) return 1;
  long long n = c->n, nt = c->n_tgt;
  GL_CUDA(c, c->o_acc.ensure(3 * n));
  GL_CUDA(c, c->o_oldacc.ensure(n));
  GL_CUDA(c, c->o_gravcost.ensure(n));
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
	k_gravity_walk<true, false><<<grid, GL_BLOCK, 0, c->stream>>>(c->tgt_sorted.p, nt, (int) c->nentries, c->e_next.p, c->e_A.p,
							      c->e_B.p, c->e_C.p, c->s_pm.p, c->s_vh.p, c->s_type.p,
							      c->perm.p, c->in_oldacc.p, c->o_acc.p, c->o_oldacc.p,
							      c->o_gravcost.p, nullptr, c->d_counter.p, dp, lf);
      else
	k_gravity_walk<false, false><<<grid, GL_BLOCK, 0, c->stream>>>(c->tgt_sorted.p, nt, (int) c->nentries, c->e_next.p, c->e_A.p,
							       c->e_B.p, c->e_C.p, c->s_pm.p, c->s_vh.p, c->s_type.p,
							       c->perm.p, c->in_oldacc.p, c->o_acc.p, c->o_oldacc.p,
							       c->o_gravcost.p, nullptr, c->d_counter.p, dp, lf);
      GL_LAUNCH_COUNT(c);
    }
  cudaEventRecord(c->ev[1], c->stream);
  c->fire_post_launch();
  unsigned long long tot = 0;
  GL_CUDA(c, cudaMemcpyAsync(&tot, c->d_counter.p, sizeof(tot), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, c->wait_stream());
  GL_CUDA(c, cudaGetLastError());
  float ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->st.ms_gravity_walk = ms;
  c->st.interactions = (long long) tot;
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
    k_gravity_walk<true, true><<<grid, GL_BLOCK, 0, c->stream>>>(c->tgt_sorted.p, nt, (int) c->nentries, c->e_next.p, c->e_A.p, c->e_B.p,
								c->e_C.p, c->s_pm.p, c->s_vh.p, c->s_type.p, c->perm.p, c->in_oldacc.p,
								nullptr, nullptr, nullptr, c->o_pot.p, c->d_counter.p, dp, 0);
  else
    k_gravity_walk<false, true><<<grid, GL_BLOCK, 0, c->stream>>>(c->tgt_sorted.p, nt, (int) c->nentries, c->e_next.p, c->e_A.p, c->e_B.p,
								 c->e_C.p, c->s_pm.p, c->s_vh.p, c->s_type.p, c->perm.p, c->in_oldacc.p,
								 nullptr, nullptr, nullptr, c->o_pot.p, c->d_counter.p, dp, 0);
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
