// gl_sph.cu - K7/K8 SPH density with in-kernel smoothing-length iteration, K10 hydro force.
//
// Reference: density() (density.c:89-704) loops "evaluate all unconverged targets, post-process,
// adjust Hsml" until every target holds DesNumNgb +- MaxNumNgbDeviation kernel-weighted neighbours;
// density_evaluate (density.c:711-1029) gathers the neighbours with ngb_treefind_variable
// (ngb.c:169-297) and forms 11 sums; hydro_force/hydro_evaluate (hydra.c:145-813, :822-1995) do the
// same with the symmetric search ngb_treefind_pairs (ngb.c:32-160).
//
// B200 design: the Hsml sequence of a target depends on that target alone (the one-sided search
// only uses h_i), so the whole iteration runs inside one kernel: a warp owns 32 consecutive SPH
// targets of the Morton order, marches once per round through the DFS entry array pruning with the
// bounding box of its still-unconverged lanes, and every lane tests each surviving gas particle
// with the reference's FP64 predicate r2 < h_i^2.  Neighbours are met in the reference's own order
// (DFS = child-slot order), sums are FP64 without FMA contraction, so Density/NumNgb agree with
// the oracle to the last bits and the Hsml iteration follows the reference's path exactly.
// Rounds continue until all lanes of the warp are done - no host round trips, no relaunch.
#include "gl_internal.h"

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

__device__ __forceinline__ double warp_min_d(double v, unsigned mask = 0xffffffffu)
{
  for(int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(mask, v, o));
  return v;
}
__device__ __forceinline__ double warp_max_d(double v, unsigned mask = 0xffffffffu)
{
  for(int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o));
  return v;
}

struct DensOut
{
  double *hsml_in, *numngb, *density, *dhsml, *divvel, *curlvel, *pressure, *sinkrho, *sinkentr, *sinkvel;
  int *pairs;
};

__global__ void __launch_bounds__(GL_BLOCK)
k_density(const int *__restrict__ tgt, long long nt, int nentries, const int *__restrict__ e_next,
	  double4 *__restrict__ e_B, const double4 *__restrict__ s_pm, double4 *__restrict__ s_vh,
	  const int *__restrict__ s_type, const double *__restrict__ s_entropy, const int *__restrict__ noff,
	  const int *__restrict__ perm, const double *__restrict__ in_entropy, const double *__restrict__ in_dtentropy,
	  const int *__restrict__ in_timebin, const int *__restrict__ in_tibegstep, DensOut out,
	  unsigned long long *__restrict__ counter, gl_dparams par)
{
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long) blockDim.x + threadIdx.x) >> 5;
  const long long t = warp * 32 + lane;
  const bool live = t < nt;
  int p = 0, ptype = 0, pin = 0;
  double px = 0, py = 0, pz = 0, vx = 0, vy = 0, vz = 0, h = 1;
  if(live)
    {
      p = tgt[t];
      double4 pm = s_pm[p];
      double4 vh = s_vh[p];
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
  double left = 0, right = 0;
  bool todo = live;
  int iter = 0;
  double rho = 0, wngb = 0, dhsmlrho = 0, divv = 0, rot0 = 0, rot1 = 0, rot2 = 0, sentr = 0, gv0 = 0, gv1 = 0, gv2 = 0;
  int numngb = 0;

  while(__any_sync(0xffffffffu, todo))
    {
      // bounding box of the search spheres of the lanes still iterating
      const double hb = h * BOX_SLACK;
      const double bxlo = warp_min_d(todo ? px - hb : 1e300), bxhi = warp_max_d(todo ? px + hb : -1e300);
      const double bylo = warp_min_d(todo ? py - hb : 1e300), byhi = warp_max_d(todo ? py + hb : -1e300);
      const double bzlo = warp_min_d(todo ? pz - hb : 1e300), bzhi = warp_max_d(todo ? pz + hb : -1e300);
      const double h2 = h * h;
      const double hinv = 1.0 / h;
      const double hinv3 = hinv * hinv * hinv;
      const double hinv4 = hinv3 * hinv;
      if(todo)
	{			// lanes that already converged keep the sums of their final round
	  rho = wngb = dhsmlrho = divv = rot0 = rot1 = rot2 = sentr = gv0 = gv1 = gv2 = 0;
	  numngb = 0;
	}
      int e = 0;
      while(e < nentries)
	{
	  const int nx = __ldg(&e_next[e]);
	  if(nx >= 0)
	    {
	      // node: prune against the group box (ngb.c:270-289 does it per target with h + len/2)
	      const double4 B = e_B[e];
	      const double half = 0.5 * B.w;
	      const bool out_of_box = (B.x + half < bxlo) || (B.x - half > bxhi) || (B.y + half < bylo) ||
		(B.y - half > byhi) || (B.z + half < bzlo) || (B.z - half > bzhi);
	      e = out_of_box ? nx : e + 1;
	      continue;
	    }
	  const int code = -1 - nx;
	  e++;
	  if((code & 7) != 0) continue;	// ngb.c:191: only gas particles are neighbours
	  const int j = code >> 3;
	  const double4 pj = gl_ld4(&s_pm[j]);
	  if(pj.w == 0) continue;	// density.c:832
	  // density.c:835-856
	  const double dx = px - pj.x, dy = py - pj.y, dz = pz - pj.z;
	  const double r2 = dx * dx + dy * dy + dz * dz;
	  const bool hit = todo && (r2 < h2);
	  if(!__any_sync(0xffffffffu, hit)) continue;
	  const double4 vj = gl_ld4(&s_vh[j]);
	  const double entr_j = __ldg(&s_entropy[j]);
	  if(hit)
	    {
	      numngb++;
	      const double r = sqrt(r2);
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
	      const double mass_j = pj.w;
	      rho += mass_j * wk;
	      wngb += NORM_COEFF * wk / hinv3;
	      dhsmlrho += -mass_j * (3 * hinv * wk + u * dwk);
	      gv0 += mass_j * wk * vj.x;
	      gv1 += mass_j * wk * vj.y;
	      gv2 += mass_j * wk * vj.z;
	      sentr += mass_j * wk * entr_j;
	      if(r > 0)
		{
		  const double fac = mass_j * dwk / r;
		  const double dvx = vx - vj.x, dvy = vy - vj.y, dvz = vz - vj.z;
		  divv += -fac * (dx * dvx + dy * dvy + dz * dvz);
		  rot0 += fac * (dz * dvy - dy * dvz);
		  rot1 += fac * (dx * dvz - dz * dvx);
		  rot2 += fac * (dy * dvx - dx * dvy);
		}
	    }
	}

      // ---- post-processing and Hsml update, density.c:434-652 (one lane = one target) ----
      if(todo)
	{
	  bool redo = (wngb < (desnumngb - par.maxdev)) || (wngb > (desnumngb + par.maxdev) && h > (1.01 * par.mingashsml));
	  if(redo && left > 0 && right > 0 && (right - left) < 1.0e-3 * left) redo = false;	// density.c:566-573
	  if(redo)
	    {
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
			  double fac = 1 - (wngb - desnumngb) / (3 * wngb) * dfac;
			  h *= (fac < 1.26) ? fac : 1.26;
			}
		      else
			h *= 1.26;
		    }
		  if(right > 0 && left == 0)
		    {
		      if(ptype == 0 && fabs(wngb - desnumngb) < 0.5 * desnumngb)
			{
			  double fac = 1 - (wngb - desnumngb) / (3 * wngb) * dfac;
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
	    }
	  else
	    todo = false;
	}
      iter++;
      if(iter > GL_MAXITER)
	{			// density.c:669-674 endrun(1155)
	  if(todo) atomicMax(&counter[2], 1155ull);
	  break;
	}
    }

  if(live)
    {
      out.hsml_in[pin] = h;
      s_vh[p].w = h;
      out.numngb[pin] = wngb;
      out.pairs[pin] = numngb;
      if(ptype == 0)
	{
	  e_B[(long long) p + noff[p + 1]].x = h;	// live Hsml for the gravity softening, forcetree.c:2046-2054
	  double dfac = dhsmlrho, curl = 0, dv = divv;
	  if(rho > 0)
	    {
	      dfac *= h / (3 * rho);
	      dfac = (dfac > -0.9) ? 1 / (1 + dfac) : 1;
	      curl = sqrt(rot0 * rot0 + rot1 * rot1 + rot2 * rot2) / rho;
	      dv = divv / rho;
	    }
	  const int tb = in_timebin[pin];
	  const int dt_step = tb ? (1 << tb) : 0;
	  const double dt_entr = (par.ti_current - (in_tibegstep[pin] + dt_step / 2)) * par.timebase_interval;
	  out.density[pin] = rho;
	  out.dhsml[pin] = dfac;
	  out.divvel[pin] = dv;
	  out.curlvel[pin] = curl;
	  out.pressure[pin] = (in_entropy[pin] + in_dtentropy[pin] * dt_entr) * pow(rho, GL_GAMMA);
	}
      else
	{			// density.c:520-546 sink (Type 5) / dust (Type 2) post-processing
	  double se = sentr, g0 = gv0, g1 = gv1, g2 = gv2;
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
  const unsigned maxit = __reduce_max_sync(0xffffffffu, (unsigned) iter);
  if(lane == 0) atomicMax(&counter[1], (unsigned long long) maxit);
}

__global__ void k_gather_d(long long n, const int *__restrict__ perm, const double *__restrict__ in, double *__restrict__ out)
{
  long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i < n) out[i] = in[perm[i]];
}

int gls_density(gl_ctx *c)
{
  if(!c->tree_valid) return c->fail("density: no tree (call gl_force_treebuild / gl_gravity_tree first, init.c:529)");
  if(c->par.ComovingIntegrationOn) return c->fail("density: ComovingIntegrationOn != 0 is not implemented");
  if(gl_prepare_targets(c)) return 1;
  long long n = c->n, nt = c->n_tgt_sph;
  GL_CUDA(c, c->o_numngb.ensure(n));
  GL_CUDA(c, c->o_density.ensure(n));
  GL_CUDA(c, c->o_dhsml.ensure(n));
  GL_CUDA(c, c->o_divvel.ensure(n));
  GL_CUDA(c, c->o_curlvel.ensure(n));
  GL_CUDA(c, c->o_pressure.ensure(n));
  GL_CUDA(c, c->o_sinkrho.ensure(n));
  GL_CUDA(c, c->o_sinkentr.ensure(n));
  GL_CUDA(c, c->o_sinkvel.ensure(3 * n));
  GL_CUDA(c, c->o_denspairs.ensure(n));
  GL_CUDA(c, c->d_counter.ensure(8));
  GL_CUDA(c, c->h_q.ensure(n + 1));	// scratch: sorted entropy lives in h_r until hydro overwrites it
  GL_CUDA(c, c->h_r.ensure(n + 1));
  double *s_entropy = (double *) c->h_r.p;
  GL_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, 8 * sizeof(unsigned long long), c->stream));
  cudaEventRecord(c->ev[0], c->stream);
  k_gather_d<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, c->perm.p, c->in_entropy.p, s_entropy);
  GL_LAUNCH_COUNT(c);
  DensOut out{c->in_hsml.p, c->o_numngb.p, c->o_density.p, c->o_dhsml.p, c->o_divvel.p, c->o_curlvel.p, c->o_pressure.p,
	      c->o_sinkrho.p, c->o_sinkentr.p, c->o_sinkvel.p, c->o_denspairs.p};
  gl_dparams dp = gl_make_dparams(c);
  if(nt > 0)
    {
      k_density<<<gl_grid(nt, GL_BLOCK), GL_BLOCK, 0, c->stream>>>(c->tgt_sph.p, nt, (int) c->nentries, c->e_next.p, c->e_B.p,
								  c->s_pm.p, c->s_vh.p, c->s_type.p, s_entropy, c->noff.p,
								  c->perm.p, c->in_entropy.p, c->in_dtentropy.p,
								  c->in_timebin.p, c->in_tibegstep.p, out, c->d_counter.p, dp);
      GL_LAUNCH_COUNT(c);
    }
  cudaEventRecord(c->ev[1], c->stream);
  unsigned long long cnt[4];
  GL_CUDA(c, cudaMemcpyAsync(cnt, c->d_counter.p, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  GL_CUDA(c, cudaGetLastError());
  float ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->st.ms_density = ms;
  c->st.density_targets = nt;
  c->st.density_passes_max = (long long) cnt[1];
  if(cnt[2] == 1155) return c->fail("density: failed to converge in neighbour iteration (endrun(1155), density.c:669)");
  if(cnt[2] == 8188) return c->fail("density: Left == Right == 0 (endrun(8188), density.c:608)");
  return 0;
}

// ---------------------------------------------------------------- hydro ----
__global__ void k_hydro_prep(long long n, const int *__restrict__ perm, const int *__restrict__ s_type,
			     const double4 *__restrict__ s_vh, const double *__restrict__ density,
			     const double *__restrict__ pressure, const double *__restrict__ dhsml,
			     const double *__restrict__ divvel, const double *__restrict__ curlvel,
			     const int *__restrict__ in_timebin, double4 *__restrict__ h_q, double2 *__restrict__ h_r)
{
  long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= n) return;
  if(s_type[i] != 0)
    {
      h_q[i] = make_double4(0, 0, 0, 0);
      h_r[i] = make_double2(0, 0);
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
  h_r[i] = make_double2(f, (double) (tb ? (1 << tb) : 0));
}

__global__ void __launch_bounds__(GL_BLOCK)
k_hydro(const int *__restrict__ tgt, long long nt, int nentries, const int *__restrict__ e_next,
	const double4 *__restrict__ e_B, const double *__restrict__ e_hmax, const double4 *__restrict__ s_pm,
	const double4 *__restrict__ s_vh, const double4 *__restrict__ h_q, const double2 *__restrict__ h_r,
	const int *__restrict__ perm, const double *__restrict__ pressure, const double *__restrict__ dhsml,
	const double *__restrict__ divvel, const double *__restrict__ curlvel, double *__restrict__ o_hacc,
	double *__restrict__ o_dtentr, double *__restrict__ o_maxsig, int *__restrict__ o_pairs, gl_dparams par)
{
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long) blockDim.x + threadIdx.x) >> 5;
  const long long t = warp * 32 + lane;
  const bool live = t < nt;
  int p = 0, pin = 0;
  double px = 0, py = 0, pz = 0, vx = 0, vy = 0, vz = 0, h_i = 0, mass = 0, rho = 1, f1 = 0, cs_i = 0, p_over_rho2_i = 0;
  int timestep = 0;
  if(live)
    {
      // hydra.c:958-979 target load
      p = tgt[t];
      pin = perm[p];
      const double4 pm = s_pm[p], vh = s_vh[p];
      px = pm.x; py = pm.y; pz = pm.z; mass = pm.w;
      vx = vh.x; vy = vh.y; vz = vh.z; h_i = vh.w;
      rho = h_q[p].x;
      const double P = pressure[pin];
      timestep = (int) h_r[p].y;
      cs_i = sqrt(GL_GAMMA * P / rho);
      const double adv = fabs(divvel[pin]);
      f1 = adv / (adv + curlvel[pin] + 0.0001 * cs_i / h_i / 1.0);
      p_over_rho2_i = P / (rho * rho);	// hydra.c:1154-1155
      p_over_rho2_i *= dhsml[pin];
    }
  const double h_i2 = h_i * h_i;
  const double bxlo = warp_min_d(live ? px : 1e300), bxhi = warp_max_d(live ? px : -1e300);
  const double bylo = warp_min_d(live ? py : 1e300), byhi = warp_max_d(live ? py : -1e300);
  const double bzlo = warp_min_d(live ? pz : 1e300), bzhi = warp_max_d(live ? pz : -1e300);
  const double hgrp = warp_max_d(live ? h_i : 0.0);
  double a0 = 0, a1 = 0, a2 = 0, dtE = 0, maxsig = 0;
  int npairs = 0;
  int e = 0;
  while(e < nentries)
    {
      const int nx = __ldg(&e_next[e]);
      if(nx >= 0)
	{
	  // ngb.c:136-152: search radius max(hmax_node, h_i); here max over the warp's targets
	  const double4 B = gl_ld4(&e_B[e]);
	  const double reach = fmax(__ldg(&e_hmax[e]), hgrp) * BOX_SLACK + 0.5 * B.w;
	  const bool out_of_box = (B.x + reach < bxlo) || (B.x - reach > bxhi) || (B.y + reach < bylo) ||
	    (B.y - reach > byhi) || (B.z + reach < bzlo) || (B.z - reach > bzhi);
	  e = out_of_box ? nx : e + 1;
	  continue;
	}
      const int code = -1 - nx;
      e++;
      if((code & 7) != 0) continue;	// ngb.c:54
      const int j = code >> 3;
      const double4 pj = gl_ld4(&s_pm[j]);
      if(pj.w == 0) continue;	// hydra.c:1236
      const double4 vj = gl_ld4(&s_vh[j]);
      const double dx = px - pj.x, dy = py - pj.y, dz = pz - pj.z;
      const double r2 = dx * dx + dy * dy + dz * dz;
      const double h_j = vj.w;
      const bool hit = live && (r2 < h_i2 || r2 < h_j * h_j);	// hydra.c:1266
      if(!__any_sync(0xffffffffu, hit)) continue;
      const double4 q = gl_ld4(&h_q[j]);
      const double2 rr = __ldg(&h_r[j]);
      if(hit)
	{
	  const double r = sqrt(r2);
	  if(r > 0)
	    {
	      npairs++;
	      const double rho_j = q.x;
	      double p_over_rho2_j = q.y;
	      const double soundspeed_j = q.w;
	      const double dvx = vx - vj.x, dvy = vy - vj.y, dvz = vz - vj.z;
	      const double vdotr = dx * dvx + dy * dvy + dz * dvz;
	      const double vdotr2 = vdotr;	// non-comoving, hydra.c:1279-1282
	      double dwk_i = 0, dwk_j = 0;
	      if(r2 < h_i2)
		{
		  const double hinv = 1.0 / h_i;
		  const double hinv4 = hinv * hinv * hinv * hinv;
		  const double u = r * hinv;
		  if(u < 0.5)
		    dwk_i = hinv4 * u * (KERNEL_COEFF_3 * u - KERNEL_COEFF_4);
		  else
		    dwk_i = hinv4 * KERNEL_COEFF_6 * (1.0 - u) * (1.0 - u);
		}
	      if(r2 < h_j * h_j)
		{
		  const double hinv = 1.0 / h_j;
		  const double hinv4 = hinv * hinv * hinv * hinv;
		  const double u = r * hinv;
		  if(u < 0.5)
		    dwk_j = hinv4 * u * (KERNEL_COEFF_3 * u - KERNEL_COEFF_4);
		  else
		    dwk_j = hinv4 * KERNEL_COEFF_6 * (1.0 - u) * (1.0 - u);
		}
	      double vsig = cs_i + soundspeed_j;	// hydra.c:1502
	      if(vsig > maxsig) maxsig = vsig;
	      double visc = 0;
	      if(vdotr2 < 0)
		{
		  const double mu_ij = 1.0 * vdotr2 / r;	// fac_mu = 1
		  vsig -= 3 * mu_ij;
		  if(vsig > maxsig) maxsig = vsig;
		  const double rho_ij = 0.5 * (rho + rho_j);
		  const double f2 = rr.x;
		  visc = 0.25 * par.artvisc * vsig * (-mu_ij) / rho_ij * (f1 + f2);
		  // hydra.c:1583-1597 viscosity limiter
		  const int ts_j = (int) rr.y;
		  const double dt = 2 * ((timestep > ts_j) ? timestep : ts_j) * par.timebase_interval;
		  if(dt > 0 && (dwk_i + dwk_j) < 0)
		    {
		      const double lim = 0.5 * 1.0 * vdotr2 / (0.5 * (mass + pj.w) * (dwk_i + dwk_j) * r * dt);
		      visc = (visc < lim) ? visc : lim;
		    }
		}
	      p_over_rho2_j = q.z;	// already x DhsmlDensityFactor_j
	      const double hfc_visc = 0.5 * pj.w * visc * (dwk_i + dwk_j) / r;
	      const double hfc = hfc_visc + pj.w * (p_over_rho2_i * dwk_i + p_over_rho2_j * dwk_j) / r;
	      a0 += (-hfc * dx);
	      a1 += (-hfc * dy);
	      a2 += (-hfc * dz);
	      dtE += (0.5 * hfc_visc * vdotr2);
	    }
	}
    }
  if(live)
    {
      o_hacc[3ll * pin] = a0;
      o_hacc[3ll * pin + 1] = a1;
      o_hacc[3ll * pin + 2] = a2;
      // hydra.c:583: DtEntropy *= GAMMA_MINUS1 / (hubble_a2 * rho^GAMMA_MINUS1)
      o_dtentr[pin] = dtE * (GL_GAMMA_MINUS1 / (1.0 * pow(rho, GL_GAMMA_MINUS1)));
      o_maxsig[pin] = maxsig;
      o_pairs[pin] = npairs;
    }
}

int gls_hydro(gl_ctx *c)
{
  if(!c->tree_valid) return c->fail("hydro_force: no tree");
  if(c->par.ComovingIntegrationOn) return c->fail("hydro_force: ComovingIntegrationOn != 0 is not implemented");
  if(c->o_density.cap < (size_t) c->n) return c->fail("hydro_force: density() has not been run (SphP.Density unknown)");
  if(gl_prepare_targets(c)) return 1;
  long long n = c->n, nt = c->n_tgt_gas;
  GL_CUDA(c, c->o_hacc.ensure(3 * n));
  GL_CUDA(c, c->o_dtentr.ensure(n));
  GL_CUDA(c, c->o_maxsig.ensure(n));
  GL_CUDA(c, c->o_hydropairs.ensure(n));
  GL_CUDA(c, c->h_q.ensure(n + 1));
  GL_CUDA(c, c->h_r.ensure(n + 1));
  gl_dparams dp = gl_make_dparams(c);
  cudaEventRecord(c->ev[0], c->stream);
  k_hydro_prep<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, c->perm.p, c->s_type.p, c->s_vh.p, c->o_density.p, c->o_pressure.p,
							c->o_dhsml.p, c->o_divvel.p, c->o_curlvel.p, c->in_timebin.p, c->h_q.p,
							c->h_r.p);
  GL_LAUNCH_COUNT(c);
  if(nt > 0)
    {
      k_hydro<<<gl_grid(nt, GL_BLOCK), GL_BLOCK, 0, c->stream>>>(c->tgt_gas.p, nt, (int) c->nentries, c->e_next.p, c->e_B.p,
								c->e_hmax.p, c->s_pm.p, c->s_vh.p, c->h_q.p, c->h_r.p, c->perm.p,
								c->o_pressure.p, c->o_dhsml.p, c->o_divvel.p, c->o_curlvel.p,
								c->o_hacc.p, c->o_dtentr.p, c->o_maxsig.p, c->o_hydropairs.p, dp);
      GL_LAUNCH_COUNT(c);
    }
  cudaEventRecord(c->ev[1], c->stream);
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  GL_CUDA(c, cudaGetLastError());
  float ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
  c->st.ms_hydro = ms;
  c->st.hydro_targets = nt;
  return 0;
}
