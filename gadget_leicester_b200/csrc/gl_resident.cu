// gl_resident.cu - run()'s loop (run.c:63-110) with the particle state resident in HBM (device side).
//
// The reference alternates find_next_sync_point_and_drift (run.c:190) -> domain_Decomposition -> compute_accelerations
// (accel.c:27) -> advance_and_find_timesteps (timestep.c:29) on P[] / SphP[] in host memory.  The drop-in call
// gl_gadget_compute_accelerations moves those structs through the host's DRAM twice per step, which costs more than
// the force computation.  Here the state the loop carries from one step to the next stays in the
// context's input-order arrays (in_pos, in_vel, in_velpred, in_hsml, in_entropy, in_dtentropy, in_timebin,
// in_tibegstep, in_oldacc, o_acc, o_hacc, o_maxsig, o_density, o_pressure, o_dhsml, o_divvel, o_curlvel) and three
// kernels restate the per-particle statements of the host-side pieces on it:
//   k_run_drift   drift_particle (predict.c:129-260) for every particle to the new All.Ti_Current - the reference
//                 drifts the active ones at the sync point (run.c:325) and all others in domain_Decomposition
//                 (domain.c:122) or lazily from the walks (forcetree.c:1912, ngb.c:58); with a new decomposition
//                 and tree in every step (how this path runs) that is one drift of everything per step;
//   k_run_kick    get_timestep (timestep.c:607) -> power-of-two bin with the "only onto a synchronised bin" rule
//                 (timestep.c:136-154) -> do_the_kick (timestep.c:364) for the particles on active bins;
//   k_run_bins    TimeBinCount[] (timestep.c:170-233 keeps it by relinking lists; a histogram is all the next sync
//                 point needs, run.c:199-216).
// The statements and their order are those of gl_gadget_drift / gl_gadget_advance_and_find_timesteps (host, bit-
// identical to the reference); on the device exp() / pow() / sqrt() are CUDA's (<= 2 ulp from glibc's), so the
// floating-point state agrees to rounding and the integer time line agrees except for a step that falls within an
// ulp of a power-of-two boundary.  Checked against eight iterations of the reference's own loop (harness op L) by
// tests/test_gpu_resident.py.  Non-comoving integration, no dynamic tree update (the tree is rebuilt every step).
#include "gl_internal.h"
#include <cub/cub.cuh>

namespace
{
const int TIMEBASE = 1 << 29;	// allvars.h:39-41 (TIMEBINS 29)

struct RunPar
{
  int ti_current;
  unsigned active_mask;		// bit b: TimeBinActive[b]
  double tbi, errtolint, courant, maxstep, minstep, minegyspec, gm1;
  double softtab[6];
  double outer, unitvel, feedvel, smbh, inner, sinkb, accdt;
};

RunPar make_runpar(const gl_timestep_params *T)
{
  RunPar r{};
  r.ti_current = T->Ti_Current;
  for(int b = 0; b < 32; b++)
    if(T->TimeBinActive[b]) r.active_mask |= (1u << b);
  r.tbi = T->Timebase_interval; r.errtolint = T->ErrTolIntAccuracy; r.courant = T->CourantFac;
  r.maxstep = T->MaxSizeTimestep; r.minstep = T->MinSizeTimestep; r.minegyspec = T->MinEgySpec; r.gm1 = T->GammaMinus1;
  for(int i = 0; i < 6; i++) r.softtab[i] = T->SofteningTable[i];
  r.outer = T->OuterBoundary; r.unitvel = T->UnitVelocity_in_cm_per_s; r.feedvel = T->FeedBackVelocity;
  r.smbh = T->SMBHmass; r.inner = T->InnerBoundary; r.sinkb = T->SinkBoundary; r.accdt = T->AccDtBlackHole;
  return r;
}

__device__ __forceinline__ void run_fail(int *err, int code, long long p)
{
  if(atomicCAS(&err[0], 0, code) == 0) err[1] = (int) p;	// keep the first
}

// predict.c:129-260 for particle i, from time0 to time1 (every particle sits at time0: the loop drifts all of them)
__global__ void k_run_drift(long long n, int time0, int time1, double tbi, double mingashsml, const int *__restrict__ type,
			    double *__restrict__ pos, const double *__restrict__ vel, double *__restrict__ velpred,
			    const double *__restrict__ gacc, const double *__restrict__ hacc, double *__restrict__ density,
			    double *__restrict__ hsml, double *__restrict__ pressure, const double *__restrict__ divvel,
			    const double *__restrict__ entropy, const double *__restrict__ dtentropy, const int *__restrict__ timebin,
			    const int *__restrict__ tibegstep)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= n) return;
  const int ty = type[i];
  if(ty == 3) return;		// predict.c:134 (-DVIRTUAL)
  const double dt = (time1 - time0) * tbi;	// predict.c:165: dt_drift = dt_gravkick = dt_hydrokick
#pragma unroll
  for(int j = 0; j < 3; j++) pos[3 * i + j] = __dadd_rn(pos[3 * i + j], __dmul_rn(vel[3 * i + j], dt));	// predict.c:169-170
  if(ty == 0)
    {
#pragma unroll
      for(int j = 0; j < 3; j++)	// predict.c:185-186 (DragAccel stays 0 here: the dust loops are not part of this loop)
	velpred[3 * i + j] = __dadd_rn(velpred[3 * i + j], __dadd_rn(__dmul_rn(gacc[3 * i + j], dt), __dmul_rn(hacc[3 * i + j], dt)));
      const double divv = divvel[i];
      const double rho = __dmul_rn(density[i], exp(__dmul_rn(-divv, dt)));	// predict.c:204
      double h = __dmul_rn(hsml[i], exp(__dmul_rn(__dmul_rn(0.333333333333, divv), dt)));	// predict.c:205
      if(h < mingashsml) h = mingashsml;
      density[i] = rho;
      hsml[i] = h;
      const int tb = timebin[i];
      const int dt_step = tb ? (1 << tb) : 0;
      const double dt_entr = (time1 - (tibegstep[i] + dt_step / 2)) * tbi;	// predict.c:213
      pressure[i] = __dmul_rn(__dadd_rn(entropy[i], __dmul_rn(dtentropy[i], dt_entr)), pow(rho, 7.0 / 5));	// predict.c:217
    }
}

// get_timestep(p, &aphys, 0), timestep.c:607-1122, on the SoA state; < 0 after a failure
__device__ int run_get_timestep(const RunPar &T, long long p, int type, double mass, double hsml, double ax, double ay, double az,
				double maxsig, int *err)
{
  double ac = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az)));
  if(ac == 0) ac = 1.0e-30;
  // TypeOfTimestepCriterion 0, timestep.c:676-708 (DeltaDustMomentum is zero without the dust loops: timestep.c:681-701
  // then re-forms the same ac and dt)
  double dt = sqrt(2 * T.errtolint * 1.0 * T.softtab[type] / ac);
  if(type == 2) dt = dt / 2;
  if(type == 0) dt = sqrt(2 * T.errtolint * 1.0 * hsml / 2.8 / ac);
  if(type == 0)
    {
      const double dt_courant = 2 * T.courant * hsml / maxsig;	// timestep.c:853-866
      if(dt_courant < dt) dt = dt_courant;
    }
  if(type == 3)
    {
      // timestep.c:889-900 (NewDensity is zero for every particle this loop can hold; see gl_run_begin)
      double dt_abs = 0.03 * T.outer / 2.9979e10 * T.unitvel * T.feedvel;
      const double dt_ff = 1.;
      if(dt_abs > dt_ff) dt_abs = dt_ff;
      if(dt > dt_abs) dt = dt_abs;
    }
  if(type == 5)
    {
      // sink particles, timestep.c:909-950
      double dt_accr = 0.03 * (T.outer / 100.), dt_a = 1.e10;
      if(mass >= 0.45 * T.smbh && T.inner > 0) dt_a = T.accdt * pow((T.inner + 0.5 * hsml), 1.5) / pow(mass, 0.5);
      if(mass < 0.45 * T.smbh && T.sinkb > 0) dt_a = T.accdt * pow((T.sinkb + 0.5 * hsml), 1.5) / pow(mass, 0.5);
      if(dt_accr > dt_a) dt_accr = dt_a;
      if(dt_accr < dt) dt = dt_accr;
    }
  if(dt >= T.maxstep) dt = T.maxstep;	// timestep.c:1042-1122 (hubble_a = 1, dt_displacement = MaxSizeTimestep)
  if(dt < T.minstep)
    {
      run_fail(err, 888, p);
      return -1;
    }
  const int ti_step = (int) (dt / T.tbi);
  if(!(ti_step > 0 && ti_step < TIMEBASE))
    {
      run_fail(err, 818, p);
      return -1;
    }
  return ti_step;
}

// the per-particle body of advance_and_find_timesteps (timestep.c:127-168, :236-260) + do_the_kick (timestep.c:364-601)
__global__ void k_run_kick(long long n, RunPar T, const int *__restrict__ type, const double *__restrict__ mass,
			   const double *__restrict__ hsml, const double *__restrict__ gacc, const double *__restrict__ hacc,
			   const double *__restrict__ maxsig, const double *__restrict__ density, double *__restrict__ vel,
			   double *__restrict__ velpred, double *__restrict__ entropy, double *__restrict__ dtentropy,
			   const double *__restrict__ dtentr_new, int *__restrict__ timebin, int *__restrict__ tibegstep, int *__restrict__ err)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= n) return;
  const int binold = timebin[i];
  if(binold < 0 || binold > 31 || !((T.active_mask >> binold) & 1u)) return;
  const int ty = type[i];
  const bool gas = ty == 0;
  const double g0 = gacc[3 * i], g1 = gacc[3 * i + 1], g2 = gacc[3 * i + 2];
  double h0 = 0, h1 = 0, h2 = 0;
  if(gas)
    {
      h0 = hacc[3 * i]; h1 = hacc[3 * i + 1]; h2 = hacc[3 * i + 2];
      // the hydro force of this step belongs to this particle: SphP.e.DtEntropy <- hydra.c:785-813
      if(dtentr_new) dtentropy[i] = dtentr_new[i];
    }
  // timestep.c:636-650 (fac1 = 1; DragAccel = 0)
  const double ax = gas ? __dadd_rn(g0, h0) : g0, ay = gas ? __dadd_rn(g1, h1) : g1, az = gas ? __dadd_rn(g2, h2) : g2;
  int ti_step = run_get_timestep(T, i, ty, mass[i], hsml[i], ax, ay, az, gas ? maxsig[i] : 1.0, err);
  if(ti_step < 0) return;
  int ti_min = TIMEBASE;	// timestep.c:141-144
  while(ti_min > ti_step) ti_min >>= 1;
  ti_step = ti_min;
  if(ti_step == 1)
    {
      run_fail(err, 112313, i);	// get_timestep_bin, timestep.c:1226-1246
      return;
    }
  int bin = -1;
  for(int t = ti_step; t; t >>= 1) bin++;
  if(ti_step == 0) bin = 0;
  if(bin > binold && !((T.active_mask >> bin) & 1u))	// timestep.c:147-154
    {
      bin = binold;
      ti_step = bin ? (1 << bin) : 0;
    }
  if(T.ti_current >= TIMEBASE)
    {
      ti_step = 0;
      bin = 0;
    }
  if((TIMEBASE - T.ti_current) < ti_step)
    {
      run_fail(err, 889, i);
      return;
    }
  if(bin != binold) timebin[i] = bin;
  const int ti_step_old = binold ? (1 << binold) : 0;
  const int beg = tibegstep[i];
  const int tstart = beg + ti_step_old / 2, tend = beg + ti_step_old + ti_step / 2, tcur = beg + ti_step_old;
  tibegstep[i] = tcur;
  // ---- do_the_kick(i, tstart, tend, tcur) ----
  if(ty == 3) return;		// -DVIRTUAL, timestep.c:379
  const double dt_kick = (tend - tstart) * T.tbi, dt_kick2 = (tend - tcur) * T.tbi;
  double v0 = vel[3 * i], v1 = vel[3 * i + 1], v2 = vel[3 * i + 2];
  if(ty != 2)			// dust is moved by the drag module (timestep.c:478-485): dv = 0
    {
      v0 = __dadd_rn(v0, __dmul_rn(g0, dt_kick));
      v1 = __dadd_rn(v1, __dmul_rn(g1, dt_kick));
      v2 = __dadd_rn(v2, __dmul_rn(g2, dt_kick));
    }
  if(gas)
    {
      v0 = __dadd_rn(v0, __dmul_rn(h0, dt_kick));
      v1 = __dadd_rn(v1, __dmul_rn(h1, dt_kick));
      v2 = __dadd_rn(v2, __dmul_rn(h2, dt_kick));
      velpred[3 * i] = __dsub_rn(__dsub_rn(v0, __dmul_rn(dt_kick2, g0)), __dmul_rn(dt_kick2, h0));
      velpred[3 * i + 1] = __dsub_rn(__dsub_rn(v1, __dmul_rn(dt_kick2, g1)), __dmul_rn(dt_kick2, h1));
      velpred[3 * i + 2] = __dsub_rn(__dsub_rn(v2, __dmul_rn(dt_kick2, g2)), __dmul_rn(dt_kick2, h2));
      double A = entropy[i], dA = dtentropy[i];
      if(__dmul_rn(dA, dt_kick) > __dmul_rn(-0.5, A))	// timestep.c:546-551
	A = __dadd_rn(A, __dmul_rn(dA, dt_kick));
      else
	A = __dmul_rn(A, 0.5);
      if(T.minegyspec != 0)
	{
	  const double minentropy = T.minegyspec * T.gm1 / pow(density[i], T.gm1);
	  if(A < minentropy)
	    {
	      A = minentropy;
	      dA = 0;
	    }
	}
      // the same limit for the predicted entropy at the middle of the NEW step (timestep.c:577-583)
      const double dt_entr = (bin ? (1 << bin) : 0) / 2 * T.tbi;
      if(__dadd_rn(A, __dmul_rn(dA, dt_entr)) < __dmul_rn(0.5, A)) dA = __dmul_rn(-0.5, A) / dt_entr;
      entropy[i] = A;
      dtentropy[i] = dA;
    }
  vel[3 * i] = v0; vel[3 * i + 1] = v1; vel[3 * i + 2] = v2;
}

__global__ void k_run_bins(long long n, const int *__restrict__ timebin, int *__restrict__ hist)
{
  __shared__ int sh[32];
  if(threadIdx.x < 32) sh[threadIdx.x] = 0;
  __syncthreads();
  for(long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
    {
      const int b = timebin[i];
      if(b >= 0 && b < 32) atomicAdd(&sh[b], 1);
    }
  __syncthreads();
  if(threadIdx.x < 32 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

__global__ void k_run_active_flags(long long n, unsigned mask, const int *__restrict__ timebin, int *__restrict__ flag)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= n) return;
  const int b = timebin[i];
  flag[i] = (b >= 0 && b < 32 && ((mask >> b) & 1u)) ? 1 : 0;
}

// VelPred gets a row per particle: rows the caller supplied (the leading n_have, SphP[i] belongs to P[i]) are kept, the
// others start from the velocities (they are only read where the type is 0)
__global__ void k_run_fill_velpred(long long n, long long n_have, const double *__restrict__ vel, double *__restrict__ velpred)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i < 3 * n && i >= 3 * n_have) velpred[i] = vel[i];
}

int run_check_err(gl_ctx *c, const char *where)
{
  int h[2] = { 0, 0 };
  GL_CUDA(c, cudaMemcpyAsync(h, c->run_err.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  GL_CUDA(c, cudaGetLastError());
  if(c->comm && c->comm->world > 1)
    {
      // the reference aborts all ranks (endrun): every rank has to learn that one of them failed
      int any = 0;
      if(gll_agree_violation(c, h[0] != 0, &any)) return 1;
      if(any && h[0] == 0) return c->fail(std::string(where) + ": another rank stopped (endrun)");
    }
  if(h[0] == 0) return 0;
  const char *what = h[0] == 888 ? "timestep below MinSizeTimestep (endrun 888)"
    : h[0] == 818 ? "a timestep of size zero was assigned on the integer timeline (endrun 818)"
    : h[0] == 112313 ? "time-step of integer size 1 not allowed (endrun 112313)" : "step beyond the end of the time line (endrun 888)";
  return c->fail(std::string(where) + ": " + what + ", particle " + std::to_string(h[1]));
}

int run_kick(gl_ctx *c, const gl_timestep_params *tp, bool commit_dtentropy)
{
  const long long n = c->n_own;
  const RunPar T = make_runpar(tp);
  GL_CUDA(c, cudaMemsetAsync(c->run_err.p, 0, 2 * sizeof(int), c->stream));
  k_run_kick<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, T, c->in_type.p, c->in_mass.p, c->in_hsml.p, c->o_acc.p, c->o_hacc.p,
						       c->o_maxsig.p, c->o_density.p, c->in_vel.p, c->in_velpred.p, c->in_entropy.p,
						       c->in_dtentropy.p, commit_dtentropy ? c->o_dtentr.p : nullptr, c->in_timebin.p,
						       c->in_tibegstep.p, c->run_err.p);
  GL_LAUNCH_COUNT(c);
  return run_check_err(c, "advance_and_find_timesteps");
}
}				// namespace

// ---------------------------------------------------------------- C ABI ----
extern "C" int gl_set_velpred(gl_ctx *c, const double *velpred3, long long n_gas, int where)
{
  cudaSetDevice(c->device);
  const long long n_all = c->n_own > 0 ? c->n_own : c->n;	// (multi-rank mode: imports of a tree are not the caller's particles)
  if(c->n <= 0) return c->fail("gl_set_velpred: no particles");
  if(n_gas < 0 || n_gas > n_all || (n_gas > 0 && !velpred3)) return c->fail("gl_set_velpred: bad arguments");
  GL_CUDA(c, c->in_velpred.ensure(3 * (size_t) std::max<long long>(n_gas, 1)));
  if(n_gas > 0)
    GL_CUDA(c, cudaMemcpyAsync(c->in_velpred.p, velpred3, 3 * n_gas * sizeof(double), where ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
			       c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n_velpred = n_gas;
  c->have_velpred = n_gas > 0;
  c->run_ready = false;		// gl_run_begin gives VelPred a row per particle again
  if(c->tree_valid && glk_apply_velpred(c)) return 1;
  return 0;
}

extern "C" int gl_set_accelerations(gl_ctx *c, const double *gravaccel3, const double *hydroaccel3, const double *maxsignalvel, int where)
{
  cudaSetDevice(c->device);
  const long long n = c->n_own > 0 ? c->n_own : c->n;	// (multi-rank mode: imports of a tree are not the caller's particles)
  if(n <= 0) return c->fail("gl_set_accelerations: no particles");
  const cudaMemcpyKind k = where ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  if(gravaccel3)
    {
      GL_CUDA(c, c->o_acc.ensure_keep(3 * (size_t) n));
      GL_CUDA(c, cudaMemcpyAsync(c->o_acc.p, gravaccel3, 3 * n * sizeof(double), k, c->stream));
    }
  if(hydroaccel3)
    {
      GL_CUDA(c, c->o_hacc.ensure_keep(3 * (size_t) n));
      GL_CUDA(c, cudaMemcpyAsync(c->o_hacc.p, hydroaccel3, 3 * n * sizeof(double), k, c->stream));
    }
  if(maxsignalvel)
    {
      GL_CUDA(c, c->o_maxsig.ensure_keep((size_t) n));
      GL_CUDA(c, cudaMemcpyAsync(c->o_maxsig.p, maxsignalvel, n * sizeof(double), k, c->stream));
    }
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// The state the loop needs must be in the context: the particle set (gl_load_particles), the SphP state
// (gl_set_sph_state or a density() of all gas particles) and the accelerations (gl_set_accelerations or a force
// computation of all particles).  Sizes are checked here; VelPred defaults to the velocities that were loaded.
// Multi-rank mode: every rank holds its own particles [0, n_own); the imports of a locally-essential tree lie behind
// them and are never drifted or kicked here - they come again with every rebuild.
extern "C" int gl_run_begin(gl_ctx *c, int Ti_Current)
{
  cudaSetDevice(c->device);
  const long long n = c->n_own;
  if(c->n <= 0 || n <= 0) return c->fail("gl_run_begin: no particles loaded");
  if(!c->have_par) return c->fail("gl_run_begin: gl_set_params not called");
  const int W = c->comm ? c->comm->world : 1;
  GL_CUDA(c, c->run_err.ensure(48));
  GL_CUDA(c, c->run_hist.ensure(32 * (size_t) (W + 1)));
  if(c->o_acc.cap < 3 * (size_t) n) return c->fail("gl_run_begin: no gravitational accelerations in the context (force computation or gl_set_accelerations)");
  // how many gas particles: a set without gas never allocates the SphP arrays (the kernels only read them where the type is 0)
  GL_CUDA(c, cudaMemsetAsync(c->run_err.p, 0, 48 * sizeof(int), c->stream));
  k_run_bins<<<148 * 4, 256, 0, c->stream>>>(n, c->in_type.p, c->run_err.p + 8);	// histogram of the types: slot 8 = type 0
  GL_LAUNCH_COUNT(c);
  int ngas = 0;
  GL_CUDA(c, cudaMemcpyAsync(&ngas, c->run_err.p + 8, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  if(ngas > 0)
    {
      if(c->o_density.cap < (size_t) n || c->o_pressure.cap < (size_t) n || c->o_divvel.cap < (size_t) n || c->o_dhsml.cap < (size_t) n ||
	 c->o_curlvel.cap < (size_t) n)
	return c->fail("gl_run_begin: no SphP state in the context (density or gl_set_sph_state)");
      if(c->o_hacc.cap < 3 * (size_t) n || c->o_maxsig.cap < (size_t) n)
	return c->fail("gl_run_begin: no hydro accelerations / signal velocities in the context (hydro_force or gl_set_accelerations)");
      GL_CUDA(c, c->o_dtentr.ensure_keep((size_t) n));
    }
  else
    {
      // the kernels are handed these pointers whatever the types are
      GL_CUDA(c, c->o_hacc.ensure_keep(8)); GL_CUDA(c, c->o_maxsig.ensure_keep(8)); GL_CUDA(c, c->o_density.ensure_keep(8));
      GL_CUDA(c, c->o_pressure.ensure_keep(8)); GL_CUDA(c, c->o_divvel.ensure_keep(8)); GL_CUDA(c, c->o_dtentr.ensure_keep(8));
    }
  const long long have = c->have_velpred ? std::min(c->n_velpred, n) : 0;
  GL_CUDA(c, c->in_velpred.ensure_keep(3 * (size_t) n));
  k_run_fill_velpred<<<gl_grid(3 * n, 256), 256, 0, c->stream>>>(n, have, c->in_vel.p, c->in_velpred.p);
  GL_LAUNCH_COUNT(c);
  c->n_velpred = n;
  c->have_velpred = true;
  c->run_ti = Ti_Current;
  c->run_ready = true;
  return 0;
}

extern "C" int gl_run_kick(gl_ctx *c, const gl_timestep_params *tp, int commit_dtentropy)
{
  cudaSetDevice(c->device);
  if(!c->run_ready) return c->fail("gl_run_kick: gl_run_begin not called");
  if(tp->Ti_Current != c->run_ti) return c->fail("gl_run_kick: Ti_Current differs from the time the particles are drifted to");
  return run_kick(c, tp, commit_dtentropy != 0);
}

// The time-line part of find_next_sync_point_and_drift (run.c:199-216, :276-290), as a plain host function of the ranks'
// TimeBinCount[] rows (hist[32 r + b]: particles of rank r on bin b; slot 31 of a row = the particles the rank holds):
// the next kick time - the reference takes MPI_Allreduce(MPI_MIN) of the ranks' own answers (run.c:218), which is the
// answer for the summed counts -, the bins that are synchronised there, and the force updates of this rank and of all.
extern "C" int gl_next_sync_point(const int *hist, int world, int rank, int Ti_Current, int *TimeBinActive, long long *nactive_own,
				  long long *nactive_all, long long *particles_all)
{
  long long glob[32], own[32];
  for(int b = 0; b < 32; b++)
    {
      glob[b] = 0;
      for(int r = 0; r < world; r++) glob[b] += hist[32 * r + b];
      own[b] = hist[32 * rank + b];
    }
  int ti_next_kick = TIMEBASE;
  for(int b = 0; b < 29; b++)	// run.c:199-216
    if(glob[b])
      {
	const int ti_next_for_bin = b > 0 ? (Ti_Current / (1 << b)) * (1 << b) + (1 << b) : Ti_Current;
	if(ti_next_for_bin < ti_next_kick) ti_next_kick = ti_next_for_bin;
      }
  long long na = own[0], ng = glob[0];
  TimeBinActive[0] = 1;	// run.c:276-288
  for(int b = 1; b < 29; b++)
    {
      TimeBinActive[b] = (ti_next_kick % (1 << b)) == 0;
      if(TimeBinActive[b])
	{
	  na += own[b];
	  ng += glob[b];
	}
    }
  for(int b = 29; b < 32; b++) TimeBinActive[b] = 0;
  if(nactive_own) *nactive_own = na;
  if(nactive_all) *nactive_all = ng;
  if(particles_all) *particles_all = glob[31];
  return ti_next_kick;
}

extern "C" int gl_run_steps(gl_ctx *c, gl_timestep_params *tp, int nsteps, gl_run_info *info)
{
  cudaSetDevice(c->device);
  if(!c->run_ready) return c->fail("gl_run_steps: gl_run_begin not called");
  if(tp->Ti_Current != c->run_ti) return c->fail("gl_run_steps: Ti_Current differs from the time the particles are drifted to");
  const long long n = c->n_own;
  const int W = (c->comm && c->comm->world > 1) ? c->comm->world : 1;
  gl_run_info I{};
  GL_CUDA(c, c->run_list.ensure((size_t) n + 1));
  GL_CUDA(c, c->run_flag.ensure((size_t) n + 1));
  GL_CUDA(c, c->scan_tmp_i.ensure(64));
  cudaEventRecord(c->ev[12], c->stream);
  for(int step = 0; step < nsteps; step++)
    {
      // ---- find_next_sync_point_and_drift, run.c:190-330 ----
      std::vector<int> hall(32 * (size_t) W);
      int *hist = hall.data();	// own counts first; slot 31 carries the number of own particles
      GL_CUDA(c, cudaMemsetAsync(c->run_hist.p, 0, 32 * sizeof(int), c->stream));
      k_run_bins<<<148 * 4, 256, 0, c->stream>>>(n, c->in_timebin.p, c->run_hist.p);
      GL_LAUNCH_COUNT(c);
      const int n_int = (int) n;
      GL_CUDA(c, cudaMemcpyAsync(c->run_hist.p + 31, &n_int, sizeof(int), cudaMemcpyHostToDevice, c->stream));
      if(W > 1)
	{
	  // TimeBinCount[] of all ranks: the reference takes MPI_Allreduce(MPI_MIN) of the ranks' next kick times
	  // (run.c:218) and sums the force updates (run.c:290); the counts give both
	  if(c->comm->allgather(c->run_hist.p, c->run_hist.p + 32, 32 * sizeof(int), c->stream))
	    return c->fail("gl_run_steps: all-gather of TimeBinCount failed: " + c->comm->err);
	  GL_CUDA(c, cudaMemcpyAsync(hist, c->run_hist.p + 32, 32 * sizeof(int) * W, cudaMemcpyDeviceToHost, c->stream));
	}
      else
	GL_CUDA(c, cudaMemcpyAsync(hist, c->run_hist.p, 32 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      GL_CUDA(c, cudaStreamSynchronize(c->stream));
      const int time0 = tp->Ti_Current;
      long long nactive = 0, nactive_glob = 0, total = 0;
      const int ti_next_kick = gl_next_sync_point(hist, W, W > 1 ? c->comm->rank : 0, time0, tp->TimeBinActive, &nactive, &nactive_glob, &total);
      if(ti_next_kick >= TIMEBASE && time0 >= TIMEBASE) break;	// the end of the time line has been reached
      tp->Ti_Current = ti_next_kick;
      const RunPar T = make_runpar(tp);
      // drift everything to the sync point (see the header of this file)
      if(ti_next_kick != time0)
	{
	  k_run_drift<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, time0, ti_next_kick, tp->Timebase_interval, c->par.MinGasHsml, c->in_type.p,
							      c->in_pos.p, c->in_vel.p, c->in_velpred.p, c->o_acc.p, c->o_hacc.p, c->o_density.p,
							      c->in_hsml.p, c->o_pressure.p, c->o_divvel.p, c->in_entropy.p, c->in_dtentropy.p,
							      c->in_timebin.p, c->in_tibegstep.p);
	  GL_LAUNCH_COUNT(c);
	}
      c->run_ti = ti_next_kick;
      // ---- domain_Decomposition: new extent, new tree (TreeReconstructFlag, domain.c:135).  The ranks keep the
      // particles they have: any assignment gives the same results, a stale one only prunes less ----
      c->par.Ti_Current = ti_next_kick;
      if(W > 1 && !getenv("GL_NO_GAP_HULLS")) c->let_gap_hulls = true;	// the caller's order is no longer this grid's Peano-Hilbert order
      gl_set_tree_reconstruct_flag(c, 1);	// (multi-rank mode: the imports of the last tree are dropped, the next build fetches them again)
      c->tree_valid = false;
      c->tgt_dirty = true;
      if(!c->domain_imposed) c->have_domain = false;
      // ---- the active list (run.c:296-320): particles on active bins, as targets of the force path ----
      if(nactive_glob >= total)
	{
	  if(gl_set_targets(c, nullptr, -1, 0)) return 1;
	}
      else
	{
	  k_run_active_flags<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, T.active_mask, c->in_timebin.p, c->run_flag.p);
	  GL_LAUNCH_COUNT(c);
	  size_t tmp = 0;
	  int *dcount = c->scan_tmp_i.p + 59;
	  cub::CountingInputIterator<int> iota(0);
	  cub::DeviceSelect::Flagged(nullptr, tmp, iota, c->run_flag.p, c->run_list.p, dcount, (int) n, c->stream);
	  GL_CUDA(c, c->cub_tmp.ensure(tmp));
	  GL_CUDA(c, cub::DeviceSelect::Flagged(c->cub_tmp.p, tmp, iota, c->run_flag.p, c->run_list.p, dcount, (int) n, c->stream));
	  c->kernel_launches += 2;
	  int hcount = 0;
	  GL_CUDA(c, cudaMemcpyAsync(&hcount, dcount, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	  GL_CUDA(c, cudaStreamSynchronize(c->stream));
	  if(hcount != nactive) return c->fail("gl_run_steps: active list and TimeBinCount disagree (\"terrible\", run.c:333)");
	  // multi-rank mode orders its own targets (a rank's particles are its stretch of the Peano-Hilbert curve, in order)
	  if(W > 1 ? gl_set_targets_ordered(c, c->run_list.p, hcount, 1) : gl_set_targets(c, c->run_list.p, hcount, 1)) return 1;
	}
      // ---- compute_accelerations, accel.c:27-106 (collective in multi-rank mode: a rank without active particles
      // still serves the others) ----
      if(nactive_glob > 0 && gl_compute_accelerations(c, 0)) return 1;
      I.interactions += c->st.interactions;
      // ---- advance_and_find_timesteps, timestep.c:29 ----
      if(nactive_glob > 0 && run_kick(c, tp, true)) return 1;
      I.steps_done++;
      I.active_total += nactive;
      I.active_last = nactive;
    }
  cudaEventRecord(c->ev[13], c->stream);
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[12], c->ev[13]);
  I.ms_total = ms;
  I.Ti_Current = tp->Ti_Current;
  if(info) *info = I;
  return 0;
}

extern "C" int gl_get_run_state(gl_ctx *c, double *pos3, double *vel3, double *velpred3, double *hsml, double *entropy, double *dtentropy,
				int *timebin, int *ti_begstep, int where)
{
  cudaSetDevice(c->device);
  const long long n = c->n_own;
  if(c->n <= 0 || n <= 0) return c->fail("gl_get_run_state: no particles");
  const cudaMemcpyKind k = where ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if(pos3) GL_CUDA(c, cudaMemcpyAsync(pos3, c->in_pos.p, 3 * n * sizeof(double), k, c->stream));
  if(vel3) GL_CUDA(c, cudaMemcpyAsync(vel3, c->in_vel.p, 3 * n * sizeof(double), k, c->stream));
  if(velpred3)
    {
      if(!c->run_ready) return c->fail("gl_get_run_state: VelPred has a row per particle only after gl_run_begin");
      GL_CUDA(c, cudaMemcpyAsync(velpred3, c->in_velpred.p, 3 * n * sizeof(double), k, c->stream));
    }
  if(hsml) GL_CUDA(c, cudaMemcpyAsync(hsml, c->in_hsml.p, n * sizeof(double), k, c->stream));
  if(entropy) GL_CUDA(c, cudaMemcpyAsync(entropy, c->in_entropy.p, n * sizeof(double), k, c->stream));
  if(dtentropy) GL_CUDA(c, cudaMemcpyAsync(dtentropy, c->in_dtentropy.p, n * sizeof(double), k, c->stream));
  if(timebin) GL_CUDA(c, cudaMemcpyAsync(timebin, c->in_timebin.p, n * sizeof(int), k, c->stream));
  if(ti_begstep) GL_CUDA(c, cudaMemcpyAsync(ti_begstep, c->in_tibegstep.p, n * sizeof(int), k, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// ---------------------------------------------------------------- the loop on the reference's own P[] / SphP[] ----
// What a reference-side binding calls instead of run.c's loop body for `nsteps` iterations: the structs are read ONCE,
// the loop runs on the device, the structs are written ONCE - where the per-step drop-in call moves them through the
// host's memory twice per step.  Not a pipeline like gl_gadget_compute_accelerations: one pass of the host pool each way.
#include "gl_host.h"
#include <chrono>

namespace
{
template <class T> inline T &fld(char *base, long long off) { return *reinterpret_cast<T *>(base + off); }
template <class T> inline const T &fld(const char *base, long long off) { return *reinterpret_cast<const T *>(base + off); }
}

extern "C" int gl_gadget_run_steps(gl_ctx *c, void *Pv, void *SphPv, long long NumPart, long long N_gas, gl_timestep_params *tp, int nsteps,
				   gl_run_info *info)
{
  cudaSetDevice(c->device);
  if(!c->have_lay) return c->fail("gl_gadget_run_steps: gl_gadget_bind not called");
  const gl_gadget_layout L = c->lay;
  const long long n = NumPart, ng = std::min(N_gas, NumPart);
  if(n <= 0) return c->fail("gl_gadget_run_steps: NumPart <= 0");
  if(L.p_Ti_current < 0) return c->fail("gl_gadget_run_steps: the layout has no P.Ti_current");
  char *P = (char *) Pv, *S = (char *) SphPv;
  // ---- gather: everything run()'s loop carries from one step to the next ----
  // staging in page-locked memory, kept by the context: pageable vectors cost more in first-touch faults and bounce
  // copies than the loop itself (measured: 217 ms per call for 1.6 M particles, 1.6 M x 500 B each way)
  const size_t need = (size_t) n * (29 * sizeof(double) + 3 * sizeof(int) + sizeof(float)) + 64;
  if(c->run_pin_cap < need)
    {
      if(c->run_pin) cudaFreeHost(c->run_pin);
      c->run_pin = nullptr;
      c->run_pin_cap = 0;
      GL_CUDA(c, cudaMallocHost(&c->run_pin, need + need / 8));
      c->run_pin_cap = need + need / 8;
    }
  double *fbase = (double *) c->run_pin;
  int *qbase = (int *) (fbase + (size_t) n * 29);
  float *gcbase = (float *) (qbase + (size_t) n * 3);
  double *pos = fbase, *vel = pos + 3 * n, *mass = vel + 3 * n, *hsml = mass + n, *oldacc = hsml + n, *entropy = oldacc + n,
    *dtentropy = entropy + n, *velpred = dtentropy + n, *gacc = velpred + 3 * n, *hacc = gacc + 3 * n, *maxsig = hacc + 3 * n,
    *density = maxsig + n, *pressure = density + n, *dhsml = pressure + n, *divvel = dhsml + n, *curlvel = divvel + n,
    *numngb = curlvel + n, *gcost = numngb + n;
  int *type = qbase, *timebin = type + n, *tibeg = timebin + n;
  std::atomic<long long> behind{0}, badgas{0};
  const int ti = tp->Ti_Current;
  static const bool dbg = getenv("GL_DEBUG_E2E") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double tl = now();
  auto stamp = [&](const char *what) {
    if(!dbg) return;
    const double t = now();
    fprintf(stderr, "[gl run_steps] %-10s %8.2f ms\n", what, t - tl);
    tl = t;
  };
  HostPool::get().run(n, 8192, [&](long long, long long a, long long b) {
    long long nb = 0, bg = 0;
    for(long long i = a; i < b; i++)
      {
	const char *pi = P + i * L.p_stride;
	const int ty = fld<short>(pi, L.p_Type);
	for(int k = 0; k < 3; k++)
	  {
	    pos[3 * i + k] = (&fld<double>(pi, L.p_Pos))[k];
	    vel[3 * i + k] = (&fld<double>(pi, L.p_Vel))[k];
	    gacc[3 * i + k] = (&fld<double>(pi, L.p_GravAccel))[k];
	  }
	mass[i] = fld<double>(pi, L.p_Mass);
	hsml[i] = fld<double>(pi, L.p_Hsml);
	oldacc[i] = fld<double>(pi, L.p_OldAcc);
	type[i] = ty;
	timebin[i] = fld<short>(pi, L.p_TimeBin);
	tibeg[i] = fld<int>(pi, L.p_Ti_begstep);
	if(ty != 3 && fld<int>(pi, L.p_Ti_current) != ti) nb++;
	if((ty == 0) != (i < ng)) bg++;
	if(i < ng)
	  {
	    const char *si = S + i * L.sph_stride;
	    for(int k = 0; k < 3; k++)
	      {
		velpred[3 * i + k] = (&fld<double>(si, L.s_VelPred))[k];
		hacc[3 * i + k] = (&fld<double>(si, L.s_HydroAccel))[k];
	      }
	    entropy[i] = fld<double>(si, L.s_Entropy);
	    dtentropy[i] = fld<double>(si, L.s_DtEntropy);
	    maxsig[i] = fld<double>(si, L.s_MaxSignalVel);
	    density[i] = fld<double>(si, L.s_Density);
	    pressure[i] = fld<double>(si, L.s_Pressure);
	    dhsml[i] = fld<double>(si, L.s_DhsmlDensityFactor);
	    divvel[i] = fld<double>(si, L.s_DivVel);
	    curlvel[i] = fld<double>(si, L.s_CurlVel);
	  }
	else
	  {
	    for(int k = 0; k < 3; k++) velpred[3 * i + k] = hacc[3 * i + k] = 0;
	    entropy[i] = dtentropy[i] = maxsig[i] = density[i] = pressure[i] = dhsml[i] = divvel[i] = curlvel[i] = 0;
	  }
      }
    if(nb) behind.fetch_add(nb);
    if(bg) badgas.fetch_add(bg);
  });
  stamp("gather");
  if(badgas.load()) return c->fail("gl_gadget_run_steps: the gas particles (Type 0) must be exactly P[0..N_gas)");
  if(behind.load())
    return c->fail("gl_gadget_run_steps: " + std::to_string(behind.load()) + " particles are not drifted to Ti_Current (P.Ti_current != All.Ti_Current): "
		   "call move_particles() / gl_gadget_drift first");
  if(gl_load_particles(c, n, pos, vel, mass, type, hsml, oldacc, timebin, tibeg, entropy, dtentropy, 0)) return 1;
  c->par.Ti_Current = ti;
  if(ng > 0)
    {
      if(gl_set_velpred(c, velpred, ng, 0)) return 1;
      if(gl_set_sph_state(c, density, pressure, dhsml, divvel, curlvel, 0)) return 1;
    }
  if(gl_set_accelerations(c, gacc, ng > 0 ? hacc : nullptr, ng > 0 ? maxsig : nullptr, 0)) return 1;
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  if(gl_run_begin(c, ti)) return 1;
  stamp("upload");
  // ---- the loop ----
  gl_run_info I{};
  if(gl_run_steps(c, tp, nsteps, &I)) return 1;
  if(info) *info = I;
  stamp("loop");
  // ---- scatter: the state and the results of the last force computation of every particle ----
  float *gc = gcbase;
  if(gl_get_run_state(c, pos, vel, velpred, hsml, entropy, dtentropy, timebin, tibeg, 0)) return 1;
  const long long keep_first = c->res_first, keep_count = c->res_count;
  c->res_first = 0;
  c->res_count = n;		// the own particles (the arrays may be longer in multi-rank mode)
  int rc = gl_get_gravity(c, gacc, oldacc, gc, 0);
  if(!rc && ng > 0) rc = gl_get_density(c, nullptr, numngb, density, dhsml, divvel, curlvel, pressure, nullptr, nullptr, nullptr, 0);
  if(!rc && ng > 0) rc = gl_get_hydro(c, hacc, nullptr, maxsig, 0);
  c->res_first = keep_first;
  c->res_count = keep_count;
  if(rc) return 1;
  stamp("download");
  (void) gcost;
  const int ti_end = tp->Ti_Current;
  const bool stepped = I.steps_done > 0;
  HostPool::get().run(n, 8192, [&](long long, long long a, long long b) {
    for(long long i = a; i < b; i++)
      {
	char *pi = P + i * L.p_stride;
	const int ty = type[i];
	for(int k = 0; k < 3; k++)
	  {
	    (&fld<double>(pi, L.p_Pos))[k] = pos[3 * i + k];
	    (&fld<double>(pi, L.p_Vel))[k] = vel[3 * i + k];
	    (&fld<double>(pi, L.p_GravAccel))[k] = gacc[3 * i + k];
	  }
	fld<double>(pi, L.p_Hsml) = hsml[i];
	fld<double>(pi, L.p_OldAcc) = oldacc[i];
	if(stepped) fld<float>(pi, L.p_GravCost) = gc[i];
	fld<short>(pi, L.p_TimeBin) = (short) timebin[i];
	fld<int>(pi, L.p_Ti_begstep) = tibeg[i];
	if(ty != 3) fld<int>(pi, L.p_Ti_current) = ti_end;	// predict.c:134: virtual particles are never drifted
	if(i < ng)
	  {
	    char *si = S + i * L.sph_stride;
	    if(stepped) fld<double>(pi, L.p_NumNgb) = numngb[i];
	    for(int k = 0; k < 3; k++)
	      {
		(&fld<double>(si, L.s_VelPred))[k] = velpred[3 * i + k];
		(&fld<double>(si, L.s_HydroAccel))[k] = hacc[3 * i + k];
	      }
	    fld<double>(si, L.s_Entropy) = entropy[i];
	    fld<double>(si, L.s_DtEntropy) = dtentropy[i];
	    fld<double>(si, L.s_MaxSignalVel) = maxsig[i];
	    fld<double>(si, L.s_Density) = density[i];
	    fld<double>(si, L.s_Pressure) = pressure[i];
	    fld<double>(si, L.s_DhsmlDensityFactor) = dhsml[i];
	    fld<double>(si, L.s_DivVel) = divvel[i];
	    fld<double>(si, L.s_CurlVel) = curlvel[i];
	    if(L.s_DragAccel >= 0 && stepped)
	      for(int k = 0; k < 3; k++) (&fld<double>(si, L.s_DragAccel))[k] = 0;	// do_the_kick, timestep.c:503-505
	  }
      }
  });
  stamp("scatter");
  return 0;
}

// ---------------------------------------------------------------- domain_exchange for the resident state ----
// domain_Decomposition() ends with domain_exchange (domain.c:665): every particle struct travels to the rank its key now
// belongs to.  With the state resident in the library the same move happens between the GPUs: the new owners come from
// gl_domain_split_collective (equal-cost stretches of the global Peano-Hilbert order, domain.c:1075 / :378), every
// particle's state is packed into one 256-byte record, the records cross NVLink in ONE all-to-all, and the receiver
// stores them in Peano-Hilbert order of the current grid (so a warp's 32 targets are neighbours again and the stretch
// key hulls are tight).  Results do not depend on who owns what: this only restores the balance and the
// pruning of the locally-essential exchange after the particles have moved.
namespace
{
#define XR 32			// doubles per record
// [0..2] pos [3..5] vel [6..8] velpred [9..11] gravaccel [12..14] hydroaccel [15] mass [16] hsml [17] oldacc [18] entropy
// [19] dtentropy [20] maxsignalvel [21] density [22] pressure [23] dhsml [24] divvel [25] curlvel [26] gravcost [27] numngb
// [28] key (bits) [29] type | timebin [30] ti_begstep | id [31] -
struct XArrays
{
  double *pos, *vel, *velpred, *acc, *hacc, *mass, *hsml, *oldacc, *entropy, *dtentropy, *maxsig, *density, *pressure, *dhsml, *divvel,
    *curlvel, *numngb;
  float *gravcost;
  int *type, *timebin, *tibeg, *id;
};

__global__ void k_x_cost(long long n, const float *__restrict__ gravcost, const int *__restrict__ timebin, double *__restrict__ cost)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= n) return;
  const int tb = timebin[i];
  cost[i] = (1.0 + (double) gravcost[i]) / (double) (1ll << (tb > 0 && tb < 40 ? tb : 0));	// domain.c:378-384
}

__global__ void k_x_pack(long long n, const int *__restrict__ order, const unsigned long long *__restrict__ key, XArrays A, int sph,
			 double *__restrict__ rec)
{
  const long long k = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(k >= n) return;
  const long long i = order[k];
  double *r = rec + XR * k;
  for(int j = 0; j < 3; j++)
    {
      r[j] = A.pos[3 * i + j];
      r[3 + j] = A.vel[3 * i + j];
      r[6 + j] = A.velpred[3 * i + j];
      r[9 + j] = A.acc[3 * i + j];
      r[12 + j] = sph ? A.hacc[3 * i + j] : 0.0;
    }
  r[15] = A.mass[i]; r[16] = A.hsml[i]; r[17] = A.oldacc[i]; r[18] = A.entropy[i]; r[19] = A.dtentropy[i];
  r[20] = sph ? A.maxsig[i] : 0.0; r[21] = sph ? A.density[i] : 0.0; r[22] = sph ? A.pressure[i] : 0.0; r[23] = sph ? A.dhsml[i] : 0.0;
  r[24] = sph ? A.divvel[i] : 0.0; r[25] = sph ? A.curlvel[i] : 0.0;
  r[26] = (double) A.gravcost[i];
  r[27] = sph ? A.numngb[i] : 0.0;
  r[28] = __longlong_as_double((long long) key[i]);
  r[29] = __hiloint2double(A.type[i], A.timebin[i]);
  r[30] = __hiloint2double(A.tibeg[i], A.id[i]);
  r[31] = 0;
}

__global__ void k_x_keys(long long n, const double *__restrict__ rec, unsigned long long *__restrict__ key, int *__restrict__ idx)
{
  const long long k = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(k >= n) return;
  key[k] = (unsigned long long) __double_as_longlong(rec[XR * k + 28]);
  idx[k] = (int) k;
}

__global__ void k_x_unpack(long long n, const int *__restrict__ order, const double *__restrict__ rec, XArrays A, int sph)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i >= n) return;
  const double *r = rec + XR * (long long) order[i];
  for(int j = 0; j < 3; j++)
    {
      A.pos[3 * i + j] = r[j];
      A.vel[3 * i + j] = r[3 + j];
      A.velpred[3 * i + j] = r[6 + j];
      A.acc[3 * i + j] = r[9 + j];
      if(sph) A.hacc[3 * i + j] = r[12 + j];
    }
  A.mass[i] = r[15]; A.hsml[i] = r[16]; A.oldacc[i] = r[17]; A.entropy[i] = r[18]; A.dtentropy[i] = r[19];
  if(sph)
    {
      A.maxsig[i] = r[20]; A.density[i] = r[21]; A.pressure[i] = r[22]; A.dhsml[i] = r[23]; A.divvel[i] = r[24]; A.curlvel[i] = r[25];
      A.numngb[i] = r[27];
    }
  A.gravcost[i] = (float) r[26];
  A.type[i] = __double2hiint(r[29]); A.timebin[i] = __double2loint(r[29]);
  A.tibeg[i] = __double2hiint(r[30]); A.id[i] = __double2loint(r[30]);
}

__global__ void k_x_default_ids(long long n, int base, int *__restrict__ id)
{
  const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if(i < n) id[i] = base + (int) i;
}

XArrays x_arrays(gl_ctx *c)
{
  return XArrays{ c->in_pos.p, c->in_vel.p, c->in_velpred.p, c->o_acc.p, c->o_hacc.p, c->in_mass.p, c->in_hsml.p, c->in_oldacc.p, c->in_entropy.p,
		  c->in_dtentropy.p, c->o_maxsig.p, c->o_density.p, c->o_pressure.p, c->o_dhsml.p, c->o_divvel.p, c->o_curlvel.p, c->o_numngb.p,
		  c->o_gravcost.p, c->in_type.p, c->in_timebin.p, c->in_tibegstep.p, c->run_id.p };
}
}				// namespace

extern "C" int gl_run_set_ids(gl_ctx *c, const int *ids, int where)
{
  cudaSetDevice(c->device);
  const long long n = c->n_own;
  if(c->n <= 0 || n <= 0 || !ids) return c->fail("gl_run_set_ids: no particles / no ids");
  GL_CUDA(c, c->run_id.ensure((size_t) n));
  GL_CUDA(c, cudaMemcpyAsync(c->run_id.p, ids, n * sizeof(int), where ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  c->run_have_id = true;
  return 0;
}

extern "C" int gl_get_run_ids(gl_ctx *c, int *ids, int where)
{
  cudaSetDevice(c->device);
  if(!c->run_have_id) return c->fail("gl_get_run_ids: no ids in the context (gl_run_set_ids, or a gl_domain_exchange that numbered the particles)");
  GL_CUDA(c, cudaMemcpyAsync(ids, c->run_id.p, c->n_own * sizeof(int), where ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int gl_domain_exchange(gl_ctx *c, int by_cost, long long *n_new, long long *n_sent)
{
  cudaSetDevice(c->device);
  if(!c->run_ready) return c->fail("gl_domain_exchange: the resident state is not set up (gl_run_begin)");
  gl_comm *cm = c->comm;
  if(!cm || cm->world < 2)
    {
      if(n_new) *n_new = c->n_own;
      if(n_sent) *n_sent = 0;
      return 0;
    }
  const int W = cm->world, me = cm->rank;
  const long long n = c->n_own;
  const bool sph = c->o_density.cap >= (size_t) n && c->o_hacc.cap >= 3 * (size_t) n;
  if(c->o_gravcost.cap < (size_t) n)
    {
      GL_CUDA(c, c->o_gravcost.ensure((size_t) n));
      GL_CUDA(c, cudaMemsetAsync(c->o_gravcost.p, 0, n * sizeof(float), c->stream));
    }
  if(sph) GL_CUDA(c, c->o_numngb.ensure_keep((size_t) n));
  // ---- ids: whoever holds a particle afterwards can say which one it is ----
  if(!c->run_have_id)
    {
      long long *cnt = (long long *) c->run_hist.p;	// 32 (W + 1) ints: room for W + 1 long longs
      GL_CUDA(c, cudaMemcpyAsync(cnt, &n, sizeof(long long), cudaMemcpyHostToDevice, c->stream));
      if(cm->allgather(cnt, cnt + 1, sizeof(long long), c->stream)) return c->fail("gl_domain_exchange: all-gather failed: " + cm->err);
      std::vector<long long> hn((size_t) W);
      GL_CUDA(c, cudaMemcpyAsync(hn.data(), cnt + 1, W * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
      GL_CUDA(c, cudaStreamSynchronize(c->stream));
      long long base = 0;
      for(int r = 0; r < me; r++) base += hn[(size_t) r];
      GL_CUDA(c, c->run_id.ensure((size_t) n));
      k_x_default_ids<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, (int) base, c->run_id.p);
      GL_LAUNCH_COUNT(c);
      c->run_have_id = true;
    }
  // ---- who goes where: equal-cost stretches of the global Peano-Hilbert order ----
  DBuf<double> cost, rec, rec2;
  DBuf<int> dest, idx, order;
  DBuf<unsigned long long> ksorted, cntdev;
  struct Release
  {
    std::function<void()> f;
    ~Release() { f(); }
  } rel{ [&] { cost.release(); rec.release(); rec2.release(); dest.release(); idx.release(); order.release(); ksorted.release(); cntdev.release(); } };
  GL_CUDA(c, dest.ensure((size_t) n));
  if(by_cost)
    {
      GL_CUDA(c, cost.ensure((size_t) n));
      k_x_cost<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, c->o_gravcost.p, c->in_timebin.p, cost.p);
      GL_LAUNCH_COUNT(c);
    }
  std::vector<long long> to_go((size_t) W);
  if(gl_domain_split_collective(c, by_cost ? cost.p : nullptr, 1, dest.p, 1, to_go.data(), nullptr)) return 1;
  // the owner is a monotone function of the key: sorted by key, the particles of every destination are one stretch
  GL_CUDA(c, idx.ensure((size_t) n)); GL_CUDA(c, order.ensure((size_t) n)); GL_CUDA(c, ksorted.ensure((size_t) n));
  GL_CUDA(c, rec.ensure((size_t) n * XR));
  {
    // iota
    k_x_default_ids<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, 0, idx.p);
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->key_ph.p, ksorted.p, idx.p, order.p, (int) n, 0, 63, c->stream);
    GL_CUDA(c, c->cub_tmp.ensure(tmp));
    GL_CUDA(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, c->key_ph.p, ksorted.p, idx.p, order.p, (int) n, 0, 63, c->stream));
    c->kernel_launches += 6;
  }
  k_x_pack<<<gl_grid(n, 256), 256, 0, c->stream>>>(n, order.p, c->key_ph.p, x_arrays(c), sph ? 1 : 0, rec.p);
  GL_LAUNCH_COUNT(c);
  // ---- counts: my row of toGo, everybody's rows all-gathered (domain.c:1274 + the MPI_Allgather of :700) ----
  long long sent = 0;
  for(int r = 0; r < W; r++) sent += to_go[(size_t) r];
  to_go[(size_t) me] = n - sent;
  GL_CUDA(c, cntdev.ensure((size_t) W * (W + 1)));
  GL_CUDA(c, cudaMemcpyAsync(cntdev.p, to_go.data(), W * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
  if(cm->allgather(cntdev.p, cntdev.p + W, W * sizeof(long long), c->stream)) return c->fail("gl_domain_exchange: all-gather failed: " + cm->err);
  std::vector<long long> all((size_t) W * W);
  GL_CUDA(c, cudaMemcpyAsync(all.data(), cntdev.p + W, (size_t) W * W * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  std::vector<size_t> soff((size_t) W), scnt((size_t) W), roff((size_t) W), rcnt((size_t) W);
  long long nn = 0;
  size_t so = 0;
  for(int r = 0; r < W; r++)
    {
      scnt[(size_t) r] = (size_t) to_go[(size_t) r] * XR * sizeof(double);
      soff[(size_t) r] = so;
      so += scnt[(size_t) r];
      const long long from_r = all[(size_t) r * W + me];
      rcnt[(size_t) r] = (size_t) from_r * XR * sizeof(double);
      roff[(size_t) r] = (size_t) nn * XR * sizeof(double);
      nn += from_r;
    }
  if(nn <= 0) return c->fail("gl_domain_exchange: this rank would be left without particles");
  if(nn >= (1ll << 27)) return c->fail("gl_domain_exchange: more than 2^27 particles on one rank");
  GL_CUDA(c, rec2.ensure((size_t) nn * XR));
  if(cm->alltoallv(rec.p, soff.data(), scnt.data(), rec2.p, roff.data(), rcnt.data(), c->stream))
    return c->fail("gl_domain_exchange: all-to-all failed: " + cm->err);
  // ---- store in Peano-Hilbert order of the current grid ----
  GL_CUDA(c, idx.ensure((size_t) nn)); GL_CUDA(c, order.ensure((size_t) nn)); GL_CUDA(c, ksorted.ensure((size_t) nn));
  GL_CUDA(c, c->key_ph.ensure((size_t) nn));
  k_x_keys<<<gl_grid(nn, 256), 256, 0, c->stream>>>(nn, rec2.p, c->key_ph.p, idx.p);
  GL_LAUNCH_COUNT(c);
  {
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->key_ph.p, ksorted.p, idx.p, order.p, (int) nn, 0, 63, c->stream);
    GL_CUDA(c, c->cub_tmp.ensure(tmp));
    GL_CUDA(c, cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, c->key_ph.p, ksorted.p, idx.p, order.p, (int) nn, 0, 63, c->stream));
    c->kernel_launches += 6;
  }
  const size_t N = (size_t) nn;
  GL_CUDA(c, c->in_pos.ensure(3 * N)); GL_CUDA(c, c->in_vel.ensure(3 * N)); GL_CUDA(c, c->in_velpred.ensure(3 * N));
  GL_CUDA(c, c->o_acc.ensure(3 * N)); GL_CUDA(c, c->in_mass.ensure(N)); GL_CUDA(c, c->in_hsml.ensure(N)); GL_CUDA(c, c->in_oldacc.ensure(N));
  GL_CUDA(c, c->in_entropy.ensure(N)); GL_CUDA(c, c->in_dtentropy.ensure(N)); GL_CUDA(c, c->o_gravcost.ensure(N));
  GL_CUDA(c, c->in_type.ensure(N)); GL_CUDA(c, c->in_timebin.ensure(N)); GL_CUDA(c, c->in_tibegstep.ensure(N)); GL_CUDA(c, c->run_id.ensure(N));
  GL_CUDA(c, c->o_oldacc.ensure(N));
  if(sph)
    {
      GL_CUDA(c, c->o_hacc.ensure(3 * N)); GL_CUDA(c, c->o_maxsig.ensure(N)); GL_CUDA(c, c->o_density.ensure(N)); GL_CUDA(c, c->o_pressure.ensure(N));
      GL_CUDA(c, c->o_dhsml.ensure(N)); GL_CUDA(c, c->o_divvel.ensure(N)); GL_CUDA(c, c->o_curlvel.ensure(N)); GL_CUDA(c, c->o_numngb.ensure(N));
      GL_CUDA(c, c->o_dtentr.ensure(N));
      GL_CUDA(c, c->o_denspairs.ensure(N));
      GL_CUDA(c, cudaMemsetAsync(c->o_denspairs.p, 0, c->o_denspairs.cap * sizeof(int), c->stream));	// no history for the new set
    }
  k_x_unpack<<<gl_grid(nn, 256), 256, 0, c->stream>>>(nn, order.p, rec2.p, x_arrays(c), sph ? 1 : 0);
  GL_LAUNCH_COUNT(c);
  GL_CUDA(c, cudaMemcpyAsync(c->o_oldacc.p, c->in_oldacc.p, nn * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  GL_CUDA(c, cudaStreamSynchronize(c->stream));
  GL_CUDA(c, cudaGetLastError());
  // ---- a new particle set on this rank ----
  c->n = nn;
  gl_fresh_particle_set(c);
  c->n_velpred = nn;
  c->have_velpred = true;
  c->n_tgt_in = -1;
  c->tgt_host.clear();
  c->tgt_ordered = false;
  c->tgt_dirty = true;
  c->res_first = 0;
  c->res_count = -1;
  c->tree_valid = false;
  c->tree_flag = true;
  if(!c->domain_imposed) c->have_domain = false;
  if(n_new) *n_new = nn;
  if(n_sent) *n_sent = sent;
  return 0;
}
