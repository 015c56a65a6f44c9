// gl_timestep.cu - the per-particle part of advance_and_find_timesteps() (timestep.c:29) on the reference's own
// P[] / SphP[]: get_timestep (timestep.c:607) -> power-of-two bin -> do_the_kick (timestep.c:364), for the particles of
// the active list, on all host cores.  Host code only: no context, no GPU.
//
// The statements and their order are the reference's under the shipped -D list (non-comoving: fac1 = fac2 = fac3 =
// hubble_a = a3inv = atime = 1, dt_displacement = All.MaxSizeTimestep, timestep.c:1131); the decisions are FP64
// comparisons, so the same libm gives the same bins and the same kicked state bit for bit (tests/test_timestep.py).
// Every particle only touches its own structs, which is what makes the loop parallel; the reference's serial part -
// relinking the time-bin lists (timestep.c:170-233) - is left to the caller (see include/gl_b200.h).
#include "gl_internal.h"
#include "gl_host.h"
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <mutex>

namespace
{
const int TIMEBASE = 1 << 29;	// allvars.h:39-41 (TIMEBINS 29)

struct Fail
{
  std::mutex m;
  std::atomic<int> code{0};
  char *err;
  int errlen;
  void set(int c, const char *fmt, ...)
  {
    std::lock_guard<std::mutex> lk(m);
    if(code.load()) return;	// keep the first message
    if(err && errlen > 0)
      {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(err, (size_t) errlen, fmt, ap);
	va_end(ap);
      }
    code.store(c);
  }
};

// get_timestep(p, &aphys, 0), timestep.c:607-1122; returns the step on the integer time line, < 0 after a failure
int get_timestep(const gl_gadget_layout &L, const gl_timestep_params &T, const char *pi, const char *si, long long p, Fail &f)
{
  const int type = *(const short *) (pi + L.p_Type);
  const double *ga = (const double *) (pi + L.p_GravAccel);
  const double mass = *(const double *) (pi + L.p_Mass), hsml = *(const double *) (pi + L.p_Hsml);
  double ax = ga[0], ay = ga[1], az = ga[2];	// fac1 = 1
  if(type == 0)
    {
      const double *ha = (const double *) (si + L.s_HydroAccel);
      ax += ha[0];
      ay += ha[1];
      az += ha[2];
      if(L.s_DragAccel >= 0)	// -DDUST, timestep.c:645-649
	{
	  const double *da = (const double *) (si + L.s_DragAccel);
	  ax += da[0];
	  ay += da[1];
	  az += da[2];
	}
    }
  double ac = sqrt(ax * ax + ay * ay + az * az);
  if(ac == 0) ac = 1.0e-30;
  // TypeOfTimestepCriterion 0, timestep.c:676-708
  double dt = sqrt(2 * T.ErrTolIntAccuracy * 1.0 * T.SofteningTable[type] / ac);
  if(T.p_DeltaDustMomentum >= 0 && type == 0 && mass > 0)
    {
      const double *ddm = (const double *) (pi + T.p_DeltaDustMomentum);
      ax += ddm[0] / mass / dt;
      ay += ddm[1] / mass / dt;
      az += ddm[2] / mass / dt;
      ac = sqrt(ax * ax + ay * ay + az * az);
      if(ac > 0)
	{
	  const double dt_new = sqrt(2 * T.ErrTolIntAccuracy * 1.0 * T.SofteningTable[type] / ac);
	  if(dt_new < dt) dt = dt_new;
	}
    }
  if(type == 2) dt = dt / 2;
  if(type == 0) dt = sqrt(2 * T.ErrTolIntAccuracy * 1.0 * hsml / 2.8 / ac);
  if(type == 0)
    {
      // Courant condition, timestep.c:853-866
      const double dt_courant = 2 * T.CourantFac * hsml / *(const double *) (si + L.s_MaxSignalVel);
      if(dt_courant < dt) dt = dt_courant;
    }
  if(type == 3)
    {
      // timestep.c:889-900; the free-fall branch reads `rho`, which the reference never sets
      double dt_abs = 0.03 * T.OuterBoundary / 2.9979e10 * T.UnitVelocity_in_cm_per_s * T.FeedBackVelocity;
      const double dt_ff = 1.;
      if(T.p_NewDensity >= 0 && *(const double *) (pi + T.p_NewDensity) > 0.)
	{
	  f.set(894, "get_timestep: Type-3 particle %lld has NewDensity > 0; the reference's criterion for it reads an "
		"uninitialised variable (timestep.c:894)", p);
	  return -1;
	}
      if(dt_abs > dt_ff) dt_abs = dt_ff;
      if(dt > dt_abs) dt = dt_abs;
    }
  if(type == 5)
    {
      // sink particles: a fraction of the orbital time at the accretion radius, timestep.c:909-950
      double dt_accr = 0.03 * (T.OuterBoundary / 100.), dt_a = 1.e10;
      if(mass >= 0.45 * T.SMBHmass && T.InnerBoundary > 0) dt_a = T.AccDtBlackHole * pow((T.InnerBoundary + 0.5 * hsml), 1.5) / pow(mass, 0.5);
      if(mass < 0.45 * T.SMBHmass && T.SinkBoundary > 0) dt_a = T.AccDtBlackHole * pow((T.SinkBoundary + 0.5 * hsml), 1.5) / pow(mass, 0.5);
      if(dt_accr > dt_a) dt_accr = dt_a;
      if(dt_accr < dt) dt = dt_accr;
    }
  // timestep.c:1042-1122 (hubble_a = 1)
  if(dt >= T.MaxSizeTimestep) dt = T.MaxSizeTimestep;
  if(dt >= T.MaxSizeTimestep) dt = T.MaxSizeTimestep;	// dt_displacement = All.MaxSizeTimestep without comoving integration
  if(dt < T.MinSizeTimestep)
    {
      f.set(888, "warning: Timestep wants to be below the limit `MinSizeTimestep' (particle %lld, dt=%g)", p, dt);
      return -1;
    }
  const int ti_step = (int) (dt / T.Timebase_interval);
  if(!(ti_step > 0 && ti_step < TIMEBASE))
    {
      f.set(818, "Error: A timestep of size zero was assigned on the integer timeline! (particle %lld, dt=%g, ti_step=%d)", p, dt, ti_step);
      return -1;
    }
  return ti_step;
}

// do_the_kick(i, tstart, tend, tcurrent), timestep.c:364-601
void do_the_kick(const gl_gadget_layout &L, const gl_timestep_params &T, char *pi, char *si, int tstart, int tend, int tcurrent)
{
  const int type = *(const short *) (pi + L.p_Type);
  if(type == 3) return;	// -DVIRTUAL, timestep.c:379
  const double dt_entr0 = (tend - tstart) * T.Timebase_interval, dt_kick = dt_entr0;
  const double dt_kick2 = (tend - tcurrent) * T.Timebase_interval;
  double *vel = (double *) (pi + L.p_Vel);
  const double *ga = (const double *) (pi + L.p_GravAccel);
  double dv[3];
  for(int j = 0; j < 3; j++)
    {
      dv[j] = ga[j] * dt_kick;
      if(type == 2) dv[j] = 0.0;	// dust is moved by the drag module, not kicked here (timestep.c:478-485)
      vel[j] += dv[j];
    }
  if(type == 0)
    {
      const double *ha = (const double *) (si + L.s_HydroAccel);
      double *vp = (double *) (si + L.s_VelPred);
      for(int j = 0; j < 3; j++)
	{
	  vel[j] += ha[j] * dt_kick;
	  vp[j] = vel[j] - dt_kick2 * ga[j] - dt_kick2 * ha[j];
	  if(L.s_DragAccel >= 0) ((double *) (si + L.s_DragAccel))[j] = 0.;
	}
      double *A = (double *) (si + L.s_Entropy), *dA = (double *) (si + L.s_DtEntropy);
      // "In case of cooling, we prevent that the entropy (and hence temperature) decreases by more than a factor 0.5"
      if(*dA * dt_entr0 > -0.5 * *A)
	*A += *dA * dt_entr0;
      else
	*A *= 0.5;
      if(T.MinEgySpec)
	{
	  const double minentropy = T.MinEgySpec * T.GammaMinus1 / pow(*(const double *) (si + L.s_Density), T.GammaMinus1);
	  if(*A < minentropy)
	    {
	      *A = minentropy;
	      *dA = 0;
	    }
	}
      // the same limit for the predicted entropy at the middle of the NEW step (timestep.c:577-583)
      const int tb = *(const short *) (pi + L.p_TimeBin);
      const double dt_entr = (tb ? (1 << tb) : 0) / 2 * T.Timebase_interval;
      if(*A + *dA * dt_entr < 0.5 * *A) *dA = -0.5 * *A / dt_entr;
    }
}
}				// namespace

extern "C" int gl_gadget_advance_and_find_timesteps(const gl_gadget_layout *lay, const gl_timestep_params *par, void *Pv, void *SphPv,
						    const int *active, long long nactive, int *binold_out, char *err, int errlen)
{
  const gl_gadget_layout L = *lay;
  const gl_timestep_params T = *par;
  char *P = (char *) Pv, *S = (char *) SphPv;
  Fail fail;
  fail.err = err;
  fail.errlen = errlen;
  if(err && errlen > 0) err[0] = 0;
  HostPool::get().run(nactive, 4096, [&](long long, long long a, long long b) {
    for(long long k = a; k < b; k++)
      {
	const long long i = active[k];
	char *pi = P + i * L.p_stride, *si = S + i * L.sph_stride;	// si is only dereferenced for gas
	short *timebin = (short *) (pi + L.p_TimeBin);
	const int binold = *timebin;
	if(binold_out) binold_out[k] = binold;
	int ti_step = get_timestep(L, T, pi, si, i, fail);
	if(ti_step < 0) continue;
	int ti_min = TIMEBASE;	// make it a power 2 subdivision, timestep.c:141-144
	while(ti_min > ti_step) ti_min >>= 1;
	ti_step = ti_min;
	if(ti_step == 1)	// get_timestep_bin, timestep.c:1226-1246
	  {
	    fail.set(112313, "time-step of integer size 1 not allowed (particle %lld)", i);
	    continue;
	  }
	int bin = -1;
	for(int t = ti_step; t; t >>= 1) bin++;
	if(ti_step == 0) bin = 0;
	if(bin > binold && T.TimeBinActive[bin] == 0)	// timestep wants to increase: leave at old step if not synchronized
	  {
	    bin = binold;
	    ti_step = bin ? (1 << bin) : 0;
	  }
	if(T.Ti_Current >= TIMEBASE)	// we here finish the last timestep
	  {
	    ti_step = 0;
	    bin = 0;
	  }
	if((TIMEBASE - T.Ti_Current) < ti_step)
	  {
	    fail.set(888, "particle %lld would step beyond the end of the time line", i);
	    continue;
	  }
	if(bin != binold) *timebin = (short) bin;
	const int ti_step_old = binold ? (1 << binold) : 0;
	int *begstep = (int *) (pi + L.p_Ti_begstep);
	const int tstart = *begstep + ti_step_old / 2;	// midpoint of old step
	const int tend = *begstep + ti_step_old + ti_step / 2;	// midpoint of new step
	*begstep += ti_step_old;
	do_the_kick(L, T, pi, si, tstart, tend, *begstep);
      }
  });
  return fail.code.load() ? 1 : 0;
}
