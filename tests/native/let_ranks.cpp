// let_ranks.cpp - the multi-rank force path driven from C++ through the C ABI only (include/gl_b200.h): two contexts
// of this process, one host thread each, as two "MPI ranks" (gl_comm_init_local), against one context that holds all
// particles.  What a reference-side binding does at NTask = 2, without Python in between.  TEST CODE.
//   usage: let_ranks <libgl_b200.so> [n]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <thread>
#include <dlfcn.h>
#include "../../include/gl_b200.h"

#define SYM(name) decltype(&name) p_##name = (decltype(&name)) dlsym(h, #name); if(!p_##name) { fprintf(stderr, "missing %s\n", #name); return 2; }

struct Set
{
  std::vector<double> pos, vel, mass, hsml, oldacc, entropy, dtentropy;
  std::vector<int> type, timebin, tibeg;
  void resize(size_t n)
  {
    pos.resize(3 * n); vel.resize(3 * n); mass.resize(n); hsml.resize(n); oldacc.assign(n, 0.0); entropy.resize(n);
    dtentropy.assign(n, 0.0); type.assign(n, 0); timebin.assign(n, 0); tibeg.assign(n, 0);
  }
};

int main(int argc, char **argv)
{
  if(argc < 2) { fprintf(stderr, "usage: %s libgl_b200.so [n]\n", argv[0]); return 2; }
  void *h = dlopen(argv[1], RTLD_NOW);
  if(!h) { fprintf(stderr, "%s\n", dlerror()); return 2; }
  const long long n = argc > 2 ? atoll(argv[2]) : 40000;
  SYM(gl_create) SYM(gl_destroy) SYM(gl_last_error) SYM(gl_set_params) SYM(gl_load_particles) SYM(gl_peano_hilbert_order)
  SYM(gl_compute_accelerations) SYM(gl_get_gravity) SYM(gl_get_pair_counts) SYM(gl_comm_local_group_create) SYM(gl_comm_local_group_destroy)
  SYM(gl_comm_init_local) SYM(gl_get_let_stats)
  gl_params par;
  memset(&par, 0, sizeof(par));
  par.ErrTolTheta = 0.3; par.ErrTolForceAcc = 0.003; par.TypeOfOpeningCriterion = 1; par.DesNumNgb = 40; par.MaxNumNgbDeviation = 0.5;
  par.BlackHoleNgbFactor = 4; par.ArtBulkViscConst = 1; par.G = 1;
  const double soft[6] = { 1e-5, 5e-5, 5e-5, 1e-7, 0.01, 5e-5 };
  for(int k = 0; k < 6; k++) { par.SofteningTable[k] = soft[k]; par.ForceSoftening[k] = 2.8 * soft[k]; }
  par.MinGasHsml = 2.8 * soft[0]; par.Timebase_interval = 1.0 / (1 << 29); par.Ti_Current = 0;
  // a centrally condensed gas ball (uniform in r: rho ~ 1/r^2), velocities +-0.05
  Set all;
  all.resize(n);
  unsigned long long s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double) (s >> 11) / 9007199254740992.0; };
  for(long long i = 0; i < n; i++)
    {
      const double r = 0.02 + 0.98 * rnd(), ct = 2 * rnd() - 1, ph = 6.283185307179586 * rnd(), st = sqrt(1 - ct * ct);
      all.pos[3 * i] = r * st * cos(ph); all.pos[3 * i + 1] = r * st * sin(ph); all.pos[3 * i + 2] = r * ct;
      for(int k = 0; k < 3; k++) all.vel[3 * i + k] = 0.1 * rnd() - 0.05;
      all.mass[i] = 1.0 / n;
      all.hsml[i] = 1.5 * r * cbrt(40.0 / n);
      all.entropy[i] = 1.0;
    }
  gl_ctx *one = nullptr;
  if(p_gl_create(&one, 0, n)) { fprintf(stderr, "no GPU\n"); return 3; }
  auto chk = [&](gl_ctx *c, int rc, const char *what) { if(rc) { fprintf(stderr, "%s: %s\n", what, p_gl_last_error(c)); exit(4); } };
  chk(one, p_gl_set_params(one, &par), "set_params");
  auto load = [&](gl_ctx *c, const Set &S, long long m) {
    return p_gl_load_particles(c, m, S.pos.data(), S.vel.data(), S.mass.data(), S.type.data(), S.hsml.data(), S.oldacc.data(), S.timebin.data(),
			       S.tibeg.data(), S.entropy.data(), S.dtentropy.data(), 0);
  };
  chk(one, load(one, all, n), "load");
  std::vector<int> order(n);
  chk(one, p_gl_peano_hilbert_order(one, n, order.data(), 0), "peano_hilbert_order");
  chk(one, load(one, all, n), "load");
  chk(one, p_gl_compute_accelerations(one, 0), "compute_accelerations (one context)");
  std::vector<double> acc1(3 * n), old1(n);
  std::vector<float> cost1(n);
  std::vector<int> dp1(n), hp1(n);
  chk(one, p_gl_get_gravity(one, acc1.data(), old1.data(), cost1.data(), 0), "get_gravity");
  chk(one, p_gl_get_pair_counts(one, dp1.data(), hp1.data(), 0), "get_pair_counts");
  p_gl_destroy(one);

  const int W = 2;
  void *group = p_gl_comm_local_group_create(W);
  long long bad[W] = { 0, 0 }, imported[W] = { 0, 0 };
  auto rank_main = [&](int r) {
    const long long lo = r * (n / W), hi = (r == W - 1) ? n : (r + 1) * (n / W), m = hi - lo;
    Set S;
    S.resize(m);
    for(long long k = 0; k < m; k++)
      {
	const long long i = order[lo + k];
	for(int d = 0; d < 3; d++) { S.pos[3 * k + d] = all.pos[3 * i + d]; S.vel[3 * k + d] = all.vel[3 * i + d]; }
	S.mass[k] = all.mass[i]; S.hsml[k] = all.hsml[i]; S.entropy[k] = all.entropy[i];
      }
    gl_ctx *c = nullptr;
    if(p_gl_create(&c, 0, m)) exit(3);
    chk(c, p_gl_set_params(c, &par), "set_params");
    chk(c, p_gl_comm_init_local(c, r, group), "gl_comm_init_local");
    chk(c, load(c, S, m), "load own");
    chk(c, p_gl_compute_accelerations(c, 0), "compute_accelerations (rank)");
    std::vector<double> acc(3 * m), old(m);
    std::vector<float> cost(m);
    std::vector<int> dp(m), hp(m);
    chk(c, p_gl_get_gravity(c, acc.data(), old.data(), cost.data(), 0), "get_gravity");
    chk(c, p_gl_get_pair_counts(c, dp.data(), hp.data(), 0), "get_pair_counts");
    for(long long k = 0; k < m; k++)
      {
	const long long i = order[lo + k];
	if(cost[k] != cost1[i] || dp[k] != dp1[i] || hp[k] != hp1[i] || memcmp(&acc[3 * k], &acc1[3 * i], 24) != 0) bad[r]++;
      }
    gl_let_stats st;
    chk(c, p_gl_get_let_stats(c, &st), "let_stats");
    imported[r] = st.imported_particles;
    p_gl_destroy(c);
  };
  std::thread t0(rank_main, 0), t1(rank_main, 1);
  t0.join();
  t1.join();
  p_gl_comm_local_group_destroy(group);
  printf("let_ranks: n %lld, rank 0 imported %lld, rank 1 imported %lld, mismatching particles %lld\n", n, imported[0], imported[1], bad[0] + bad[1]);
  puts(bad[0] + bad[1] == 0 ? "OK" : "MISMATCH");
  return bad[0] + bad[1] == 0 ? 0 : 1;
}
