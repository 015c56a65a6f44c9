// host_scatter_probe.cpp - what do the drop-in's host-side pack / scatter loops cost on the GPU box's CPU?
// g++ -O2 -pthread -o host_scatter_probe host_scatter_probe.cpp
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <chrono>
#include <algorithm>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
template <class F> static void pfor(int nth, long long n, F f)
{
  std::vector<std::thread> th;
  long long chunk = (n + nth - 1) / nth;
  for(int t = 0; t < nth; t++)
    {
      long long a = t * chunk, b = std::min(n, a + chunk);
      if(a >= b) break;
      th.emplace_back([=]() { f(a, b); });
    }
  for(auto &x : th) x.join();
}
int main(int argc, char **argv)
{
  long long n = 1600000;
  int nth = argc > 1 ? atoi(argv[1]) : (int) std::thread::hardware_concurrency();
  int pf = argc > 2 ? atoi(argv[2]) : 0;	// prefetch distance in particles (0 = none)
  char *P = (char *) aligned_alloc(4096, n * 536 + 4096), *S = (char *) aligned_alloc(4096, n * 264 + 4096);
  double *st = (double *) aligned_alloc(4096, n * 24 * 8);
  pfor(nth, n, [=](long long a, long long b) { memset(P + a * 536, 1, (b - a) * 536); memset(S + a * 264, 1, (b - a) * 264); for(long long i = a * 24; i < b * 24; i++) st[i] = 1.0; });
  printf("threads %d prefetch %d\n", nth, pf);
  for(int rep = 0; rep < 3; rep++)
    {
      double t0 = now();
      pfor(nth, n, [=](long long a, long long b) { memcpy(st + a * 12, st + n * 12 + a * 12, (b - a) * 96); });
      double t1 = now();
      printf("memcpy 154 MB (read+write 307 MB)     %.2f ms  %.1f GB/s\n", t1 - t0, n * 192e-6 / (t1 - t0));
      // pack A: 3 lines of P
      t0 = now();
      pfor(nth, n, [=](long long a, long long b) {
	for(long long i = a; i < b; i++)
	  {
	    const char *pi = P + i * 536;
	    if(pf) { __builtin_prefetch(pi + pf * 536); __builtin_prefetch(pi + pf * 536 + 64); __builtin_prefetch(pi + pf * 536 + 520); }
	    const double *x = (const double *) pi;
	    st[3 * i] = x[0]; st[3 * i + 1] = x[1]; st[3 * i + 2] = x[2];
	    st[3 * n + 3 * i] = x[3]; st[3 * n + 3 * i + 1] = x[4]; st[3 * n + 3 * i + 2] = x[5];
	    st[6 * n + i] = x[6]; st[7 * n + i] = x[12]; st[8 * n + i] = x[14];
	    ((int *) (st + 9 * n))[i] = *(const int *) (pi + 524); ((int *) (st + 10 * n))[i] = *(const short *) (pi + 532);
	  }
      });
      t1 = now();
      printf("pack A (P: 3 lines read)               %.2f ms\n", t1 - t0);
      t0 = now();
      pfor(nth, n, [=](long long a, long long b) {
	for(long long i = a; i < b; i++)
	  {
	    const char *si = S + i * 264;
	    if(pf) { __builtin_prefetch(si + pf * 264); __builtin_prefetch(si + pf * 264 + 88); }
	    const double *x = (const double *) si;
	    st[12 * n + 3 * i] = x[2]; st[12 * n + 3 * i + 1] = x[3]; st[12 * n + 3 * i + 2] = x[4];
	    st[15 * n + i] = x[0]; st[16 * n + i] = x[11];
	  }
      });
      t1 = now();
      printf("pack B (SphP: 2 lines read)            %.2f ms\n", t1 - t0);
      // scatter P gravity: GravAccel, OldAcc (line 1), GravCost (line 8)
      t0 = now();
      pfor(nth, n, [=](long long a, long long b) {
	for(long long i = a; i < b; i++)
	  {
	    char *pi = P + i * 536;
	    if(pf) { __builtin_prefetch(pi + pf * 536 + 64, 1); __builtin_prefetch(pi + pf * 536 + 520, 1); }
	    double *g = (double *) (pi + 64);
	    g[0] = st[3 * i]; g[1] = st[3 * i + 1]; g[2] = st[3 * i + 2]; g[4] = st[6 * n + i];
	    *(float *) (pi + 520) = (float) st[7 * n + i];
	  }
      });
      t1 = now();
      printf("scatter P gravity (2 lines RMW)        %.2f ms\n", t1 - t0);
      t0 = now();
      pfor(nth, n, [=](long long a, long long b) {
	for(long long i = a; i < b; i++)
	  {
	    char *pi = P + i * 536;
	    if(pf) { __builtin_prefetch(pi + pf * 536 + 112, 1); __builtin_prefetch(pi + pf * 536 + 128, 1); }
	    *(double *) (pi + 112) = st[8 * n + i]; *(double *) (pi + 120) = st[6 * n + i]; *(unsigned *) (pi + 128) = 0;
	  }
      });
      t1 = now();
      printf("scatter P density (1-2 lines RMW)      %.2f ms\n", t1 - t0);
      t0 = now();
      pfor(nth, n, [=](long long a, long long b) {
	for(long long i = a; i < b; i++)
	  {
	    char *si = S + i * 264;
	    if(pf) { __builtin_prefetch(si + pf * 264 + 8, 1); __builtin_prefetch(si + pf * 264 + 80, 1); __builtin_prefetch(si + pf * 264 + 152, 1); }
	    double *x = (double *) si;
	    x[1] = st[i]; x[5] = st[n + i]; x[10] = st[2 * n + i]; x[11] = st[3 * n + i];
	    x[12] = st[4 * n + 3 * i]; x[13] = st[4 * n + 3 * i + 1]; x[14] = st[4 * n + 3 * i + 2];
	    x[19] = st[8 * n + i]; x[20] = st[9 * n + i]; x[21] = st[10 * n + i];
	  }
      });
      t1 = now();
      printf("scatter SphP fused (3 lines RMW)       %.2f ms\n", t1 - t0);
    }
  return 0;
}
