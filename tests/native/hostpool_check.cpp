// hostpool_check.cpp - CPU check of the drop-in layer's host worker pool (gadget_leicester_b200/csrc/gl_host.h):
// every chunk of every job runs exactly once, chunk bounds tile [0,n), jobs are handed out in FIFO order, submit()
// returns at once and wait() blocks until the job is complete, empty jobs complete immediately, and jobs submitted
// from several threads at once all finish.  Built and run by tests/test_host_logic.py (g++, no CUDA).
#include "gl_host.h"
#include <cstdio>
#include <numeric>

static int fails = 0;
#define CHECK(c) do { if(!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); fails++; } } while(0)

int main()
{
  HostPool pool(4);
  CHECK(pool.threads() == 4);
  // 1. coverage: every index touched exactly once, chunk indices consistent with their bounds
  {
    const long long n = 100003, chunk = 257;
    std::vector<std::atomic<int>> hit(n);
    for(auto &h : hit) h.store(0);
    std::atomic<long long> bad{0};
    pool.run(n, chunk, [&](long long k, long long a, long long b) {
      if(a != k * chunk || b > n || b - a > chunk || b <= a) bad++;
      for(long long i = a; i < b; i++) hit[i]++;
    });
    long long wrong = 0;
    for(auto &h : hit) wrong += (h.load() != 1);
    CHECK(wrong == 0);
    CHECK(bad.load() == 0);
  }
  // 2. submit() does not block; wait() does; FIFO: the first chunk of job B is not handed out before all chunks of job A are
  {
    std::atomic<int> a_started{0}, order_violation{0};
    const int na = 64;
    auto ja = pool.submit(na, 1, [&](long long, long long, long long) {
      a_started++;
      std::this_thread::sleep_for(std::chrono::milliseconds(2));
    });
    auto jb = pool.submit(8, 1, [&](long long, long long, long long) {
      if(ja->next.load() < na) order_violation++;	// every chunk of A must have been handed out by now
    });
    CHECK(!jb->is_finished() || a_started.load() == na);	// B cannot be done before all of A was handed out
    ja->wait();
    jb->wait();
    CHECK(a_started.load() == na);
    CHECK(order_violation.load() == 0);
  }
  // 3. empty job
  {
    auto j = pool.submit(0, 16, [&](long long, long long, long long) { fails++; });
    j->wait();
    CHECK(j->is_finished());
  }
  // 4. concurrent submitters
  {
    std::atomic<long long> sum{0};
    std::vector<std::thread> th;
    for(int t = 0; t < 6; t++)
      th.emplace_back([&]() {
	for(int r = 0; r < 20; r++)
	  pool.run(1000, 37, [&](long long, long long a, long long b) { sum += (b - a); });
      });
    for(auto &x : th) x.join();
    CHECK(sum.load() == 6ll * 20 * 1000);
  }
  printf(fails ? "HOSTPOOL FAIL\n" : "HOSTPOOL OK\n");
  return fails ? 1 : 0;
}
