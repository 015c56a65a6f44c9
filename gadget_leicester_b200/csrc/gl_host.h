// gl_host.h - host-side worker pool of the drop-in layer (pack P[]/SphP[] into the pinned SoA staging, scatter
// results back).  Those loops are bound by host memory bandwidth, so they are spread over all cores but ONE: the
// thread that drives the GPU (kernel launches, stream waits) must never queue behind a worker, otherwise the
// device idles for a scheduler time slice.  Work is handed out in small chunks from an atomic counter (a static
// split makes the whole loop wait for whichever thread the OS de-scheduled).  Jobs are FIFO and may overlap at
// their ends; submit() returns at once, HostJob::wait() blocks without spinning.
#pragma once
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#include <cstdlib>
#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>

struct HostJob
{
  long long n = 0, chunk = 1, nchunks = 0;
  std::function<void(long long, long long, long long)> fn;	// (chunk index, first, last+1)
  std::atomic<long long> next{0}, done{0};
  std::mutex m;
  std::condition_variable cv;
  bool finished = false;
  void wait()
  {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return finished; });
  }
  bool is_finished()
  {
    std::lock_guard<std::mutex> lk(m);
    return finished;
  }
};

class HostPool
{
public:
  static HostPool &get()
  {
    static HostPool pool(default_threads());
    return pool;
  }
  static int default_threads()
  {
    if(const char *e = getenv("GL_HOST_THREADS"))
      if(atoi(e) > 0) return atoi(e);
    int hw = (int) std::thread::hardware_concurrency();
    if(hw < 2) hw = 2;
    int share = 1;		// ranks of one node share the cores (torchrun exports LOCAL_WORLD_SIZE)
    if(const char *e = getenv("LOCAL_WORLD_SIZE"))
      if(atoi(e) > 1) share = atoi(e);
    int t = (hw - 1) / share;
    if(t < 1) t = 1;
    if(t > 63) t = 63;
    return t;
  }
  explicit HostPool(int nthreads)
  {
    for(int t = 0; t < nthreads; t++) th.emplace_back([this]() { worker(); });
  }
  ~HostPool()
  {
    {
      std::lock_guard<std::mutex> lk(m);
      stop = true;
    }
    cv.notify_all();
    for(auto &x : th) x.join();
  }
  int threads() const { return (int) th.size(); }

  std::shared_ptr<HostJob> submit(long long n, long long chunk, std::function<void(long long, long long, long long)> fn)
  {
    auto j = std::make_shared<HostJob>();
    j->n = n;
    j->chunk = chunk < 1 ? 1 : chunk;
    j->nchunks = n > 0 ? (n + j->chunk - 1) / j->chunk : 0;
    j->fn = std::move(fn);
    if(j->nchunks == 0)
      {
	j->finished = true;
	return j;
      }
    {
      std::lock_guard<std::mutex> lk(m);
      q.push_back(j);
    }
    cv.notify_all();
    return j;
  }
  void run(long long n, long long chunk, std::function<void(long long, long long, long long)> fn) { submit(n, chunk, std::move(fn))->wait(); }

private:
  std::vector<std::thread> th;
  std::mutex m;
  std::condition_variable cv;
  std::deque<std::shared_ptr<HostJob>> q;
  bool stop = false;

  void worker()
  {
    // workers yield to the GPU-driving thread: when it wakes from a stream wait it must get a core at once
    setpriority(PRIO_PROCESS, (id_t) syscall(SYS_gettid), 10);
    for(;;)
      {
	std::shared_ptr<HostJob> j;
	{
	  std::unique_lock<std::mutex> lk(m);
	  cv.wait(lk, [&] { return stop || !q.empty(); });
	  if(stop && q.empty()) return;
	  j = q.front();
	}
	const long long k = j->next.fetch_add(1);
	if(k >= j->nchunks)
	  {
	    // nothing left to hand out: retire the job from the queue (the last chunks may still be running elsewhere)
	    std::lock_guard<std::mutex> lk(m);
	    if(!q.empty() && q.front() == j) q.pop_front();
	    continue;
	  }
	const long long a = k * j->chunk, b = (a + j->chunk < j->n) ? a + j->chunk : j->n;
	j->fn(k, a, b);
	if(j->done.fetch_add(1) + 1 == j->nchunks)
	  {
	    {
	      std::lock_guard<std::mutex> lk(j->m);
	      j->finished = true;
	    }
	    j->cv.notify_all();
	  }
      }
  }
};
