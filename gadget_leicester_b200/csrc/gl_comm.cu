// gl_comm.cu - transports behind gl_comm.h: NCCL (one process per GPU, NVLink / NVSwitch) and an in-process group.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the library has no link-time dependency on it, a single-rank
// user never loads it, and inside a torch process the already-loaded NCCL is shared.  The communicator is created
// from a 128-byte unique id that the CALLER broadcasts with whatever it has (the reference's MPI_Bcast, a file,
// torch.distributed) - the library never opens a socket of its own for the rendezvous.
#include "gl_comm.h"
#include <nccl.h>
#include <dlfcn.h>
#include <mutex>
#include <condition_variable>
#include <chrono>
#include <vector>
#include <cstring>

// ---------------------------------------------------------------- NCCL ----
namespace
{
struct NcclApi
{
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string &err)
  {
    if(h) return true;
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    for(const char *nm : names)
      if((h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
    if(!h)
      {
	err = std::string("gl_comm: cannot load libnccl.so.2: ") + dlerror();
	return false;
      }
#define GL_SYM(field, name)                                                          \
  field = reinterpret_cast<decltype(field)>(dlsym(h, name));                         \
  if(!field)                                                                         \
    {                                                                                \
      err = std::string("gl_comm: libnccl lacks ") + name;                           \
      h = nullptr;                                                                   \
      return false;                                                                  \
    }
    GL_SYM(GetUniqueId, "ncclGetUniqueId");
    GL_SYM(CommInitRank, "ncclCommInitRank");
    GL_SYM(CommDestroy, "ncclCommDestroy");
    GL_SYM(AllGather, "ncclAllGather");
    GL_SYM(Send, "ncclSend");
    GL_SYM(Recv, "ncclRecv");
    GL_SYM(GroupStart, "ncclGroupStart");
    GL_SYM(GroupEnd, "ncclGroupEnd");
    GL_SYM(GetErrorString, "ncclGetErrorString");
#undef GL_SYM
    return true;
  }
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

struct NcclComm : gl_comm
{
  ncclComm_t comm = nullptr;
  int fail(const char *what, ncclResult_t r)
  {
    err = std::string("NCCL ") + what + ": " + g_nccl.GetErrorString(r);
    return 1;
  }
  ~NcclComm() override
  {
    if(comm) g_nccl.CommDestroy(comm);
  }
  const char *kind() const override { return "nccl"; }
  int allgather(const void *send, void *recv, size_t bytes, cudaStream_t st) override
  {
    ncclResult_t r = g_nccl.AllGather(send, recv, bytes, ncclInt8, comm, st);
    return r == ncclSuccess ? 0 : fail("ncclAllGather", r);
  }
  int alltoallv(const void *send, const size_t *soff, const size_t *scnt, void *recv, const size_t *roff, const size_t *rcnt,
		cudaStream_t st) override
  {
    // the personalised exchange that replaces the reference's hypercube of MPI_Sendrecv pairs (gravtree.c:236-262):
    // all sends and receives of one step are fused into one NCCL group, so every NVLink carries its pair at once
    if(scnt[rank])
      {
	cudaError_t e = cudaMemcpyAsync((char *) recv + roff[rank], (const char *) send + soff[rank], scnt[rank], cudaMemcpyDeviceToDevice, st);
	if(e != cudaSuccess)
	  {
	    err = std::string("alltoallv self copy: ") + cudaGetErrorString(e);
	    return 1;
	  }
      }
    ncclResult_t r = g_nccl.GroupStart();
    if(r != ncclSuccess) return fail("ncclGroupStart", r);
    for(int p = 0; p < world; p++)
      {
	if(p == rank) continue;
	if(scnt[p] && (r = g_nccl.Send((const char *) send + soff[p], scnt[p], ncclInt8, p, comm, st)) != ncclSuccess) return fail("ncclSend", r);
	if(rcnt[p] && (r = g_nccl.Recv((char *) recv + roff[p], rcnt[p], ncclInt8, p, comm, st)) != ncclSuccess) return fail("ncclRecv", r);
      }
    r = g_nccl.GroupEnd();
    return r == ncclSuccess ? 0 : fail("ncclGroupEnd", r);
  }
};
}	// namespace

int gl_comm_nccl_unique_id(void *id128, std::string &err)
{
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if(!g_nccl.load(err)) return 1;
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if(r != ncclSuccess)
    {
      err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r);
      return 1;
    }
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
  return 0;
}

gl_comm *gl_comm_make_nccl(int rank, int world, const void *id128, int device, std::string &err)
{
  {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if(!g_nccl.load(err)) return nullptr;
  }
  cudaSetDevice(device);
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  NcclComm *c = new NcclComm();
  c->rank = rank;
  c->world = world;
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
  if(r != ncclSuccess)
    {
      err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r);
      c->comm = nullptr;
      delete c;
      return nullptr;
    }
  return c;
}

// ---------------------------------------------------------------- in-process group ----
namespace
{
struct LocalGroup
{
  int world;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  long long generation = 0;
  bool aborted = false;
  // what every rank publishes for the current collective
  std::vector<const void *> send;
  std::vector<const size_t *> soff, scnt;
  explicit LocalGroup(int w) : world(w), send(w), soff(w), scnt(w) {}
  // returns false when the group was aborted (a rank failed) or nobody came for 120 s
  bool barrier()
  {
    std::unique_lock<std::mutex> lk(mu);
    if(aborted) return false;
    const long long gen = generation;
    if(++arrived == world)
      {
	arrived = 0;
	generation++;
	cv.notify_all();
	return true;
      }
    const bool ok = cv.wait_for(lk, std::chrono::seconds(120), [&] { return generation != gen || aborted; });
    if(!ok || aborted)
      {
	aborted = true;
	cv.notify_all();
	return false;
      }
    return true;
  }
  void abort()
  {
    std::lock_guard<std::mutex> lk(mu);
    aborted = true;
    cv.notify_all();
  }
};

struct LocalComm : gl_comm
{
  LocalGroup *g = nullptr;
  const char *kind() const override { return "local"; }
  int bad(const std::string &m)
  {
    err = m;
    g->abort();
    return 1;
  }
  int allgather(const void *send, void *recv, size_t bytes, cudaStream_t st) override
  {
    if(cudaStreamSynchronize(st) != cudaSuccess) return bad("local allgather: stream error before the exchange");
    g->send[rank] = send;
    if(!g->barrier()) return bad("local allgather: a rank of the group failed or did not arrive");
    for(int r = 0; r < world; r++)
      if(cudaMemcpyAsync((char *) recv + (size_t) r * bytes, g->send[r], bytes, cudaMemcpyDefault, st) != cudaSuccess)
	return bad("local allgather: copy failed");
    if(cudaStreamSynchronize(st) != cudaSuccess) return bad("local allgather: copy failed");
    if(!g->barrier()) return bad("local allgather: a rank of the group failed");	// nobody reuses its send buffer before everyone has read it
    return 0;
  }
  int alltoallv(const void *send, const size_t *soff, const size_t *scnt, void *recv, const size_t *roff, const size_t *rcnt,
		cudaStream_t st) override
  {
    if(cudaStreamSynchronize(st) != cudaSuccess) return bad("local alltoallv: stream error before the exchange");
    g->send[rank] = send;
    g->soff[rank] = soff;
    g->scnt[rank] = scnt;
    if(!g->barrier()) return bad("local alltoallv: a rank of the group failed or did not arrive");
    for(int s = 0; s < world; s++)
      {
	const size_t cnt = g->scnt[s][rank];
	if(cnt != rcnt[s]) return bad("local alltoallv: send and receive counts disagree");
	if(cnt == 0) continue;
	if(cudaMemcpyAsync((char *) recv + roff[s], (const char *) g->send[s] + g->soff[s][rank], cnt, cudaMemcpyDefault, st) != cudaSuccess)
	  return bad("local alltoallv: copy failed");
      }
    if(cudaStreamSynchronize(st) != cudaSuccess) return bad("local alltoallv: copy failed");
    if(!g->barrier()) return bad("local alltoallv: a rank of the group failed");
    return 0;
  }
};
}	// namespace

void *gl_comm_local_group_create_impl(int world) { return world >= 1 && world <= 64 ? new LocalGroup(world) : nullptr; }
void gl_comm_local_group_destroy_impl(void *group) { delete static_cast<LocalGroup *>(group); }
gl_comm *gl_comm_make_local(void *group, int rank, std::string &err)
{
  LocalGroup *g = static_cast<LocalGroup *>(group);
  if(!g || rank < 0 || rank >= g->world)
    {
      err = "gl_comm_init_local: bad group or rank";
      return nullptr;
    }
  LocalComm *c = new LocalComm();
  c->g = g;
  c->rank = rank;
  c->world = g->world;
  return c;
}
