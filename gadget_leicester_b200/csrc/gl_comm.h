// gl_comm.h - the two collectives the multi-rank force path needs, over NCCL or between contexts of one process.
//
// The reference exchanges particles and partial results with MPI_Sendrecv / MPI_Allgather(v) pairs inside every force
// loop (gravtree.c:175-347, density.c:193-342, hydra.c:274-520) and pseudo-particle moments with MPI_Allgatherv
// (forcetree.c:963).  Here every rank owns one GPU; what crosses ranks is (1) small fixed-size blocks every rank
// needs from every other (domain extent, key ranges, target boxes, export counts): an all-gather, and (2) the
// locally-essential tree items a rank selects for each other rank: a personalised all-to-all.  Both on device buffers.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <string>

struct gl_comm
{
  int rank = 0, world = 1;
  std::string err;
  virtual ~gl_comm() {}
  // every rank contributes `bytes` from send (device), recv (device) gets world * bytes in rank order
  virtual int allgather(const void *send, void *recv, size_t bytes, cudaStream_t st) = 0;
  // rank r sends send + soff[d] .. + scnt[d] bytes to rank d and receives rcnt[s] bytes from rank s at recv + roff[s]
  // (host arrays of `world` entries; the self block is copied locally)
  virtual int alltoallv(const void *send, const size_t *soff, const size_t *scnt, void *recv, const size_t *roff,
			const size_t *rcnt, cudaStream_t st) = 0;
  virtual const char *kind() const = 0;
};

gl_comm *gl_comm_make_nccl(int rank, int world, const void *id128, int device, std::string &err);
int gl_comm_nccl_unique_id(void *id128, std::string &err);
// in-process group: `world` contexts of ONE process (one host thread each) exchange through device-to-device copies;
// this is how the multi-rank path is tested on a single GPU and how several GPUs of one process are driven
void *gl_comm_local_group_create_impl(int world);
void gl_comm_local_group_destroy_impl(void *group);
gl_comm *gl_comm_make_local(void *group, int rank, std::string &err);
