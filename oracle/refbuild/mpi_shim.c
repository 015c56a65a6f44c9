/* Single-rank implementation of the MPI calls the reference hot path makes.
 * TEST INFRASTRUCTURE (oracle/_ref build only).                                */
#include <mpi.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <time.h>

static size_t tsize(MPI_Datatype t)
{
  /* the datatype constants in stub/mpi.h are the element sizes, except MPI_FLOAT (5, kept distinct from MPI_INT) */
  return t == MPI_FLOAT ? sizeof(float) : (size_t) t;
}
int MPI_Init(int *argc, char ***argv) { (void) argc; (void) argv; return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Abort(MPI_Comm c, int code) { (void) c; fflush(stdout); exit(code ? code : 1); }
int MPI_Barrier(MPI_Comm c) { (void) c; return 0; }
int MPI_Comm_rank(MPI_Comm c, int *r) { (void) c; *r = 0; return 0; }
int MPI_Comm_size(MPI_Comm c, int *s) { (void) c; *s = 1; return 0; }
double MPI_Wtime(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c)
{ (void) b; (void) n; (void) t; (void) root; (void) c; return 0; }
int MPI_Allgather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c)
{ (void) nr; (void) tr; (void) c; if(s != r) memmove(r, s, ns * tsize(ts)); return 0; }
int MPI_Allgatherv(const void *s, int ns, MPI_Datatype ts, void *r, const int *nr, const int *displs,
                   MPI_Datatype tr, MPI_Comm c)
{
  (void) nr; (void) c;
  char *dst = (char *) r + displs[0] * tsize(tr);
  if((const void *) dst != s) memmove(dst, s, ns * tsize(ts));
  return 0;
}
int MPI_Gather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, int root, MPI_Comm c)
{ (void) nr; (void) tr; (void) root; (void) c; if(s != r) memmove(r, s, ns * tsize(ts)); return 0; }
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{ (void) op; (void) c; if(s != r) memmove(r, s, n * tsize(t)); return 0; }
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{ (void) op; (void) root; (void) c; if(s != r) memmove(r, s, n * tsize(t)); return 0; }
int MPI_Sendrecv(const void *s, int ns, MPI_Datatype ts, int dest, int stag, void *r, int nr, MPI_Datatype tr,
                 int src, int rtag, MPI_Comm c, MPI_Status *st)
{
  (void) dest; (void) stag; (void) nr; (void) tr; (void) src; (void) rtag; (void) c; (void) st;
  if(s != r) memmove(r, s, ns * tsize(ts));
  return 0;
}
static int p2p_unreachable(const char *what)
{
  fprintf(stderr, "mpi_shim: %s called with NTask==1 (there is no peer)\n", what);
  exit(97);
  return 0;
}
int MPI_Send(const void *s, int n, MPI_Datatype t, int d, int tag, MPI_Comm c)
{ (void) s; (void) n; (void) t; (void) d; (void) tag; (void) c; return p2p_unreachable("MPI_Send"); }
int MPI_Ssend(const void *s, int n, MPI_Datatype t, int d, int tag, MPI_Comm c)
{ (void) s; (void) n; (void) t; (void) d; (void) tag; (void) c; return p2p_unreachable("MPI_Ssend"); }
int MPI_Recv(void *r, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st)
{ (void) r; (void) n; (void) t; (void) src; (void) tag; (void) c; (void) st; return p2p_unreachable("MPI_Recv"); }
int MPI_Isend(const void *s, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *rq)
{ (void) s; (void) n; (void) t; (void) d; (void) tag; (void) c; (void) rq; return p2p_unreachable("MPI_Isend"); }
int MPI_Irecv(void *r, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *rq)
{ (void) r; (void) n; (void) t; (void) src; (void) tag; (void) c; (void) rq; return p2p_unreachable("MPI_Irecv"); }
int MPI_Waitall(int n, MPI_Request *rq, MPI_Status *st) { (void) n; (void) rq; (void) st; return 0; }
