/* Single-rank stand-in for <mpi.h>, used ONLY to compile the reference's own
 * hot-path sources (from /root/reference, never copied) into oracle/_ref/.
 * TEST INFRASTRUCTURE - not part of the product.  See oracle/refbuild/README.md.
 * Every collective degenerates to a memmove of the send buffer (NTask == 1).   */
#ifndef GL_STUB_MPI_H
#define GL_STUB_MPI_H
#include <stddef.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_BYTE   1
#define MPI_INT    4
#define MPI_FLOAT  5
#define MPI_DOUBLE 8
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_STATUS_IGNORE   ((MPI_Status *) 0)
#define MPI_STATUSES_IGNORE ((MPI_Status *) 0)
#define MPI_SUCCESS 0
int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm c, int code);
int MPI_Barrier(MPI_Comm c);
int MPI_Comm_rank(MPI_Comm c, int *r);
int MPI_Comm_size(MPI_Comm c, int *s);
double MPI_Wtime(void);
int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Allgather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c);
int MPI_Allgatherv(const void *s, int ns, MPI_Datatype ts, void *r, const int *nr, const int *displs,
                   MPI_Datatype tr, MPI_Comm c);
int MPI_Gather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, int root, MPI_Comm c);
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c);
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c);
int MPI_Sendrecv(const void *s, int ns, MPI_Datatype ts, int dest, int stag, void *r, int nr, MPI_Datatype tr,
                 int src, int rtag, MPI_Comm c, MPI_Status *st);
int MPI_Send(const void *s, int n, MPI_Datatype t, int dest, int tag, MPI_Comm c);
int MPI_Ssend(const void *s, int n, MPI_Datatype t, int dest, int tag, MPI_Comm c);
int MPI_Recv(void *r, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st);
int MPI_Isend(const void *s, int n, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request *rq);
int MPI_Irecv(void *r, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *rq);
int MPI_Waitall(int n, MPI_Request *rq, MPI_Status *st);
#endif
