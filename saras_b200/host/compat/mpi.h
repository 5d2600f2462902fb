/*
 * mpi.h -- "mpi-lite": stand-in for the MPI subset SARAS's lib/poisson, lib/grid, lib/parallel and
 * lib/io/parser call, for machines without an MPI installation (this container).  Two modes:
 *
 *   thread ranks  saras_mpi_run(n, fn, arg) runs fn on n threads, thread r being MPI rank r of one
 *                 in-process world.  Point-to-point follows the only pattern the Poisson solver uses
 *                 (all ranks: post Irecvs, then the same sequence of blocking Sends, then Waitall):
 *                 the k-th MPI_Send of every rank executes in lock step (barrier, direct copy into the
 *                 posted receive buffer, barrier).  That is one legal MPI schedule -- the one in which
 *                 every x-face message has landed before any y-face row is read -- and it makes the
 *                 reference's otherwise racy edge pads deterministic.
 *   process rank  one process per rank (torchrun): rank/size come from SARAS_MPI_RANK/SIZE or
 *                 RANK/WORLD_SIZE; collectives go through hooks installed with saras_mpi_set_hooks
 *                 (the Python tests install torch.distributed-backed ones).  Without threads and
 *                 without the environment variables the world has one rank.
 *
 * With a real MPI installation this header is simply left off the include path.
 */
#ifndef SARAS_COMPAT_MPI_H
#define SARAS_COMPAT_MPI_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_PROC_NULL (-2)
#define MPI_ANY_TAG (-1)
#define MPI_REQUEST_NULL (-1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)

#define MPI_BYTE 1
#define MPI_CHAR 1
#define MPI_INT 2
#define MPI_FLOAT 3
#define MPI_DOUBLE 4
#define MPI_LONG 5
#define MPI_UNSIGNED 6

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *newcomm);
int MPI_Comm_free(MPI_Comm *comm);
int MPI_Barrier(MPI_Comm comm);
double MPI_Wtime(void);

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype, MPI_Datatype *newtype);
#define MPI_ORDER_C 56
int MPI_Type_create_subarray(int ndims, const int *sizes, const int *subsizes, const int *starts, int order, MPI_Datatype oldtype,
                             MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *type);
int MPI_Type_free(MPI_Datatype *type);

int MPI_Irecv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int MPI_Waitall(int count, MPI_Request *reqs, MPI_Status *statuses);
int MPI_Wait(MPI_Request *req, MPI_Status *status);

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int recvcount,
                  MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Allgatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, const int *recvcounts,
                   const int *displs, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);

/* ---- mpi-lite extensions ---- */
/* run fn(arg) on nranks threads forming one world; returns when all have finished */
void saras_mpi_run(int nranks, void (*fn)(void *), void *arg);
/* process-rank mode: world-communicator collectives delegated to the host application */
typedef struct {
    void (*barrier)(void *ctx);
    void (*allgather_bytes)(void *ctx, const void *send, void *recv, int nbytes_each);   /* recv: size*nbytes_each */
    void *ctx;
} saras_mpi_hooks;
void saras_mpi_set_hooks(const saras_mpi_hooks *hooks);
void saras_mpi_set_process_rank(int rank, int size);

#ifdef __cplusplus
}
#endif
#endif
