// Single-rank MPI stand-in used ONLY to compile the unmodified reference sources
// (under /root/reference) into oracle/_ref/ as the parity oracle / CPU baseline.
// TEST INFRASTRUCTURE — never linked into the product library.
//
// The container has no MPI.  Every communicator has exactly one rank: collectives are
// no-ops (MPI_IN_PLACE semantics), point-to-point calls abort (they are unreachable with
// one learner rank and forked/socket environments).
#ifndef SMB200_ORACLE_MPI_STUB_H
#define SMB200_ORACLE_MPI_STUB_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>

typedef int MPI_Comm;
typedef int MPI_Request;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Errhandler;
struct MPI_Status { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; };

#define MPI_COMM_NULL   0
#define MPI_COMM_WORLD  1
#define MPI_COMM_SELF   2
#define MPI_REQUEST_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_IN_PLACE ((void*)1)
#define MPI_SUCCESS 0
#define MPI_UNDEFINED (-32766)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_ERRORS_RETURN 1

#define MPI_DATATYPE_NULL 0
#define MPI_BYTE 1
#define MPI_INT 2
#define MPI_LONG 3
#define MPI_UNSIGNED_LONG 4
#define MPI_FLOAT 5
#define MPI_DOUBLE 6
#define MPI_LONG_DOUBLE 7
#define MPI_UNSIGNED 8
#define MPI_CHAR 9
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3

static inline void smb200_mpi_stub_unreachable(const char* what) {
  std::fprintf(stderr, "mpi_stub: %s is not available with the single-rank stand-in\n", what);
  std::abort();
}

static inline int MPI_Init_thread(int*, char***, int required, int* provided)
{ if (provided) *provided = required; return MPI_SUCCESS; }
static inline int MPI_Init(int*, char***) { return MPI_SUCCESS; }
static inline int MPI_Query_thread(int* provided) { *provided = MPI_THREAD_MULTIPLE; return MPI_SUCCESS; }
static inline int MPI_Finalize() { return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm, int code) { std::exit(code ? code : 1); return MPI_SUCCESS; }
static inline int MPI_Comm_set_errhandler(MPI_Comm, MPI_Errhandler) { return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm, int* size) { *size = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_dup(MPI_Comm c, MPI_Comm* out) { *out = c; return MPI_SUCCESS; }
static inline int MPI_Comm_split(MPI_Comm c, int color, int, MPI_Comm* out)
{ *out = (color == MPI_UNDEFINED) ? MPI_COMM_NULL : c; return MPI_SUCCESS; }
static inline int MPI_Comm_free(MPI_Comm* c) { if (c) *c = MPI_COMM_NULL; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
static inline size_t smb200_mpi_stub_sizeof(MPI_Datatype t) {
  switch (t) { case MPI_BYTE: case MPI_CHAR: return 1; case MPI_INT: case MPI_FLOAT: case MPI_UNSIGNED: return 4;
    case MPI_LONG: case MPI_UNSIGNED_LONG: case MPI_DOUBLE: return 8; case MPI_LONG_DOUBLE: return sizeof(long double); }
  return 0;
}
static inline int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, MPI_Comm)
{ if (s != MPI_IN_PLACE) std::memcpy(r, s, n * smb200_mpi_stub_sizeof(t)); return MPI_SUCCESS; }
static inline int MPI_Iallreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c, MPI_Request* q)
{ MPI_Allreduce(s, r, n, t, o, c); if (q) *q = MPI_REQUEST_NULL; return MPI_SUCCESS; }
static inline int MPI_Iallgather(const void* s, int n, MPI_Datatype t, void* r, int, MPI_Datatype, MPI_Comm, MPI_Request* q)
{ if (s != MPI_IN_PLACE) std::memcpy(r, s, n * smb200_mpi_stub_sizeof(t)); if (q) *q = MPI_REQUEST_NULL; return MPI_SUCCESS; }
static inline int MPI_Iallgatherv(const void* s, int n, MPI_Datatype t, void* r, const int*, const int*, MPI_Datatype, MPI_Comm, MPI_Request* q)
{ if (s != MPI_IN_PLACE) std::memcpy(r, s, n * smb200_mpi_stub_sizeof(t)); if (q) *q = MPI_REQUEST_NULL; return MPI_SUCCESS; }
static inline int MPI_Wait(MPI_Request* q, MPI_Status*) { if (q) *q = MPI_REQUEST_NULL; return MPI_SUCCESS; }
static inline int MPI_Test(MPI_Request* q, int* flag, MPI_Status*) { if (q) *q = MPI_REQUEST_NULL; *flag = 1; return MPI_SUCCESS; }
static inline int MPI_Request_free(MPI_Request* q) { if (q) *q = MPI_REQUEST_NULL; return MPI_SUCCESS; }
static inline int MPI_Iprobe(int, int, MPI_Comm, int* flag, MPI_Status*) { *flag = 0; return MPI_SUCCESS; }
static inline int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm) { smb200_mpi_stub_unreachable("MPI_Send"); return 1; }
static inline int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*) { smb200_mpi_stub_unreachable("MPI_Recv"); return 1; }
static inline int MPI_Isend(const void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*) { smb200_mpi_stub_unreachable("MPI_Isend"); return 1; }
static inline int MPI_Irecv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*) { smb200_mpi_stub_unreachable("MPI_Irecv"); return 1; }
static inline int MPI_Get_count(const MPI_Status*, MPI_Datatype, int* n) { *n = 0; return MPI_SUCCESS; }

#endif
