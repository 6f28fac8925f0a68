/* mpi.h -- single-rank stand-in for the MPI calls of the reference's micro-benchmark driver
 * (examples/cpp/micro-bench/mb_client.cc: one MPI rank = one client connection; the image has no MPI).  Every
 * collective degenerates to a copy: rank 0 of 1.  Only for building that driver offline (integration/stack). */
#ifndef STACK_SHIM_MPI_H
#define STACK_SHIM_MPI_H
#include <string.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_WORLD 0
#define MPI_INT 4
#define MPI_DOUBLE 8
#define MPI_LONG_LONG 9
#define MPI_LONG 10
#define MPI_INT64_T 11
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_SUCCESS 0
static inline int mpi_shim_size(MPI_Datatype t) { return t == MPI_INT ? 4 : 8; }
static inline int MPI_Init(int*, char***) { return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int* n) { *n = 1; return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) {
  memcpy(r, s, (size_t)n * mpi_shim_size(t));
  return 0;
}
static inline int MPI_Gather(const void* s, int n, MPI_Datatype t, void* r, int, MPI_Datatype, int, MPI_Comm) {
  memcpy(r, s, (size_t)n * mpi_shim_size(t));
  return 0;
}
static inline int MPI_Gatherv(const void* s, int n, MPI_Datatype t, void* r, const int*, const int* displs, MPI_Datatype,
                              int, MPI_Comm) {
  memcpy((char*)r + (size_t)(displs ? displs[0] : 0) * mpi_shim_size(t), s, (size_t)n * mpi_shim_size(t));
  return 0;
}
#endif
