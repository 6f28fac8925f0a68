// Stand-in for src/core/lib/iomgr/wakeup_fd_posix.h (the real one drags in
// gRPC's error/slice machinery).  TEST INFRASTRUCTURE ONLY.  Same struct and
// call names; eventfd-backed (gpr_stubs.cc), like wakeup_fd_eventfd.cc.
#ifndef ORACLE_SHIM_WAKEUP_FD_POSIX_H
#define ORACLE_SHIM_WAKEUP_FD_POSIX_H
#include <grpc/support/log.h>  // the real header reaches it through iomgr/error.h
typedef void* grpc_error_handle;
#ifndef GRPC_ERROR_NONE
#define GRPC_ERROR_NONE nullptr
#endif
typedef struct grpc_wakeup_fd {
  int read_fd;
  int write_fd;
} grpc_wakeup_fd;
grpc_error_handle grpc_wakeup_fd_init(grpc_wakeup_fd* fd_info);
grpc_error_handle grpc_wakeup_fd_consume_wakeup(grpc_wakeup_fd* fd_info);
grpc_error_handle grpc_wakeup_fd_wakeup(grpc_wakeup_fd* fd_info);
void grpc_wakeup_fd_destroy(grpc_wakeup_fd* fd_info);
#endif
