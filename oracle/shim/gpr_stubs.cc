// Minimal gpr pieces the reference ibverbs files link against.
// TEST INFRASTRUCTURE ONLY.
#include <grpc/support/log.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/eventfd.h>
#include <unistd.h>

#include "src/core/lib/gpr/env.h"
#include "src/core/lib/iomgr/wakeup_fd_posix.h"

extern "C" void gpr_log(const char* file, int line, gpr_log_severity severity, const char* format, ...) {
  static int verbose = getenv("ORACLE_REF_VERBOSE") != nullptr;
  if (severity != GPR_LOG_SEVERITY_ERROR && !verbose) return;
  va_list ap;
  va_start(ap, format);
  fprintf(stderr, "[ref %s:%d] ", file, line);
  vfprintf(stderr, format, ap);
  fputc('\n', stderr);
  va_end(ap);
}
extern "C" int gpr_should_log(gpr_log_severity) { return 1; }
extern "C" void gpr_log_message(const char* file, int line, gpr_log_severity s, const char* m) {
  gpr_log(file, line, s, "%s", m);
}
char* gpr_getenv(const char* name) {
  const char* v = getenv(name);
  return v ? strdup(v) : nullptr;
}
grpc_error_handle grpc_wakeup_fd_init(grpc_wakeup_fd* fd_info) {
  fd_info->read_fd = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC);
  fd_info->write_fd = -1;
  return fd_info->read_fd < 0 ? (grpc_error_handle)1 : GRPC_ERROR_NONE;
}
grpc_error_handle grpc_wakeup_fd_consume_wakeup(grpc_wakeup_fd* fd_info) {
  eventfd_t v;
  (void)eventfd_read(fd_info->read_fd, &v);
  return GRPC_ERROR_NONE;
}
grpc_error_handle grpc_wakeup_fd_wakeup(grpc_wakeup_fd* fd_info) {
  (void)eventfd_write(fd_info->read_fd, 1);
  return GRPC_ERROR_NONE;
}
void grpc_wakeup_fd_destroy(grpc_wakeup_fd* fd_info) {
  if (fd_info->read_fd >= 0) close(fd_info->read_fd);
  fd_info->read_fd = -1;
}
