// Stand-in for the reference's include/grpcpp/stats_time.h (HdrHistogram-backed
// scoped timers; HdrHistogram_c is not vendored).  TEST INFRASTRUCTURE ONLY:
// lets pair.cc compile unmodified; the profiler is a no-op.
#ifndef ORACLE_SHIM_STATS_TIME_H
#define ORACLE_SHIM_STATS_TIME_H
typedef enum {
  GRPC_STATS_TIME_PAIR_SEND,
  GRPC_STATS_TIME_PAIR_RECV,
  GRPC_STATS_TIME_MAX_OP_SIZE
} grpc_stats_time;
class GRPCProfiler {
 public:
  explicit GRPCProfiler(grpc_stats_time) {}
  ~GRPCProfiler() {}
};
#endif
