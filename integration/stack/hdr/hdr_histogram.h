/* hdr/hdr_histogram.h -- minimal stand-in for HdrHistogram_c (which the reference does not vendor and
 * its CMake treats as a hard dependency, cmake/hdrhistorgram.cmake).  Own implementation of the subset
 * include/grpcpp/stats_time.h and examples/cpp/micro-bench use: log-linear buckets with 2^11
 * sub-buckets per octave (better than 3 significant figures).  Only used to let the reference stack
 * build and run offline (integration/stack); not part of the product library. */
#ifndef STACK_HDR_HISTOGRAM_H
#define STACK_HDR_HISTOGRAM_H
#include <stdbool.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
struct hdr_histogram {
  int64_t lowest, highest;
  int64_t total_count, min_value, max_value;
  double sum;
  int32_t counts_len;
  int64_t* counts;
};
struct hdr_iter {
  const struct hdr_histogram* h;
  int32_t index;
  int64_t count, value, cumulative_count;
};
int hdr_init(int64_t lowest, int64_t highest, int significant_figures, struct hdr_histogram** result);
void hdr_close(struct hdr_histogram* h);
void hdr_reset(struct hdr_histogram* h);
int64_t hdr_add(struct hdr_histogram* h, const struct hdr_histogram* from);
bool hdr_record_value(struct hdr_histogram* h, int64_t value);
bool hdr_record_values(struct hdr_histogram* h, int64_t value, int64_t count);
int64_t hdr_min(const struct hdr_histogram* h);
int64_t hdr_max(const struct hdr_histogram* h);
double hdr_mean(const struct hdr_histogram* h);
int64_t hdr_value_at_percentile(const struct hdr_histogram* h, double percentile);
void hdr_iter_init(struct hdr_iter* it, const struct hdr_histogram* h);
bool hdr_iter_next(struct hdr_iter* it);
#ifdef __cplusplus
}
#endif
#endif
