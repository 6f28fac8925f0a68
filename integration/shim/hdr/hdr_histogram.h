// Declarations of the HdrHistogram_c subset that include/grpcpp/stats_time.h touches (the reference
// does not vendor HdrHistogram_c; CMake makes it a hard dependency).  Compile-check only.
#ifndef INTEGRATION_SHIM_HDR_HISTOGRAM_H
#define INTEGRATION_SHIM_HDR_HISTOGRAM_H
#include <stdbool.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
struct hdr_histogram;
int hdr_init(int64_t lowest, int64_t highest, int significant_figures, struct hdr_histogram** result);
void hdr_close(struct hdr_histogram* h);
void hdr_reset(struct hdr_histogram* h);
int64_t hdr_add(struct hdr_histogram* h, const struct hdr_histogram* from);
bool hdr_record_value(struct hdr_histogram* h, int64_t value);
int64_t hdr_max(const struct hdr_histogram* h);
double hdr_mean(const struct hdr_histogram* h);
int64_t hdr_value_at_percentile(const struct hdr_histogram* h, double percentile);
#ifdef __cplusplus
}
#endif
#endif
