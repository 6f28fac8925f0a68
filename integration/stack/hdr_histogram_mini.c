/* See hdr/hdr_histogram.h.  Bucket b = octave * 2048 + mantissa, octave = position of the top bit
 * above bit 11; values below 4096 are exact. */
#include "hdr/hdr_histogram.h"

#include <stdlib.h>
#include <string.h>

enum { kSubBits = 11, kSub = 1 << kSubBits };

static int32_t bucket_of(int64_t v) {
  if (v < 0) v = 0;
  if (v < 2 * kSub) return (int32_t)v;
  int top = 63 - __builtin_clzll((unsigned long long)v); /* >= kSubBits + 1 */
  int shift = top - kSubBits;
  return (int32_t)(((int64_t)(shift + 1) << kSubBits) + ((v >> shift) - kSub));
}
static int64_t value_of(int32_t b) {
  if (b < 2 * kSub) return b;
  int shift = (b >> kSubBits) - 1;
  int64_t mant = (b & (kSub - 1)) + kSub;
  return ((mant + 1) << shift) - 1; /* highest equivalent value of the bucket */
}

int hdr_init(int64_t lowest, int64_t highest, int significant_figures, struct hdr_histogram** result) {
  (void)significant_figures;
  if (lowest < 1 || highest < 2 * lowest) return 22;
  struct hdr_histogram* h = (struct hdr_histogram*)calloc(1, sizeof *h);
  if (!h) return 12;
  h->lowest = lowest;
  h->highest = highest;
  h->counts_len = bucket_of(highest) + 1;
  h->counts = (int64_t*)calloc((size_t)h->counts_len, sizeof(int64_t));
  if (!h->counts) {
    free(h);
    return 12;
  }
  h->min_value = INT64_MAX;
  *result = h;
  return 0;
}
void hdr_close(struct hdr_histogram* h) {
  if (!h) return;
  free(h->counts);
  free(h);
}
void hdr_reset(struct hdr_histogram* h) {
  memset(h->counts, 0, sizeof(int64_t) * (size_t)h->counts_len);
  h->total_count = 0;
  h->max_value = 0;
  h->min_value = INT64_MAX;
  h->sum = 0;
}
bool hdr_record_values(struct hdr_histogram* h, int64_t value, int64_t count) {
  if (value < 0 || value > h->highest) return false;
  h->counts[bucket_of(value)] += count;
  h->total_count += count;
  h->sum += (double)value * (double)count;
  if (value > h->max_value) h->max_value = value;
  if (value < h->min_value) h->min_value = value;
  return true;
}
bool hdr_record_value(struct hdr_histogram* h, int64_t value) { return hdr_record_values(h, value, 1); }
int64_t hdr_add(struct hdr_histogram* h, const struct hdr_histogram* from) {
  int64_t dropped = 0;
  for (int32_t b = 0; b < from->counts_len; b++) {
    if (!from->counts[b]) continue;
    if (b < h->counts_len) {
      h->counts[b] += from->counts[b];
      h->total_count += from->counts[b];
    } else {
      dropped += from->counts[b];
    }
  }
  h->sum += from->sum;
  if (from->max_value > h->max_value) h->max_value = from->max_value;
  if (from->min_value < h->min_value) h->min_value = from->min_value;
  return dropped;
}
int64_t hdr_min(const struct hdr_histogram* h) { return h->total_count ? h->min_value : 0; }
int64_t hdr_max(const struct hdr_histogram* h) { return h->max_value; }
double hdr_mean(const struct hdr_histogram* h) { return h->total_count ? h->sum / (double)h->total_count : 0.0; }
int64_t hdr_value_at_percentile(const struct hdr_histogram* h, double percentile) {
  if (!h->total_count) return 0;
  if (percentile > 100.0) percentile = 100.0;
  int64_t want = (int64_t)((percentile / 100.0) * (double)h->total_count + 0.5);
  if (want < 1) want = 1;
  int64_t seen = 0;
  for (int32_t b = 0; b < h->counts_len; b++) {
    seen += h->counts[b];
    if (seen >= want) {
      int64_t v = value_of(b);
      return v > h->max_value ? h->max_value : v;
    }
  }
  return h->max_value;
}
void hdr_iter_init(struct hdr_iter* it, const struct hdr_histogram* h) {
  it->h = h;
  it->index = -1;
  it->count = it->value = it->cumulative_count = 0;
}
bool hdr_iter_next(struct hdr_iter* it) {
  if (it->index + 1 >= it->h->counts_len) return false;
  it->index++;
  it->count = it->h->counts[it->index];
  it->value = value_of(it->index);
  it->cumulative_count += it->count;
  return true;
}
