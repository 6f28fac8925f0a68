// dev_helpers.cc -- TEST INFRASTRUCTURE: the product's frame / credit arithmetic (the B200_HD inlines of
// grpc-rdma_b200/csrc/b200_dev.cuh that the kernels and the host runtime share) compiled for the host, so the
// CPU tests can compare it with the reference's own functions (ring_buffer.h:180-189, ring_buffer.cc:99-116).
#include <stddef.h>
#include <stdint.h>

#define __align__(n) __attribute__((aligned(n)))  // nvcc spelling, for the host compiler
#include "../../grpc-rdma_b200/csrc/b200_dev.cuh"

extern "C" {
uint64_t dev_round_up8(uint64_t v) { return b200::round_up8(v); }
uint64_t dev_encoded_size(uint64_t p) { return b200::encoded_size(p); }
uint64_t dev_calc_writable(uint64_t s) { return b200::calc_writable(s); }
uint64_t dev_free_size(uint64_t cap, uint64_t head, uint64_t tail) { return b200::free_size(cap, head, tail); }
uint64_t dev_writable_size(uint64_t cap, uint64_t head, uint64_t tail) { return b200::writable_size(cap, head, tail); }
uint64_t dev_sizeof_pairdev() { return sizeof(b200::PairDev); }
uint64_t dev_sizeof_svccmd() { return sizeof(b200::SvcCmd); }
uint64_t dev_sizeof_svcdone() { return sizeof(b200::SvcDone); }
uint64_t dev_sizeof_eagerrec() { return sizeof(b200::EagerRec); }
uint64_t dev_offset_stamp2() { return offsetof(b200::SvcCmd, stamp2); }
// the eager checksum exactly as the owner warp (lane-strided XOR of eager_word) and the host (serial) compute it
uint64_t dev_eager_checksum(const uint8_t* payload, uint32_t size, uint64_t at, int lanes) {
  uint64_t per_lane[64] = {0};
  const uint32_t words = (size + 7) >> 3;
  for (uint32_t j = 0; j < words; j++) {
    uint64_t w = 0;
    for (uint32_t k = 0; k < 8 && 8 * j + k < size; k++) w |= (uint64_t)payload[8 * j + k] << (8 * k);
    per_lane[j % lanes] ^= b200::eager_word(w, j);
  }
  uint64_t cs = 0;
  for (int l = 0; l < lanes; l++) cs ^= per_lane[l];
  return cs ^ b200::eager_mix(at * 31 + size);
}
}
