// lat_probe.cu -- EXPERIMENT HELPER (not product code): host <-> resident-kernel signalling latencies
// over PCIe, the budget of the unary path.  One resident warp polls a command word in pinned host
// memory and answers into pinned host memory; the host measures the round trip of 20000 pings per
// variant and prints p50 / p99 in microseconds.
//
//   v0  bare echo: poll (ld.relaxed.sys u32 by lane 0) -> st answer
//   v1  poll 128-byte line by 8 lanes (uint4 each), stamp in word 0 -> st answer
//   v2  v1 + read 1 KiB payload from pinned host memory (second trip) -> st answer
//   v3  v2 + write the 1 KiB to device memory (ring) + fence.sys -> st answer
//   v4  v1, answer = 1 KiB payload pushed to host (32 lanes x 2 x 16 B) + fence.sys + flag
//   v5  v1 with payload inline: 1 KiB + stamps in every 64-byte line, polled warp-wide (no 2nd trip),
//       answer = 1 KiB pushed to host + flag (the "eager" one-way path both ways)
//   v6  v0 + dependent device-memory load (PairDev line) before the answer
//   v7  v0 with the answer released by st.release.sys (instead of plain volatile store)
//   v8  v4 without the fence (posted writes in program order from one warp: payload lanes then flag
//       by lane 0 after __syncwarp) -- checks data integrity on the host side
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <vector>

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e = (x);                                                         \
    if (e != cudaSuccess) {                                                      \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                    \
      exit(1);                                                                   \
    }                                                                            \
  } while (0)

__device__ __forceinline__ uint32_t ld_sys_u32(const void* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_sys_v4(const void* p) {
  uint4 r;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_sys_v4(void* p, uint4 v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_sys_u32(void* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_rel_sys_u32(void* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct Shared {            // pinned host memory
  uint8_t cmd[2048];       // line 0: stamp + header; payload behind it (v2..v4) or stamped lines (v5)
  uint8_t payload[2048];   // v2/v3: payload the command points to
  uint8_t ans[2048];       // answer area: [0..1024) payload push, flag at 1024
};

__global__ void k_probe(Shared* sh, uint8_t* dring, uint64_t* dstate, int variant, int n) {
  const uint32_t lane = threadIdx.x;
  volatile uint32_t* flag = (volatile uint32_t*)(sh->ans + 1024);
  for (uint32_t k = 1; k <= (uint32_t)n; k++) {
    uint4 a = make_uint4(0, 0, 0, 0), b = a;
    if (variant == 0 || variant == 6 || variant == 7) {
      if (lane == 0)
        while (ld_sys_u32(sh->cmd) != k) {
        }
      __syncwarp();
    } else if (variant == 5) {
      // 1 KiB payload in 60-byte pieces: every 64-byte line carries the stamp in its last word.
      // 18 lines = 1152 B: lanes 0..31 load 16 B each, three rounds (72 x 16 B)
      while (true) {
        uint4 v[3];
        bool ok = true;
#pragma unroll
        for (int r = 0; r < 3; r++) {
          const uint32_t idx = r * 32 + lane;  // 16-byte chunk index
          if (idx < 72) {
            v[r] = ld_sys_v4(sh->cmd + 16 * idx);
            if ((idx & 3) == 3 && v[r].w != k) ok = false;  // last chunk of a line holds the stamp
          }
        }
        if (__all_sync(0xffffffffu, ok)) {
          a = v[0];
          b = v[1];
          break;
        }
      }
    } else {
      while (true) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (lane < 8) v = ld_sys_v4(sh->cmd + 16 * lane);
        const uint32_t stamp = __shfl_sync(0xffffffffu, v.x, 0);
        if (stamp == k) break;
      }
    }
    if (variant == 2 || variant == 3) {  // second trip: payload from pinned memory
      a = ld_sys_v4(sh->payload + 16 * lane);
      b = ld_sys_v4(sh->payload + 512 + 16 * lane);
      if (variant == 3) {
        *(uint4*)(dring + 16 * lane) = a;
        *(uint4*)(dring + 512 + 16 * lane) = b;
        __threadfence_system();
      }
      // the loads must have completed before the answer
      if ((a.x ^ b.x) == 0xdeadbeefu) sh->ans[5] = 1;
    }
    if (variant == 6) {
      const uint64_t s = *(volatile uint64_t*)(dstate + (k & 1023) * 16);
      if (s == 0xdeadbeefull) sh->ans[5] = 1;
    }
    if (variant == 4 || variant == 5 || variant == 8) {
      a.x = k;
      b.w = k;
      st_sys_v4(sh->ans + 16 * lane, a);
      st_sys_v4(sh->ans + 512 + 16 * lane, b);
      if (variant != 8) __threadfence_system();
    }
    __syncwarp();
    if (lane == 0) {
      if (variant == 7) st_rel_sys_u32((void*)flag, k);
      else st_sys_u32((void*)flag, k);
    }
  }
}

static uint64_t now_ns() {
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (uint64_t)t.tv_sec * 1000000000ull + t.tv_nsec;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 20000;
  CK(cudaSetDevice(0));
  Shared* sh;
  CK(cudaHostAlloc(&sh, sizeof(Shared), cudaHostAllocMapped | cudaHostAllocPortable));
  uint8_t* dring;
  uint64_t* dstate;
  CK(cudaMalloc(&dring, 1 << 20));
  CK(cudaMalloc(&dstate, 1 << 20));
  CK(cudaMemset(dstate, 0, 1 << 20));
  cudaStream_t s;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  for (int variant = 0; variant <= 8; variant++) {
    memset(sh, 0, sizeof(Shared));
    for (int i = 0; i < 1024; i++) sh->payload[i] = (uint8_t)(i * 13 + 1);
    k_probe<<<1, 32, 0, s>>>(sh, dring, dstate, variant, n);
    CK(cudaGetLastError());
    std::vector<uint64_t> rtt(n);
    volatile uint32_t* flag = (volatile uint32_t*)(sh->ans + 1024);
    uint64_t bad = 0;
    for (uint32_t k = 1; k <= (uint32_t)n; k++) {
      const uint64_t t0 = now_ns();
      if (variant == 5) {
        for (int line = 17; line >= 0; line--) {
          uint32_t* w = (uint32_t*)(sh->cmd + 64 * line);
          for (int j = 0; j < 15; j++) w[j] = k * 31 + line + j;
          __atomic_store_n(&w[15], k, __ATOMIC_RELEASE);
        }
      } else {
        if (variant >= 2) sh->payload[0] = (uint8_t)k;
        __atomic_store_n((uint32_t*)sh->cmd, k, __ATOMIC_RELEASE);
      }
      while (*flag != k) __builtin_ia32_pause();
      if (variant == 4 || variant == 5 || variant == 8) {
        // integrity of the pushed payload: first and last stamped words
        const uint32_t w0 = *(volatile uint32_t*)(sh->ans);
        const uint32_t wl = *(volatile uint32_t*)(sh->ans + 512 + 16 * 31 + 12);
        if (w0 != k || wl != k) bad++;
      }
      rtt[k - 1] = now_ns() - t0;
    }
    CK(cudaStreamSynchronize(s));
    std::sort(rtt.begin(), rtt.end());
    printf("v%d  p50 %.2f us  p99 %.2f us  min %.2f us  torn %llu\n", variant, rtt[n / 2] / 1e3, rtt[n * 99 / 100] / 1e3,
           rtt[0] / 1e3, (unsigned long long)bad);
    fflush(stdout);
  }
  return 0;
}
