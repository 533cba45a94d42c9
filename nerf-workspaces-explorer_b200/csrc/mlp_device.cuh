// Device-side pieces shared by the fused forward kernel (mlp.cu) and the backward kernels (train.cu).
#pragma once
#include "mlp.cuh"
#include "sm100_ptx.cuh"

namespace nwx {

using namespace ptx;

constexpr int kTileM = 128;                     // points per tile (= TMEM lanes)
constexpr int kThreads = 512;
constexpr int kHBytes = kTileM * kHidden * 2;   // 65536: one activation tile, 4 K-blocks of 16 KB
constexpr int kABlock = kTileM * 64 * 2;        // 16384: one [128 x 64] bf16 K-block of A
constexpr uint64_t kWaitTimeoutNs = 10ull * 1000 * 1000 * 1000;   // wall-clock bound of one barrier wait

template <bool kPair, int kStages>
struct SmemLayout {
  static constexpr uint32_t kStageBytes = kPair ? kKBlockBytes / 2 : kKBlockBytes;
  static constexpr uint32_t h0 = 0;
  static constexpr uint32_t pe0 = 2 * kHBytes;
  static constexpr uint32_t w0 = pe0 + 2 * kABlock;
  static constexpr uint32_t bar0 = w0 + kStages * kStageBytes;
  // barrier slots (8 B each)
  static constexpr uint32_t w_full = bar0;
  static constexpr uint32_t w_empty = w_full + 8 * kStages;
  static constexpr uint32_t w_peer = w_empty + 8 * kStages;
  static constexpr uint32_t acc_full = w_peer + 8 * kStages;
  static constexpr uint32_t a_ready = acc_full + 16;
  static constexpr uint32_t pe_ready = a_ready + 16;
  static constexpr uint32_t pe_free = pe_ready + 16;
  static constexpr uint32_t tmem_slot = pe_free + 16;
  static constexpr uint32_t total = tmem_slot + 16;
  static constexpr uint32_t alloc_bytes = total + 1024;   // slack for manual 1024 B alignment
};

struct WaitCtx {
  uint32_t* diag;
  uint32_t code;
};

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Bounded by WALL CLOCK (a spin count is exceeded legitimately under compute-sanitizer, cuda-gdb or heavy
// co-tenancy): a protocol bug must not hang the GPU, so after kWaitTimeoutNs the waiter writes what it was
// waiting for into the context's host-mapped diagnostics word (always allocated, readable after the context is
// poisoned: nwx_ctx_last_diag) and aborts the grid.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, const WaitCtx& w) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFu) == 0) {                // each probe may park the warp for a while already
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kWaitTimeoutNs) {
        if (w.diag) {
          w.diag[0] = 0xDEAD0000u | w.code;
          w.diag[1] = blockIdx.x;
          w.diag[2] = bar;
          w.diag[3] = parity;
          __threadfence_system();
        }
        __trap();
      }
    }
  }
}

}  // namespace nwx
