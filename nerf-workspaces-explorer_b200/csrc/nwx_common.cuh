// Shared helpers for libnwx (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <atomic>

#include "../../include/nwx.h"
#include "rng.cuh"

#define NWX_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return NWX_E_CUDA + (int)_e;      \
  } while (0)

#define NWX_REQUIRE(cond)                \
  do {                                   \
    if (!(cond)) return NWX_E_INVALID;   \
  } while (0)

extern std::atomic<int64_t> g_nwx_launches;   // defined in context.cu

// Every kernel launch goes through this so bench.py can report gpu_launches.
#define NWX_LAUNCHED()                                              \
  do {                                                              \
    g_nwx_launches.fetch_add(1, std::memory_order_relaxed);         \
    cudaError_t _e = cudaGetLastError();                            \
    if (_e != cudaSuccess) return NWX_E_CUDA + (int)_e;             \
  } while (0)

namespace nwx {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// Streaming (read-once) loads/stores that do not pollute L1.
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
// SM count of the CURRENT device (a process may drive several GPUs)
inline int num_sms() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (n[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}
// cudaFuncSetAttribute is per device: one flag per (kernel, device) instead of one per process
struct PerDeviceOnce {
  std::atomic<uint64_t> mask{0};
  bool need(int dev) const { return ((mask.load(std::memory_order_acquire) >> dev) & 1ull) == 0; }
  void done(int dev) { mask.fetch_or(1ull << dev, std::memory_order_release); }
};

// Internal launchers with an optional in-kernel random source (rng.on) in place of the tensors
int launch_coarse_z(const float* rays, int ray_dim, int64_t N, int S, const float* t_vals, const float* t_rand,
                    const RngSpec& rng, float* z_out, cudaStream_t st);
int launch_composite_fwd(const float* raw, const float* z, const float* rays_d, int d_stride, const float* noise,
                         const RngSpec& rng, int64_t N, int S, int white_bkgd, float* rgb, float* disp, float* acc,
                         float* depth, float* weights, int32_t* flags, cudaStream_t st, uint8_t* rgb8 = nullptr);
int launch_composite_bwd(const float* raw, const float* z, const float* rays_d, int d_stride, const float* noise,
                         const RngSpec& rng, const float* d_rgb, int64_t N, int S, int white_bkgd, float* d_raw,
                         cudaStream_t st);
int launch_sample_pdf(const float* z_c, const float* w_c, int Sc, const float* u, const RngSpec& rng, const float* u_lin,
                      int n_imp, int64_t N, float* z_samples, float* z_fine, int64_t* inds, float* z_std, cudaStream_t st,
                      float* z_samples_scratch = nullptr);   // z_samples == NULL: not wanted (scratch for kernels that need it)

// Embedding.embed on rows of stride x_stride (rays.cu)
int launch_embed(const float* x, int x_stride, int64_t P, int num_freqs, float scalar_factor, float* out,
                 cudaStream_t st);

}  // namespace nwx
