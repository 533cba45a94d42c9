// K4: alpha compositing (raw2outputs, reference nerf/models/model_utils.py:33-100) and its
// analytic backward.  One warp per ray.
//   forward : lane l owns the K = ceil(S/32) CONSECUTIVE samples [l*K, l*K+K) of the ray, so the per-sample
//             chain (distance -> alpha -> local transmittance product) is lane-local and ONE fp64 warp scan per
//             ray (over the lanes' products) replaces a scan per 32-sample chunk; the lanes' K x 16 B runs tile
//             the ray's contiguous [S,4] block, so every fetched sector is used.  The uint8 image
//             (to8b_np, model_utils.py:9) is written by the same lane that writes rgb: no extra launch.
//   backward: sample s lives in lane s%32, chunk s/32 (coalesced 128 B / 512 B warp transactions), one warp
//             scan per 32-sample chunk with a running carry.
// HBM-bound on paper: 20 B read (+4 noise) and 4 B written per sample, 28 B per ray.
//
// Numerics follow torch-CPU op for op (no FMA contraction; expf/division correctly rounded
// variants).  torch.cumprod accumulates fp32 input in DOUBLE and rounds every output
// (ATen cpu_cum_base_kernel, acc_type<float,false>), so the scan runs in fp64 as well: the
// transmittance then matches the reference bit for bit except where a 1e-16 relative
// difference in the double product straddles an fp32 rounding boundary.
#include <cstdlib>
#include <cstring>

#include "mlp_device.cuh"   // mbarrier / bulk-copy wrappers and the bounded barrier wait

namespace nwx {

constexpr int kCompWarps = 8;   // warps (= rays) per block

__device__ __forceinline__ double warp_incl_prod(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double up = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v *= up;
  }
  return v;
}
__device__ __forceinline__ double warp_incl_sum(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double up = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += up;
  }
  return v;
}

__device__ __forceinline__ float sigmoidf_rn(float x) {        // torch.sigmoid: 1/(1+exp(-x))
  return __frcp_rn(__fadd_rn(1.0f, expf(-x)));                 // correctly rounded 1/y == correctly rounded reciprocal
}

// Forward colour path: sigmoid to ~1e-7 absolute (MUFU.EX2 + MUFU.RCP) instead of the correctly rounded
// expf + reciprocal (~35 instructions, three per sample: they made the forward instruction-bound).  Colours
// carry the 1e-3 tolerance of the bf16 MLP; the weights (alpha, transmittance), which decide the
// importance-sampling indices, keep the exact op-for-op arithmetic.  The backward keeps sigmoidf_rn.
__device__ __forceinline__ float sigmoidf_fast(float x) {      // FMUL, MUFU.EX2, FADD, MUFU.RCP
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}

// numpy (255 * clip(x, 0, 1)).astype(uint8): fp32 multiply, truncation; NaN -> 0  (model_utils.py:9)
__device__ __forceinline__ uint8_t to8b_one(float v) {
  v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
  return (uint8_t)(int)__fmul_rn(255.0f, v);
}

struct RaySample {
  float c[3];     // sigmoid(raw rgb)
  float alpha;    // 1 - exp(-relu(sigma+noise)*dist)
  float e;        // exp(-relu(sigma+noise)*dist) = 1 - alpha before rounding of the subtraction
  float t;        // 1 - alpha + 1e-10
  float dist;
  float z;
  bool pos;       // sigma + noise > 0
};

// Loads chunk j of a ray and evaluates the per-sample terms of model_utils.py:49-71.
template <int K>
__device__ __forceinline__ void load_ray(const float* __restrict__ raw, const float* __restrict__ z,
                                         const float* __restrict__ noise, const RngSpec& rng, int64_t ray, int S,
                                         int lane, float dnorm, RaySample (&sm)[K]) {
  float zr[K];
  float4 rw[K];
  float nz[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {                                 // issue every load before any math
    const int s = j * 32 + lane;
    const bool ok = s < S;
    const int64_t idx = ray * S + (ok ? s : S - 1);
    zr[j] = ldg_stream(z + idx);
    rw[j] = ldg_stream4(reinterpret_cast<const float4*>(raw) + idx);
    nz[j] = noise ? ldg_stream(noise + idx) : (rng.on ? rng_normal(rng, (uint64_t)idx) : 0.0f);
  }
  const bool noisy = noise != nullptr || rng.on;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int s = j * 32 + lane;
    float znext = __shfl_down_sync(kFull, zr[j], 1);
    const float zhead = __shfl_sync(kFull, zr[(j + 1 < K) ? j + 1 : j], 0);
    if (lane == 31) znext = zhead;
    float dist = (s >= S - 1) ? 1e10f : __fsub_rn(znext, zr[j]);          // :51,:56
    dist = __fmul_rn(dist, dnorm);                                        // :60
    const float sg = noisy ? __fadd_rn(rw[j].w, nz[j]) : rw[j].w;          // :71
    const float r = fmaxf(sg, 0.0f);
    const float e = expf(__fmul_rn(-r, dist));                            // :49
    RaySample& o = sm[j];
    o.pos = sg > 0.0f;
    o.e = e;
    o.alpha = __fsub_rn(1.0f, e);
    o.t = __fadd_rn(__fsub_rn(1.0f, o.alpha), 1e-10f);                    // :75
    o.c[0] = sigmoidf_rn(rw[j].x); o.c[1] = sigmoidf_rn(rw[j].y); o.c[2] = sigmoidf_rn(rw[j].z);  // :62
    o.dist = dist;
    o.z = zr[j];
    if (s >= S) { o.alpha = 0.0f; o.t = 1.0f; o.e = 1.0f; o.pos = false; }   // padding lanes: neutral
  }
}

__device__ __forceinline__ float ray_dnorm(const float* __restrict__ rays_d, int d_stride, int64_t ray) {
  const float dx = __ldg(rays_d + ray * d_stride + 0), dy = __ldg(rays_d + ray * d_stride + 1),
              dz = __ldg(rays_d + ray * d_stride + 2);
  return __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));   // torch.norm, :60
}

// Forward.  Lane-contiguous layout: lane l owns the K consecutive samples [l*K, l*K + K) of the ray, so the
// per-sample chain (dist -> alpha -> local transmittance product) is lane-local and ONE fp64 warp scan per ray
// (over the lanes' products) replaces one scan per 32-sample chunk.
//
// Two front ends feed the same arithmetic (composite_ray):
//   * bulk (default): the ray's contiguous [S,4] raw block and its S depths are brought into shared memory by TWO
//     1-D TMA bulk copies per ray (cp.async.bulk + mbarrier complete_tx, issued by lane 0), and the lanes read their
//     K x 16 B runs with LDS.128.  The next ray's copies are issued as soon as the registers are loaded, so they
//     fly underneath this ray's math.  HBM sees perfectly linear 3 KB + 768 B requests and the LSU no strided
//     global wavefronts: the direct version below needs 32 L1 wavefronts per LDG.128 (each lane in its own
//     128-byte line) -- ~210 per fine ray, which is what bounded it at 0.37 ms per frame.
//   * direct: per-lane __ldg of the same runs (any alignment, any S); also the A/B baseline (NWX_COMPOSITE=direct).
// LPR = lanes per ray: 32 (one ray per warp) or 16 (two rays per warp: the per-ray part -- one scan, five sums, the
// output tail -- is shared by both halves of the warp, which is what a 64-sample coarse pass is made of); `lane` is
// the lane's index within its ray's LPR lanes, `live` is false for the missing second ray of an odd tail.
template <int K, int LPR = 32>
__device__ __forceinline__ void composite_ray(const float4 (&rw)[K], const float (&zr)[K + 1], const float (&nz)[K], bool noisy,
                                              int64_t ray, int S, int lane, float dnorm,
                                              int white_bkgd, float* __restrict__ rgb, float* __restrict__ disp,
                                              float* __restrict__ acc, float* __restrict__ depth, float* __restrict__ weights,
                                              uint8_t* __restrict__ rgb8, int& bad, bool live = true) {
  const int s0 = lane * K;
  const int64_t base = ray * S;
  float alpha[K];
  double tloc[K];                                             // product of t over my samples before j
  double p = 1.0;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int s = s0 + j;
    float dist = (s >= S - 1) ? 1e10f : __fsub_rn(zr[j + 1], zr[j]);       // :51,:56
    dist = __fmul_rn(dist, dnorm);                                        // :60
    const float sg = noisy ? __fadd_rn(rw[j].w, nz[j]) : rw[j].w;          // :71
    float a = __fsub_rn(1.0f, expf(__fmul_rn(-fmaxf(sg, 0.0f), dist)));   // :49
    float t = __fadd_rn(__fsub_rn(1.0f, a), 1e-10f);                      // :75
    if (s >= S) { a = 0.0f; t = 1.0f; }                                   // padding samples: neutral
    alpha[j] = a;
    tloc[j] = p;
    p *= (double)t;
  }
  double incl = p;                                            // :75 cumprod (exclusive), fp64 like torch
#pragma unroll
  for (int o = 1; o < LPR; o <<= 1) {
    const double up = __shfl_up_sync(kFull, incl, o, LPR);
    if (lane >= o) incl *= up;
  }
  double excl = __shfl_up_sync(kFull, incl, 1, LPR);
  if (lane == 0) excl = 1.0;
  float a_r = 0.f, a_g = 0.f, a_b = 0.f, a_d = 0.f, a_w = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int s = s0 + j;
    const float T = (float)(excl * tloc[j]);
    const float w = __fmul_rn(alpha[j], T);
    if (s < S) {
      if (weights && live) weights[base + s] = w;
      a_r += __fmul_rn(w, sigmoidf_fast(rw[j].x));                        // :62,:84
      a_g += __fmul_rn(w, sigmoidf_fast(rw[j].y));
      a_b += __fmul_rn(w, sigmoidf_fast(rw[j].z));
      a_d += __fmul_rn(w, zr[j]);
      a_w += w;
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {                              // :84, :93, :95 (xor butterflies stay inside the ray's lanes)
    a_r += __shfl_xor_sync(kFull, a_r, o); a_g += __shfl_xor_sync(kFull, a_g, o); a_b += __shfl_xor_sync(kFull, a_b, o);
    a_d += __shfl_xor_sync(kFull, a_d, o); a_w += __shfl_xor_sync(kFull, a_w, o);
  }
  if (lane == 0 && live) {
    const float q = __fdiv_rn(a_d, a_w);                               // :94; 0/0 = NaN on empty rays and
    const float dspv = __fdiv_rn(1.0f, (q != q) ? q : fmaxf(1e-10f, q));   // torch.max propagates NaN
    if (white_bkgd) {                                                   // :98
      const float bg = __fsub_rn(1.0f, a_w);
      a_r = __fadd_rn(a_r, bg); a_g = __fadd_rn(a_g, bg); a_b = __fadd_rn(a_b, bg);
    }
    if (rgb) { rgb[ray * 3 + 0] = a_r; rgb[ray * 3 + 1] = a_g; rgb[ray * 3 + 2] = a_b; }
    if (rgb8) {                                                         // to8b_np, model_utils.py:9
      rgb8[ray * 3 + 0] = to8b_one(a_r); rgb8[ray * 3 + 1] = to8b_one(a_g); rgb8[ray * 3 + 2] = to8b_one(a_b);
    }
    if (disp) disp[ray] = dspv;
    if (acc) acc[ray] = a_w;
    if (depth) depth[ray] = a_d;
    // NaN / Inf screening (inference handler:273-275): the sum of the magnitudes is finite iff every value is
    const float mag = fabsf(a_r) + fabsf(a_g) + fabsf(a_b) + fabsf(dspv) + fabsf(a_w) + fabsf(a_d);
    if (!(mag < INFINITY)) {
      const float chk[6] = {a_r, a_g, a_b, dspv, a_w, a_d};
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        if (chk[c] != chk[c]) bad |= 1;
        else if (fabsf(chk[c]) == INFINITY) bad |= 2;
      }
    }
  }
}

template <int K>
__global__ void __launch_bounds__(kCompWarps * 32, 4)
composite_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                     const float* __restrict__ rays_d, int d_stride, const float* __restrict__ noise,
                     const RngSpec rng, int64_t N, int S, int white_bkgd, float* __restrict__ rgb, float* __restrict__ disp,
                     float* __restrict__ acc, float* __restrict__ depth, float* __restrict__ weights,
                     int32_t* __restrict__ flags, uint8_t* __restrict__ rgb8) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kCompWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kCompWarps;
  const bool noisy = noise != nullptr || rng.on;
  const int s0 = lane * K;
  int bad = 0;
  for (int64_t ray = warp0; ray < N; ray += nwarps) {
    const int64_t base = ray * S;
    float zr[K + 1], nz[K];
    float4 rw[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {                               // every load of the ray before any math
      const int s = s0 + j;
      const int64_t idx = base + (s < S ? s : S - 1);
      zr[j] = __ldg(z + idx);
      rw[j] = __ldg(reinterpret_cast<const float4*>(raw) + idx);
      nz[j] = noise ? __ldg(noise + idx) : (rng.on ? rng_normal(rng, (uint64_t)idx) : 0.0f);
    }
    zr[K] = __shfl_down_sync(kFull, zr[0], 1);                  // z of the sample after my last one
    composite_ray<K>(rw, zr, nz, noisy, ray, S, lane, ray_dnorm(rays_d, d_stride, ray), white_bkgd, rgb, disp, acc, depth,
                     weights, rgb8, bad);
  }
  if (flags && bad) atomicOr(flags, bad);
}

// Bulk front end (see above).  Requires S % 4 == 0 and 16-byte aligned raw / z (the launcher checks).
template <int K>
__global__ void __launch_bounds__(kCompWarps * 32, 4)
composite_fwd_bulk_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                          const float* __restrict__ rays_d, int d_stride, const float* __restrict__ noise,
                          const RngSpec rng, int64_t N, int S, int white_bkgd, float* __restrict__ rgb, float* __restrict__ disp,
                          float* __restrict__ acc, float* __restrict__ depth, float* __restrict__ weights,
                          int32_t* __restrict__ flags, uint8_t* __restrict__ rgb8) {
  __shared__ __align__(128) float4 s_raw[kCompWarps][K * 32];
  __shared__ __align__(16) float s_z[kCompWarps][K * 32];
  __shared__ __align__(8) uint64_t s_bar[kCompWarps];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * kCompWarps + wib;
  const int64_t nwarps = (int64_t)gridDim.x * kCompWarps;
  const bool noisy = noise != nullptr || rng.on;
  const int s0 = lane * K;
  const uint32_t bar = smem_u32(&s_bar[wib]), dst_raw = smem_u32(&s_raw[wib][0]), dst_z = smem_u32(&s_z[wib][0]);
  const uint32_t raw_bytes = (uint32_t)S * 16u, z_bytes = (uint32_t)S * 4u;
  const WaitCtx wc{nullptr, 0x5100u};
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  auto fetch = [&](int64_t ray) {                               // lane 0: two linear bulk copies, one barrier phase
    mbar_arrive_expect_tx(bar, raw_bytes + z_bytes);
    bulk_g2s(dst_raw, raw + ray * S * 4, raw_bytes, bar);
    bulk_g2s(dst_z, z + ray * S, z_bytes, bar);
  };
  if (warp0 < N && lane == 0) fetch(warp0);
  uint32_t phase = 0;
  int bad = 0;
  for (int64_t ray = warp0; ray < N; ray += nwarps) {
    const int64_t base = ray * S;
    float zr[K + 1], nz[K];
    float4 rw[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {                               // the noise (training) does not go through smem
      const int s = s0 + j;
      const int64_t idx = base + (s < S ? s : S - 1);
      nz[j] = noise ? __ldg(noise + idx) : (rng.on ? rng_normal(rng, (uint64_t)idx) : 0.0f);
    }
    const float dnorm = ray_dnorm(rays_d, d_stride, ray);       // issued before the wait: its latency hides behind it
    mbar_wait(bar, phase, wc);
    phase ^= 1u;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int s = s0 + j, sc = s < S ? s : S - 1;
      rw[j] = s_raw[wib][sc];
      zr[j] = s_z[wib][sc];
    }
    zr[K] = __shfl_down_sync(kFull, zr[0], 1);                  // z of the sample after my last one
    __syncwarp();                                               // every lane's reads of the buffer are ordered before ...
    if (lane == 0 && ray + nwarps < N) {
      fence_proxy_async_smem();                                 // ... the async proxy's writes into it (generic -> async)
      fetch(ray + nwarps);                                      // next ray's bytes fly underneath this ray's math
    }
    composite_ray<K>(rw, zr, nz, noisy, ray, S, lane, dnorm, white_bkgd, rgb, disp, acc, depth, weights, rgb8, bad);
  }
  if (flags && bad) atomicOr(flags, bad);
}

// Bulk front end, TWO rays per warp (16 lanes x K samples each) for short rays (S <= 64: the coarse pass).  The two rays
// are neighbours in memory, so one pair of bulk copies brings both.  Same arithmetic per ray as the kernel above (a
// lane's run of K samples starts at a different sample, so sums associate differently: outputs agree to rounding,
// weights -- a function of the fp64 scan -- are identical).
template <int K>
__global__ void __launch_bounds__(kCompWarps * 32, 4)
composite_fwd_bulk2_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                           const float* __restrict__ rays_d, int d_stride, const float* __restrict__ noise,
                           const RngSpec rng, int64_t N, int S, int white_bkgd, float* __restrict__ rgb, float* __restrict__ disp,
                           float* __restrict__ acc, float* __restrict__ depth, float* __restrict__ weights,
                           int32_t* __restrict__ flags, uint8_t* __restrict__ rgb8) {
  __shared__ __align__(128) float4 s_raw[kCompWarps][2 * K * 16];
  __shared__ __align__(16) float s_z[kCompWarps][2 * K * 16];
  __shared__ __align__(8) uint64_t s_bar[kCompWarps];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, half = lane >> 4, sl = lane & 15;
  const int64_t pairs = (N + 1) >> 1;
  const int64_t warp0 = (int64_t)blockIdx.x * kCompWarps + wib;
  const int64_t nwarps = (int64_t)gridDim.x * kCompWarps;
  const bool noisy = noise != nullptr || rng.on;
  const int s0 = sl * K;
  const uint32_t bar = smem_u32(&s_bar[wib]), dst_raw = smem_u32(&s_raw[wib][0]), dst_z = smem_u32(&s_z[wib][0]);
  const WaitCtx wc{nullptr, 0x5200u};
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  auto fetch = [&](int64_t pair) {                              // lane 0: both rays' bytes are contiguous in HBM
    const uint32_t n = (2 * pair + 1 < N) ? 2u : 1u;
    mbar_arrive_expect_tx(bar, n * (uint32_t)S * 20u);
    bulk_g2s(dst_raw, raw + 2 * pair * S * 4, n * (uint32_t)S * 16u, bar);
    bulk_g2s(dst_z, z + 2 * pair * S, n * (uint32_t)S * 4u, bar);
  };
  if (warp0 < pairs && lane == 0) fetch(warp0);
  uint32_t phase = 0;
  int bad = 0;
  for (int64_t pair = warp0; pair < pairs; pair += nwarps) {
    const int64_t ray_raw = 2 * pair + half;
    const bool live = ray_raw < N;
    const int64_t ray = live ? ray_raw : N - 1;                 // the missing twin of an odd tail recomputes its neighbour
    const int64_t base = ray * S;
    float zr[K + 1], nz[K];
    float4 rw[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int s = s0 + j;
      const int64_t idx = base + (s < S ? s : S - 1);
      nz[j] = noise ? __ldg(noise + idx) : (rng.on ? rng_normal(rng, (uint64_t)idx) : 0.0f);
    }
    const float dnorm = ray_dnorm(rays_d, d_stride, ray);
    mbar_wait(bar, phase, wc);
    phase ^= 1u;
    const int hoff = (live ? half : 0) * S;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int s = s0 + j, sc = s < S ? s : S - 1;
      rw[j] = s_raw[wib][hoff + sc];
      zr[j] = s_z[wib][hoff + sc];
    }
    zr[K] = __shfl_down_sync(kFull, zr[0], 1);                  // lane 15 / 31 get a foreign value: their last sample has dist 1e10
    __syncwarp();
    if (lane == 0 && pair + nwarps < pairs) {
      fence_proxy_async_smem();
      fetch(pair + nwarps);
    }
    composite_ray<K, 16>(rw, zr, nz, noisy, ray, S, sl, dnorm, white_bkgd, rgb, disp, acc, depth, weights, rgb8, bad, live);
  }
  if (flags && bad) atomicOr(flags, bad);
}

// NWX_COMPOSITE=direct forces the per-lane global loads (A/B measurements)
static bool composite_two_rays_enabled() {      // NWX_COMPOSITE=one_ray: one ray per warp also for short rays (A/B)
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("NWX_COMPOSITE");
    on = (e && strcmp(e, "one_ray") == 0) ? 0 : 1;
  }
  return on == 1;
}
static bool composite_bulk_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("NWX_COMPOSITE");
    on = (e && strcmp(e, "direct") == 0) ? 0 : 1;
  }
  return on == 1;
}

// Backward: d(loss)/d(raw) from d(loss)/d(rgb_map).  With w_i = alpha_i T_i, T_i = prod_{j<i} t_j,
// t_j = 1 - alpha_j + 1e-10 and G_i = g . c_i (minus sum(g) under a white background):
//   dL/dalpha_i = G_i T_i - (sum_{k>i} G_k w_k) / t_i
//   dL/dsigma_i = dL/dalpha_i * dist_i * exp(-relu(sigma_i) dist_i) * [sigma_i > 0]
//   dL/draw_c_i = w_i g_c * c_ic (1 - c_ic)
template <int K>
__global__ void __launch_bounds__(kCompWarps * 32)
composite_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                     const float* __restrict__ rays_d, int d_stride, const float* __restrict__ noise,
                     const RngSpec rng, const float* __restrict__ d_rgb, int64_t N, int S, int white_bkgd,
                     float* __restrict__ d_raw) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kCompWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kCompWarps;
  for (int64_t ray = warp0; ray < N; ray += nwarps) {
    RaySample sm[K];
    load_ray<K>(raw, z, noise, rng, ray, S, lane, ray_dnorm(rays_d, d_stride, ray), sm);
    const float g0 = __ldg(d_rgb + ray * 3 + 0), g1 = __ldg(d_rgb + ray * 3 + 1), g2 = __ldg(d_rgb + ray * 3 + 2);
    const float gbg = white_bkgd ? (g0 + g1 + g2) : 0.0f;
    float T[K], w[K], G[K];
    double pre[K];     // inclusive prefix of G_k w_k
    double carry = 1.0, run = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const double incl = warp_incl_prod((double)sm[j].t, lane);
      double excl = __shfl_up_sync(kFull, incl, 1);
      if (lane == 0) excl = 1.0;
      T[j] = (float)(carry * excl);
      carry *= __shfl_sync(kFull, incl, 31);
      w[j] = sm[j].alpha * T[j];
      G[j] = g0 * sm[j].c[0] + g1 * sm[j].c[1] + g2 * sm[j].c[2] - gbg;
      const double gw = (j * 32 + lane < S) ? (double)G[j] * (double)w[j] : 0.0;
      const double inc = warp_incl_sum(gw, lane);
      pre[j] = run + inc;
      run += __shfl_sync(kFull, inc, 31);
    }
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int s = j * 32 + lane;
      if (s < S) {
        const float suffix = (float)(run - pre[j]);
        const float dalpha = G[j] * T[j] - suffix / sm[j].t;
        const float dsig = sm[j].pos ? dalpha * sm[j].dist * sm[j].e : 0.0f;
        float4 o;
        o.x = w[j] * g0 * sm[j].c[0] * (1.0f - sm[j].c[0]);
        o.y = w[j] * g1 * sm[j].c[1] * (1.0f - sm[j].c[1]);
        o.z = w[j] * g2 * sm[j].c[2] * (1.0f - sm[j].c[2]);
        o.w = dsig;
        reinterpret_cast<float4*>(d_raw)[ray * S + s] = o;
      }
    }
  }
}

static inline unsigned comp_grid(int64_t N) {
  int64_t blocks = (N + kCompWarps - 1) / kCompWarps;
  const int64_t cap = (int64_t)num_sms() * 8;     // up to 8 blocks x 8 warps resident per SM
  return (unsigned)(blocks < cap ? blocks : cap);
}

}  // namespace nwx

#define NWX_DISPATCH_K(S, CALL)                       \
  switch (((S) + 31) / 32) {                          \
    case 1: { constexpr int K = 1; CALL; } break;     \
    case 2: { constexpr int K = 2; CALL; } break;     \
    case 3: { constexpr int K = 3; CALL; } break;     \
    case 4: { constexpr int K = 4; CALL; } break;     \
    case 5: { constexpr int K = 5; CALL; } break;     \
    case 6: { constexpr int K = 6; CALL; } break;     \
    case 7: { constexpr int K = 7; CALL; } break;     \
    case 8: { constexpr int K = 8; CALL; } break;     \
    default: return NWX_E_INVALID;                    \
  }

int nwx::launch_composite_fwd(const float* raw, const float* z, const float* rays_d, int d_stride, const float* noise,
                              const RngSpec& rng, int64_t N, int S, int white_bkgd, float* rgb, float* disp, float* acc,
                              float* depth, float* weights, int32_t* flags, cudaStream_t st, uint8_t* rgb8) {
  NWX_REQUIRE(d_stride >= 3 && S >= 1 && S <= 256 && N >= 0);
  if (N == 0) return NWX_OK;
  NWX_REQUIRE(raw && z && rays_d && (rgb || rgb8));
  const bool bulk = nwx::composite_bulk_enabled() && (S % 4) == 0 && ((reinterpret_cast<uintptr_t>(raw) & 15u) == 0) &&
                    ((reinterpret_cast<uintptr_t>(z) & 15u) == 0);
  if (bulk && S <= 64 && nwx::composite_two_rays_enabled()) {        // short rays: two per warp, K = ceil(S / 16) <= 4
    const unsigned grid = nwx::comp_grid((N + 1) / 2);
    switch ((S + 15) / 16) {
      case 1: nwx::composite_fwd_bulk2_kernel<1><<<grid, nwx::kCompWarps * 32, 0, st>>>(raw, z, rays_d, d_stride, noise, rng, N, S, white_bkgd, rgb, disp, acc, depth, weights, flags, rgb8); break;
      case 2: nwx::composite_fwd_bulk2_kernel<2><<<grid, nwx::kCompWarps * 32, 0, st>>>(raw, z, rays_d, d_stride, noise, rng, N, S, white_bkgd, rgb, disp, acc, depth, weights, flags, rgb8); break;
      case 3: nwx::composite_fwd_bulk2_kernel<3><<<grid, nwx::kCompWarps * 32, 0, st>>>(raw, z, rays_d, d_stride, noise, rng, N, S, white_bkgd, rgb, disp, acc, depth, weights, flags, rgb8); break;
      default: nwx::composite_fwd_bulk2_kernel<4><<<grid, nwx::kCompWarps * 32, 0, st>>>(raw, z, rays_d, d_stride, noise, rng, N, S, white_bkgd, rgb, disp, acc, depth, weights, flags, rgb8); break;
    }
  } else if (bulk) {
    NWX_DISPATCH_K(S, (nwx::composite_fwd_bulk_kernel<K><<<nwx::comp_grid(N), nwx::kCompWarps * 32, 0, st>>>(
                          raw, z, rays_d, d_stride, noise, rng, N, S, white_bkgd, rgb, disp, acc, depth, weights, flags, rgb8)));
  } else {
    NWX_DISPATCH_K(S, (nwx::composite_fwd_kernel<K><<<nwx::comp_grid(N), nwx::kCompWarps * 32, 0, st>>>(
                          raw, z, rays_d, d_stride, noise, rng, N, S, white_bkgd, rgb, disp, acc, depth, weights, flags, rgb8)));
  }
  NWX_LAUNCHED();
  return NWX_OK;
}

int nwx::launch_composite_bwd(const float* raw, const float* z, const float* rays_d, int d_stride, const float* noise,
                              const RngSpec& rng, const float* d_rgb, int64_t N, int S, int white_bkgd, float* d_raw,
                              cudaStream_t st) {
  NWX_REQUIRE(d_stride >= 3 && S >= 1 && S <= 256 && N >= 0);
  if (N == 0) return NWX_OK;
  NWX_REQUIRE(raw && z && rays_d && d_rgb && d_raw);
  NWX_DISPATCH_K(S, (nwx::composite_bwd_kernel<K><<<nwx::comp_grid(N), nwx::kCompWarps * 32, 0, st>>>(
                        raw, z, rays_d, d_stride, noise, rng, d_rgb, N, S, white_bkgd, d_raw)));
  NWX_LAUNCHED();
  return NWX_OK;
}

extern "C" int nwx_composite_fwd(const float* raw, const float* z, const float* rays_d, int d_stride,
                                 const float* noise, int64_t N, int S, int white_bkgd, float* rgb,
                                 float* disp, float* acc, float* depth, float* weights, int32_t* flags,
                                 void* stream) {
  NWX_REQUIRE(rgb || N == 0);
  return nwx::launch_composite_fwd(raw, z, rays_d, d_stride, noise, nwx::RngSpec{}, N, S, white_bkgd, rgb, disp, acc,
                                   depth, weights, flags, (cudaStream_t)stream, nullptr);
}

extern "C" int nwx_composite_bwd(const float* raw, const float* z, const float* rays_d, int d_stride,
                                 const float* noise, const float* weights, const float* d_rgb, int64_t N,
                                 int S, int white_bkgd, float* d_raw, void* stream) {
  (void)weights;   // recomputed from raw: cheaper than re-reading 4 B/sample and exact for alpha = 0
  return nwx::launch_composite_bwd(raw, z, rays_d, d_stride, noise, nwx::RngSpec{}, d_rgb, N, S, white_bkgd, d_raw,
                                   (cudaStream_t)stream);
}
