// K1: pinhole ray generation from c2w poses + coarse (stratified) depth sampling.
//
// Replaces create_rays / _get_rays_camera / _get_rays_world (reference nerf/rays/rays.py:6-71)
// and the z_vals prologue of _volumetric_rendering (inference handler:216-220, training
// handler:547-562).  HBM-streaming kernels: 44 B written per ray, 256 B per ray of z.
//
// Bit-exactness: every fp32 operation is issued in the order torch-CPU executes it, with
// explicit round-to-nearest intrinsics so nvcc cannot contract mul+add into FMA where ATen
// does not:
//   d   = (R0*x + R1*y) + R2        -- ATen's small-matrix bmm loop, no FMA
//   |d| = sqrt(fma(dz,dz, fma(dy,dy, dx*dx)))   -- ATen's vectorised norm reduction uses FMA
// (both orders were determined against the reference's output, see DESIGN.md "bit-exact ops").
#include "nwx_common.cuh"

namespace nwx {

constexpr int kRaygenThreads = 256;

template <bool kViewDirs>
__global__ void __launch_bounds__(kRaygenThreads)
raygen_kernel(const float* __restrict__ c2w, int H, int W, float fx, float fy, float cx, float cy,
              float near, float far, int64_t ray0, int64_t nrays, float* __restrict__ out) {
  constexpr int kDim = kViewDirs ? 11 : 8;
  __shared__ float tile[kRaygenThreads * kDim];
  const int64_t base = (int64_t)blockIdx.x * kRaygenThreads;
  const int64_t local = base + threadIdx.x;
  if (local < nrays) {
    const int64_t ray = ray0 + local;
    const int64_t hw = (int64_t)H * W;
    const int b = (int)(ray / hw);
    const int pix = (int)(ray - (int64_t)b * hw);
    const int row = pix / W, col = pix - row * W;
    const float* T = c2w + (size_t)b * 16;
    const float x = __fdiv_rn(__fsub_rn((float)col, cx), fx);   // rays.py:52
    const float y = __fdiv_rn(__fsub_rn((float)row, cy), fy);   // rays.py:53
    float d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)                                 // rays.py:67 (z component is 1)
      d[k] = __fadd_rn(__fadd_rn(__fmul_rn(T[4 * k + 0], x), __fmul_rn(T[4 * k + 1], y)), T[4 * k + 2]);
    float* r = tile + threadIdx.x * kDim;
    r[0] = T[3]; r[1] = T[7]; r[2] = T[11];                     // rays.py:68
    r[3] = d[0]; r[4] = d[1]; r[5] = d[2];
    r[6] = near; r[7] = far;                                    // rays.py:26
    if (kViewDirs) {
      const float n = __fsqrt_rn(__fmaf_rn(d[2], d[2], __fmaf_rn(d[1], d[1], __fmul_rn(d[0], d[0]))));
      r[8] = __fdiv_rn(d[0], n); r[9] = __fdiv_rn(d[1], n); r[10] = __fdiv_rn(d[2], n);   // rays.py:24
    }
  }
  __syncthreads();
  // coalesced write-out of the block's contiguous [<=256, kDim] slab
  const int64_t remaining = nrays - base;
  const int count = (int)(remaining < kRaygenThreads ? remaining : kRaygenThreads) * kDim;
  float* dst = out + base * kDim;
  for (int i = threadIdx.x; i < count; i += kRaygenThreads) dst[i] = tile[i];
}

__global__ void __launch_bounds__(256)
coarse_z_kernel(const float* __restrict__ rays, int ray_dim, int64_t total, int S,
                const float* __restrict__ t_vals, const float* __restrict__ t_rand, const RngSpec rng,
                float* __restrict__ z_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t n = i / S;
    const int s = (int)(i - n * S);
    const float near = __ldg(rays + n * ray_dim + 6), far = __ldg(rays + n * ray_dim + 7);
    auto zlin = [&](int k) {                                    // inference handler:218
      const float t = __ldg(t_vals + k);
      return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, t)), __fmul_rn(far, t));
    };
    float z = zlin(s);
    if (t_rand != nullptr || rng.on) {                          // training handler:555-562
      const float lower = (s == 0) ? z : __fmul_rn(0.5f, __fadd_rn(z, zlin(s - 1)));
      const float upper = (s == S - 1) ? z : __fmul_rn(0.5f, __fadd_rn(zlin(s + 1), z));
      const float tr = t_rand ? ldg_stream(t_rand + i) : rng_uniform(rng, (uint64_t)i);
      z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), tr));
    }
    z_out[i] = z;
  }
}

// The render path's case (no jitter, S % 4 == 0, 16-byte aligned output): thread = four consecutive samples of a
// ray, one 128-bit store; the same correctly rounded expression per element as coarse_z_kernel.
__global__ void __launch_bounds__(256)
coarse_z_det4_kernel(const float* __restrict__ rays, int ray_dim, int64_t total4, int S4,
                     const float* __restrict__ t_vals, float4* __restrict__ z_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    const int64_t n = i / S4;
    const int s4 = (int)(i - n * S4);
    const float near = __ldg(rays + n * ray_dim + 6), far = __ldg(rays + n * ray_dim + 7);
    const float4 t = __ldg(reinterpret_cast<const float4*>(t_vals) + s4);
    auto zlin = [&](float tk) { return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, tk)), __fmul_rn(far, tk)); };
    z_out[i] = make_float4(zlin(t.x), zlin(t.y), zlin(t.z), zlin(t.w));
  }
}

__global__ void __launch_bounds__(256)
to8b_kernel(const float* __restrict__ x, int64_t n, uint8_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    // numpy: (255 * clip(x,0,1)).astype(uint8) -- fp32 multiply, truncation; NaN stays NaN -> 0
    float v = x[i];
    v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    out[i] = (uint8_t)(int)__fmul_rn(255.0f, v);
  }
}

// Embedding.embed (embedding.py:44-48) for callers that want the encoding materialised; the
// render path never does (the MLP kernel generates the features in registers).
__global__ void __launch_bounds__(256)
embed_kernel(const float* __restrict__ x, int x_stride, int64_t P, int L, float scale, float* __restrict__ out) {
  const int dim = 3 + 6 * L;
  const int64_t total = P * dim, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t p = i / dim;
    const int c = (int)(i - p * dim);
    const int a = c < 3 ? c : (c - 3) % 3;
    const float xs = __fdiv_rn(__ldg(x + p * x_stride + a), scale);
    float v = xs;
    if (c >= 3) {
      const int k = (c - 3) / 6;
      const float arg = __fmul_rn(xs, exp2f((float)k));
      v = ((c - 3) % 6 < 3) ? sinf(arg) : cosf(arg);
    }
    out[i] = v;
  }
}

int launch_embed(const float* x, int x_stride, int64_t P, int num_freqs, float scalar_factor, float* out,
                 cudaStream_t st) {
  if (P == 0) return NWX_OK;
  int64_t blocks = (P * (3 + 6 * num_freqs) + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  embed_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, x_stride, P, num_freqs, scalar_factor, out);
  NWX_LAUNCHED();
  return NWX_OK;
}

}  // namespace nwx

extern "C" int nwx_embed(const float* x, int64_t P, int num_freqs, float scalar_factor, float* out, void* stream) {
  NWX_REQUIRE(P >= 0 && num_freqs >= 0 && num_freqs <= 16 && scalar_factor != 0.0f);
  if (P == 0) return NWX_OK;
  NWX_REQUIRE(x && out);
  return nwx::launch_embed(x, 3, P, num_freqs, scalar_factor, out, (cudaStream_t)stream);
}

extern "C" int nwx_raygen(const float* c2w, int B, int H, int W, float fx, float fy, float cx, float cy,
                          float near, float far, int use_view_dirs, int64_t ray0, int64_t nrays,
                          float* rays_out, void* stream) {
  NWX_REQUIRE(B > 0 && H > 0 && W > 0 && ray0 >= 0 && nrays >= 0 && ray0 + nrays <= (int64_t)B * H * W);
  if (nrays == 0) return NWX_OK;                 // an empty shard is legal (and has a null output pointer)
  NWX_REQUIRE(c2w && rays_out);
  const unsigned grid = (unsigned)((nrays + nwx::kRaygenThreads - 1) / nwx::kRaygenThreads);
  auto st = (cudaStream_t)stream;
  if (use_view_dirs)
    nwx::raygen_kernel<true><<<grid, nwx::kRaygenThreads, 0, st>>>(c2w, H, W, fx, fy, cx, cy, near, far,
                                                                  ray0, nrays, rays_out);
  else
    nwx::raygen_kernel<false><<<grid, nwx::kRaygenThreads, 0, st>>>(c2w, H, W, fx, fy, cx, cy, near, far,
                                                                   ray0, nrays, rays_out);
  NWX_LAUNCHED();
  return NWX_OK;
}

int nwx::launch_coarse_z(const float* rays, int ray_dim, int64_t N, int S, const float* t_vals, const float* t_rand,
                         const RngSpec& rng, float* z_out, cudaStream_t st) {
  NWX_REQUIRE(ray_dim >= 8 && S >= 2 && N >= 0);
  if (N == 0) return NWX_OK;                     // empty shard: pointers may be null
  NWX_REQUIRE(rays && t_vals && z_out);
  const int64_t total = N * S;
  const int64_t cap = (int64_t)nwx::num_sms() * 16;
  if (t_rand == nullptr && !rng.on && S % 4 == 0 && ((uintptr_t)z_out & 15) == 0 && ((uintptr_t)t_vals & 15) == 0) {
    int64_t blocks = (total / 4 + 255) / 256;
    if (blocks > cap) blocks = cap;
    nwx::coarse_z_det4_kernel<<<(unsigned)blocks, 256, 0, st>>>(rays, ray_dim, total / 4, S / 4, t_vals,
                                                               reinterpret_cast<float4*>(z_out));
    NWX_LAUNCHED();
    return NWX_OK;
  }
  int64_t blocks = (total + 255) / 256;
  if (blocks > cap) blocks = cap;
  nwx::coarse_z_kernel<<<(unsigned)blocks, 256, 0, st>>>(rays, ray_dim, total, S, t_vals, t_rand, rng, z_out);
  NWX_LAUNCHED();
  return NWX_OK;
}

extern "C" int nwx_coarse_z(const float* rays, int ray_dim, int64_t N, int S, const float* t_vals,
                            const float* t_rand, float* z_out, void* stream) {
  return nwx::launch_coarse_z(rays, ray_dim, N, S, t_vals, t_rand, nwx::RngSpec{}, z_out, (cudaStream_t)stream);
}

// Fills out[i] with the library's counter-based draws (kind 0: U[0,1), 1: N(0,1)*scale) for element
// index i of (seed, offset, stream): what the kernels generate in place when no tensor is injected.
namespace nwx {
__global__ void __launch_bounds__(256)
rng_fill_kernel(int kind, RngSpec rng, int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = kind == 0 ? rng_uniform(rng, (uint64_t)i) : rng_normal(rng, (uint64_t)i);
}
}  // namespace nwx

extern "C" int nwx_rng_fill(int kind, uint64_t seed, uint64_t offset, uint32_t rng_stream, float scale, int64_t n,
                            float* out, void* stream) {
  NWX_REQUIRE((kind == 0 || kind == 1) && n >= 0);
  if (n == 0) return NWX_OK;
  NWX_REQUIRE(out);
  nwx::RngSpec r;
  r.seed = seed; r.offset = offset; r.stream = rng_stream; r.scale = scale; r.on = 1;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)nwx::num_sms() * 16) blocks = (int64_t)nwx::num_sms() * 16;
  nwx::rng_fill_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(kind, r, n, out);
  NWX_LAUNCHED();
  return NWX_OK;
}

// _sample_training_data (training handler:341-370) on the device: ONE random image of the bank and n pixel
// indices with replacement, then the gather of the sampled rays and ground-truth pixels.  Thread = one output
// float (coalesced writes); every thread of a ray re-derives the ray's pixel index from the counter-based
// generator (stream 5, element = sample number), the image index is element 0 of stream 4.
namespace nwx {
__device__ __forceinline__ uint32_t rng_below(const RngSpec& r, uint64_t idx, uint32_t n) {
  uint32_t x[4];
  philox4x32_10(r.seed, r.offset, r.stream, idx, x);
  return (uint32_t)(((uint64_t)x[0] * (uint64_t)n) >> 32);      // uniform in [0, n)
}
__global__ void __launch_bounds__(256)
sample_batch_kernel(const float* __restrict__ rays_bank, const float* __restrict__ rgb_bank, int num_img, uint32_t num_ray,
                    int ray_dim, int64_t n, RngSpec rng_img, RngSpec rng_pix, float* __restrict__ rays_out,
                    float* __restrict__ gt_out, int64_t* __restrict__ idx_out) {
  const int width = ray_dim + 3;
  const int64_t total = n * width, stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t img = rng_below(rng_img, 0, (uint32_t)num_img);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / width;
    const int c = (int)(i - r * width);
    const int64_t pix = rng_below(rng_pix, (uint64_t)r, num_ray);
    const int64_t src = img * num_ray + pix;
    if (c < ray_dim) rays_out[r * ray_dim + c] = __ldg(rays_bank + src * ray_dim + c);
    else gt_out[r * 3 + (c - ray_dim)] = __ldg(rgb_bank + src * 3 + (c - ray_dim));
    if (idx_out && c == 0) {
      idx_out[1 + r] = pix;
      if (r == 0) idx_out[0] = img;
    }
  }
}
}  // namespace nwx

extern "C" int nwx_sample_training_batch(const float* rays_bank, const float* rgb_bank, int num_img, int64_t num_ray,
                                         int ray_dim, int64_t n, uint64_t seed, uint64_t offset, float* rays_out,
                                         float* gt_out, int64_t* idx_out, void* stream) {
  NWX_REQUIRE(num_img >= 1 && num_ray >= 1 && num_ray <= 0xFFFFFFFFll && ray_dim >= 8 && n >= 0);
  if (n == 0) return NWX_OK;
  NWX_REQUIRE(rays_bank && rgb_bank && rays_out && gt_out);
  nwx::RngSpec ri, rp;
  ri.seed = rp.seed = seed; ri.offset = rp.offset = offset; ri.stream = 4; rp.stream = 5; ri.on = rp.on = 1;
  int64_t blocks = (n * (ray_dim + 3) + 255) / 256;
  const int64_t cap = (int64_t)nwx::num_sms() * 8;
  if (blocks > cap) blocks = cap;
  nwx::sample_batch_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      rays_bank, rgb_bank, num_img, (uint32_t)num_ray, ray_dim, n, ri, rp, rays_out, gt_out, idx_out);
  NWX_LAUNCHED();
  return NWX_OK;
}

extern "C" int nwx_to8b(const float* x, int64_t n, uint8_t* out, void* stream) {
  NWX_REQUIRE(n >= 0);
  if (n == 0) return NWX_OK;
  NWX_REQUIRE(x && out);
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)nwx::num_sms() * 16;
  if (blocks > cap) blocks = cap;
  nwx::to8b_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, out);
  NWX_LAUNCHED();
  return NWX_OK;
}
