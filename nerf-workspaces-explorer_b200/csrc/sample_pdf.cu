// K2: hierarchical importance resampling (sample_pdf, reference nerf/rays/rays.py:74-121) fused
// with the coarse/fine depth merge sort(cat(z, z_samples)) (inference handler:243).
//
// One warp per ray.  Per ray the warp: stages z / weights in shared memory (coalesced loads),
// builds the 63-entry CDF, inverts it for 128 uniforms with binary searches over shared memory,
// and merges the sorted fine samples with the coarse depths by rank (merge-path), writing every
// output row coalesced.  1.3 KB of HBM traffic per ray.
//
// Bit-exact indices.  The searchsorted result depends on every bit of the CDF, so the CDF is
// built exactly as torch-CPU builds it (orders determined against the reference's output, see
// DESIGN.md "bit-exact ops"; tests/golden/sample_pdf.npz pins it):
//   * torch.sum over the 62 contiguous weights: ATen's AVX2 cascade (sum_stub is registered
//     without an AVX-512 variant, so every x86 host runs the 8-lane kernel): with 8-float
//     vectors v0..v6 and tail x56..x61:  p = v0+v4+v5+v6+v1+v2+v3 (lane-wise, in that order),
//     total = ((0 + x56 + ... + x61) + p[0]) + ... + p[7].  cascade_sum_emul() implements the
//     general-length rule (ilp 4, vec 8).
//   * pdf = w / total: correctly rounded fp32 division.
//   * torch.cumsum: fp32 input accumulated in DOUBLE, each output rounded to fp32.  pdf values
//     are >= 1e-5/1.0007 > 2^-17 and sum to ~1, so every partial sum is an exact multiple of
//     2^-40 below 2 and fits a double: any summation order gives the same doubles, and the warp
//     scan is bit-identical to the sequential loop.
//   * the lerp of rays.py:113-119 is evaluated op by op without FMA contraction.
#include <cstdlib>
#include <cstring>

#include "nwx_common.cuh"

namespace nwx {

constexpr int kPdfWarps = 8;
constexpr int kMaxBins = 128;     // M = Sc-1 <= 127
constexpr int kMaxImp = 256;

struct PdfSmem {
  float bins[kMaxBins];
  float cdf[kMaxBins];
  float x[kMaxBins];          // weights + 1e-5, then pdf
  float zc[kMaxBins];         // coarse depths
  float smp[kMaxImp];         // fine samples (sorted in place when u is random)
  float merged[kMaxBins + kMaxImp];
};

// torch.sum(x[0..n), dim=-1) on a contiguous fp32 row, n >= 8, bit-exact (see file header).
__device__ __forceinline__ float cascade_sum_emul(const float* x, int n, int lane) {
  const int V = n >> 3;            // full 8-float vectors
  const int ilp = V >> 2;          // rows of 4 vectors handled by multi_row_sum (< 16 for n < 512)
  float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
  if (lane < 8) {
    for (int i = 0; i < ilp; ++i) {
      p0 = __fadd_rn(p0, x[(4 * i + 0) * 8 + lane]);
      p1 = __fadd_rn(p1, x[(4 * i + 1) * 8 + lane]);
      p2 = __fadd_rn(p2, x[(4 * i + 2) * 8 + lane]);
      p3 = __fadd_rn(p3, x[(4 * i + 3) * 8 + lane]);
    }
    for (int i = 4 * ilp; i < V; ++i) p0 = __fadd_rn(p0, x[i * 8 + lane]);
    p0 = __fadd_rn(__fadd_rn(__fadd_rn(p0, p1), p2), p3);
  }
  float acc = 0.f;
  for (int k = V * 8; k < n; ++k) acc = __fadd_rn(acc, x[k]);       // scalar tail first
#pragma unroll
  for (int l = 0; l < 8; ++l) acc = __fadd_rn(acc, __shfl_sync(kFull, p0, l));
  return acc;                                                        // identical in every lane
}

__device__ __forceinline__ int upper_bound(const float* a, int n, float v) {   // #(a[i] <= v)
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int lower_bound(const float* a, int n, float v) {   // #(a[i] < v)
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// In-warp bitonic sort of a[0..n2) in shared memory, n2 a power of two <= 256.
__device__ __forceinline__ void warp_bitonic_sort(float* a, int n2, int lane) {
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n2 >> 1); t += 32) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // index with bit j clear
        const int p = i | j;
        const bool up = (i & k) == 0;
        const float lo = a[i], hi = a[p];
        if ((lo > hi) == up) { a[i] = hi; a[p] = lo; }
      }
      __syncwarp();
    }
  }
}

// kFromCoarse: inputs are (z_c, w_c) [N,Sc] and bins/weights are derived as the handlers do
// (inference handler:236-237); otherwise inputs are (bins [N,M], weights [N,M-1]) as in rays.py:74.
template <bool kFromCoarse>
__global__ void __launch_bounds__(kPdfWarps * 32)
sample_pdf_kernel(const float* __restrict__ in_a, const float* __restrict__ in_b, int Sc, int M,
                  const float* __restrict__ u, const RngSpec rng, const float* __restrict__ u_lin, int n_imp,
                  int64_t N, float* __restrict__ z_samples, float* __restrict__ z_fine, int64_t* __restrict__ inds_out,
                  float* __restrict__ z_std, float* __restrict__ cdf_out) {
  __shared__ PdfSmem smem[kPdfWarps];
  PdfSmem& sm = smem[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kPdfWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kPdfWarps;
  const int nw = M - 1;                                         // number of pdf weights
  const bool det = (u == nullptr) && !rng.on;
  // Deterministic uniforms are the same sorted array for every ray (linspace, rays.py:95): searchsorted of sorted
  // queries is then a merge.  For every CDF entry find the first uniform that reaches it (a guess from the
  // spacing plus an exact fix-up against the real values), histogram and prefix-sum: ind(q) = #{i : first(i) <= q}
  // -- bit-identical to upper_bound(cdf, u_q) for any sorted u, with 63 short searches instead of 128 long ones.
  bool fast = det;
  if (det) {
    for (int q = lane; q + 1 < n_imp; q += 32) fast = fast && (__ldg(u_lin + q) <= __ldg(u_lin + q + 1));
    fast = __all_sync(kFull, fast);
  }

  for (int64_t ray = warp0; ray < N; ray += nwarps) {
    // ---- stage inputs ----
    if (kFromCoarse) {
      for (int i = lane; i < Sc; i += 32) sm.zc[i] = ldg_stream(in_a + ray * Sc + i);
      for (int i = lane; i < nw; i += 32)
        sm.x[i] = __fadd_rn(ldg_stream(in_b + ray * Sc + i + 1), 1e-5f);            // rays.py:87, w[1:-1]
      __syncwarp();
      for (int i = lane; i < M; i += 32)
        sm.bins[i] = __fmul_rn(0.5f, __fadd_rn(sm.zc[i + 1], sm.zc[i]));            // handler:236
    } else {
      for (int i = lane; i < M; i += 32) sm.bins[i] = ldg_stream(in_a + ray * M + i);
      for (int i = lane; i < nw; i += 32) sm.x[i] = __fadd_rn(ldg_stream(in_b + ray * nw + i), 1e-5f);
    }
    __syncwarp();

    // ---- CDF (rays.py:88-90) ----
    const float total = cascade_sum_emul(sm.x, nw, lane);
    double carry = 0.0;
    if (lane == 0) sm.cdf[0] = 0.0f;
    for (int base = 0; base < nw; base += 32) {
      const int i = base + lane;
      double v = (i < nw) ? (double)__fdiv_rn(sm.x[i], total) : 0.0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += up;
      }
      v += carry;
      if (i < nw) sm.cdf[i + 1] = (float)v;
      carry = __shfl_sync(kFull, v, 31);
    }
    __syncwarp();
    if (cdf_out)
      for (int i = lane; i < M; i += 32) cdf_out[ray * M + i] = sm.cdf[i];

    int* ind_of = reinterpret_cast<int*>(sm.merged);                               // [n_imp], fast path only
    if (fast) {
      for (int q = lane; q < n_imp; q += 32) ind_of[q] = 0;
      __syncwarp();
      for (int i = lane; i < M; i += 32) {
        const float c = sm.cdf[i];
        int g = (int)ceilf(c * (float)(n_imp - 1));
        g = min(max(g, 0), n_imp);
        while (g > 0 && __ldg(u_lin + g - 1) >= c) --g;                              // first(i) = #{q : u_q < cdf[i]}
        while (g < n_imp && __ldg(u_lin + g) < c) ++g;
        if (g < n_imp) atomicAdd(&ind_of[g], 1);
      }
      __syncwarp();
      int carry_i = 0;
      for (int base = 0; base < n_imp; base += 32) {
        const int q = base + lane;
        int v = (q < n_imp) ? ind_of[q] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int up = __shfl_up_sync(kFull, v, o);
          if (lane >= o) v += up;
        }
        v += carry_i;
        if (q < n_imp) ind_of[q] = v;
        carry_i = __shfl_sync(kFull, v, 31);
      }
      __syncwarp();
    }

    // ---- invert (rays.py:103-119) ----
    double s1 = 0.0;
    bool sorted = true;
    for (int base = 0; base < n_imp; base += 32) {
      const int q = base + lane;
      float smp = 0.f;
      if (q < n_imp) {
        const float uu = det ? __ldg(u_lin + q)
                             : (u ? ldg_stream(u + ray * n_imp + q) : rng_uniform(rng, (uint64_t)(ray * n_imp + q)));
        const int ind = fast ? ind_of[q] : upper_bound(sm.cdf, M, uu);                // :103 right=True
        const int below = max(ind - 1, 0), above = min(ind, M - 1);                   // :104-105
        const float cb = sm.cdf[below], ca = sm.cdf[above];
        const float bb = sm.bins[below], ba = sm.bins[above];
        float denom = __fsub_rn(ca, cb);                                              // :113
        if (denom < 1e-5f) denom = 1.0f;                                              // :114
        const float t = __fdiv_rn(__fsub_rn(uu, cb), denom);                          // :118
        smp = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));                         // :119
        sm.smp[q] = smp;
        z_samples[ray * n_imp + q] = smp;
        if (inds_out) inds_out[ray * n_imp + q] = ind;
        s1 += smp;
      }
      // sortedness probe (needed only to skip the sort): compare with the previous element
      float prev = __shfl_up_sync(kFull, smp, 1);
      if (lane == 0) prev = (base == 0) ? smp : sm.smp[base - 1];
      if (q < n_imp && !(prev <= smp)) sorted = false;
      __syncwarp();                              // sm.smp[base+31] is read by lane 0 next round
    }
    __syncwarp();

    if (z_std) {                                                                      // handler:267
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(kFull, s1, o);
      const double mean = s1 / n_imp;
      double s2 = 0.0;
      for (int q = lane; q < n_imp; q += 32) { const double dlt = sm.smp[q] - mean; s2 += dlt * dlt; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(kFull, s2, o);
      if (lane == 0) z_std[ray] = (float)sqrt(s2 / n_imp);
    }

    // ---- merge = torch.sort(cat([z_c, z_samples])) values (handler:243) ----
    const bool all_sorted = __all_sync(kFull, sorted);
    if (kFromCoarse && z_fine && fast && all_sorted) {
      // Sorted samples whose CDF bin is known: a sample drawn from bin [below, above] lies between the mid-points
      // around z_c[above], so its rank among the coarse depths is above or above + 1 -- start there and fix up
      // exactly.  The coarse depths then fill the slots the samples left free, in order (coarse first on ties,
      // as torch.sort of cat([z_c, z_samples]) with distinct keys; equal keys carry equal values).
      uint32_t* occ = reinterpret_cast<uint32_t*>(sm.x);        // pdf values are dead: <= 12 occupancy words
      const int tot = Sc + n_imp, nwords = (tot + 31) >> 5;
      if (lane < nwords) occ[lane] = 0u;
      __syncwarp();
      for (int q = lane; q < n_imp; q += 32) {
        const float v = sm.smp[q];
        int r = min(ind_of[q], M - 1);
        while (r < Sc && sm.zc[r] <= v) ++r;                    // r = #{j : z_c[j] <= v}
        while (r > 0 && sm.zc[r - 1] > v) --r;
        const int pos = q + r;
        atomicOr(&occ[pos >> 5], 1u << (pos & 31));
      }
      __syncwarp();
      const uint32_t myw = (lane < nwords) ? occ[lane] : 0u;
      int incl = __popc(myw);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += up;
      }
      const int excl = incl - __popc(myw);
      for (int w = 0; w < nwords; ++w) {
        const uint32_t word = __shfl_sync(kFull, myw, w);
        const int k = __shfl_sync(kFull, excl, w) + __popc(word & ((1u << lane) - 1u));   // samples before this slot
        const int slot = 32 * w + lane;
        if (slot < tot) z_fine[ray * tot + slot] = ((word >> lane) & 1u) ? sm.smp[k] : sm.zc[slot - k];
      }
    } else if (kFromCoarse && z_fine) {
      if (!all_sorted) {                         // random u (training): sort the samples first
        int n2 = 1;
        while (n2 < n_imp) n2 <<= 1;
        for (int q = n_imp + lane; q < n2; q += 32) sm.smp[q] = INFINITY;
        __syncwarp();
        warp_bitonic_sort(sm.smp, n2, lane);
      }
      for (int i = lane; i < Sc; i += 32) {
        const float v = sm.zc[i];
        sm.merged[i + lower_bound(sm.smp, n_imp, v)] = v;       // coarse first on ties
      }
      for (int q = lane; q < n_imp; q += 32) {
        const float v = sm.smp[q];
        sm.merged[q + upper_bound(sm.zc, Sc, v)] = v;
      }
      __syncwarp();
      const int tot = Sc + n_imp;
      for (int i = lane; i < tot; i += 32) z_fine[ray * tot + i] = sm.merged[i];
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Specialisation for the shape every shipped config renders with (office_*_config.yaml:22-23: 64 coarse
// depths -> 63 bins, 128 deterministic uniforms).  Same arithmetic, same bits as the generic kernel above,
// ~3x fewer instructions: every loop bound is a compile-time constant, a lane owns CONTIGUOUS elements
// (the pair 2l, 2l+1 of the 64-wide arrays, the quad 4l..4l+3 of the 128 samples) so each prefix sum is a
// lane-local sum plus ONE warp scan, inputs arrive as 8-byte coalesced loads, the CDF and the bins sit
// interleaved in shared memory (one LDS.64 per gather), the uniforms are staged once per block, and the
// sample store is skipped when nobody asked for z_samples (the render only consumes the merged depths).
// ------------------------------------------------------------------------------------------------
struct __align__(16) PdfSmem64 {
  float2 cb[64];            // {cdf[i], bins[i]}, i < 63
  float x[64];              // weights + 1e-5 (62 used)
  float zc[64];             // coarse depths (kFromCoarse)
  float smp[128];           // fine samples
  int cnt[128];             // histogram of first(i), then free
  uint32_t occ[8];          // occupancy bit mask of the 192 merged slots
};

template <bool kFromCoarse>
__global__ void __launch_bounds__(kPdfWarps * 32)
sample_pdf_det64_kernel(const float* __restrict__ in_a, const float* __restrict__ in_b, const float* __restrict__ u_lin,
                        int64_t N, float* __restrict__ z_samples, float* __restrict__ z_fine,
                        int64_t* __restrict__ inds_out, float* __restrict__ z_std, float* __restrict__ cdf_out) {
  constexpr int Sc = 64, M = 63, nw = 62, NI = 128, TOT = Sc + NI;
  __shared__ PdfSmem64 smem[kPdfWarps];
  __shared__ __align__(16) float s_u[NI];
  PdfSmem64& sm = smem[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kPdfWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kPdfWarps;
  for (int q = threadIdx.x; q < NI; q += blockDim.x) s_u[q] = __ldg(u_lin + q);
  __syncthreads();
  // the histogram inversion needs sorted uniforms (torch.linspace is); anything else takes binary searches
  bool u_sorted = true;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int q = 4 * lane + j;
    if (q + 1 < NI) u_sorted = u_sorted && (s_u[q] <= s_u[q + 1]);
  }
  u_sorted = __all_sync(kFull, u_sorted);
  const float4 u4 = reinterpret_cast<const float4*>(s_u)[lane];
  const float uq[4] = {u4.x, u4.y, u4.z, u4.w};

  for (int64_t ray = warp0; ray < N; ray += nwarps) {
    // ---- stage: lane owns elements 2l, 2l+1 ----
    float b0, b1 = 0.f, x0 = 0.f, x1 = 0.f;
    if (kFromCoarse) {
      const float2 zz = __ldcs(reinterpret_cast<const float2*>(in_a + ray * Sc) + lane);
      const float2 ww = __ldcs(reinterpret_cast<const float2*>(in_b + ray * Sc) + lane);
      reinterpret_cast<float2*>(sm.zc)[lane] = zz;
      const float z_next = __shfl_down_sync(kFull, zz.x, 1), w_next = __shfl_down_sync(kFull, ww.x, 1);
      b0 = __fmul_rn(0.5f, __fadd_rn(zz.y, zz.x));                                    // handler:236, bins[2l]
      if (lane < 31) {
        b1 = __fmul_rn(0.5f, __fadd_rn(z_next, zz.y));                                // bins[2l+1]
        x0 = __fadd_rn(ww.y, 1e-5f);                                                  // rays.py:87 on w[1:-1]: x[2l] = w[2l+1]
        x1 = __fadd_rn(w_next, 1e-5f);                                                // x[2l+1] = w[2l+2]
      }
    } else {
      b0 = ldg_stream(in_a + ray * M + 2 * lane);
      if (lane < 31) {
        b1 = ldg_stream(in_a + ray * M + 2 * lane + 1);
        x0 = __fadd_rn(ldg_stream(in_b + ray * nw + 2 * lane), 1e-5f);
        x1 = __fadd_rn(ldg_stream(in_b + ray * nw + 2 * lane + 1), 1e-5f);
      }
    }
    reinterpret_cast<float2*>(sm.x)[lane] = make_float2(x0, x1);
    reinterpret_cast<int4*>(sm.cnt)[lane] = make_int4(0, 0, 0, 0);
    if (lane < 8) sm.occ[lane] = 0u;
    __syncwarp();

    // ---- CDF (rays.py:88-90): cascade sum, correctly rounded division, exact fp64 prefix ----
    const float total = cascade_sum_emul(sm.x, nw, lane);
    const double d0 = lane < 31 ? (double)__fdiv_rn(x0, total) : 0.0, d1 = lane < 31 ? (double)__fdiv_rn(x1, total) : 0.0;
    double incl = d0 + d1;                                       // all partial sums are exact in double (file header)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += up;
    }
    const float c1 = (float)(incl - d1), c2 = (float)incl;       // cdf[2l+1], cdf[2l+2]  (lane < 31)
    float* cbf = reinterpret_cast<float*>(sm.cb);
    cbf[2 * (2 * lane) + 1] = b0;
    if (lane == 0) cbf[0] = 0.0f;                                 // cdf[0]
    if (lane < 31) {
      cbf[2 * (2 * lane + 1) + 1] = b1;
      cbf[2 * (2 * lane + 1)] = c1;
      cbf[2 * (2 * lane + 2)] = c2;
    }
    if (cdf_out && lane < 31) {
      if (lane == 0) cdf_out[ray * M] = 0.0f;
      cdf_out[ray * M + 2 * lane + 1] = c1;
      cdf_out[ray * M + 2 * lane + 2] = c2;
    }

    // ---- searchsorted(cdf, u, right=True) (rays.py:103) ----
    int ind[4];
    if (u_sorted) {
      // first(i) = #{q : u_q < cdf[i]}; ind(q) = #{i : first(i) <= q}; cdf[0] = 0 <= every u contributes the 1
      if (lane < 31) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float c = e == 0 ? c1 : c2;
          int g = (int)ceilf(c * (float)(NI - 1));
          g = min(max(g, 0), NI);
          while (g > 0 && s_u[g - 1] >= c) --g;
          while (g < NI && s_u[g] < c) ++g;
          if (g < NI) atomicAdd(&sm.cnt[g], 1);
        }
      }
      __syncwarp();
      const int4 c4 = reinterpret_cast<const int4*>(sm.cnt)[lane];
      const int a0 = c4.x, a1 = a0 + c4.y, a2 = a1 + c4.z, a3 = a2 + c4.w;
      int scan = a3;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(kFull, scan, o);
        if (lane >= o) scan += up;
      }
      const int base = scan - a3 + 1;
      ind[0] = base + a0; ind[1] = base + a1; ind[2] = base + a2; ind[3] = base + a3;
    } else {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int lo = 0, hi = M;                                      // #(cdf[i] <= u)
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (sm.cb[mid].x <= uq[j]) lo = mid + 1; else hi = mid;
        }
        ind[j] = lo;
      }
    }

    // ---- invert (rays.py:104-119): lane owns samples 4l..4l+3 ----
    float smp[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int below = max(ind[j] - 1, 0), above = min(ind[j], M - 1);               // :104-105
      const float2 lo = sm.cb[below], hi = sm.cb[above];
      float denom = __fsub_rn(hi.x, lo.x);                                            // :113
      if (denom < 1e-5f) denom = 1.0f;                                                // :114
      const float t = __fdiv_rn(__fsub_rn(uq[j], lo.x), denom);                       // :118
      smp[j] = __fadd_rn(lo.y, __fmul_rn(t, __fsub_rn(hi.y, lo.y)));                  // :119
    }
    reinterpret_cast<float4*>(sm.smp)[lane] = make_float4(smp[0], smp[1], smp[2], smp[3]);
    if (z_samples) reinterpret_cast<float4*>(z_samples + ray * NI)[lane] = make_float4(smp[0], smp[1], smp[2], smp[3]);
    if (inds_out) {
      longlong2* io = reinterpret_cast<longlong2*>(inds_out + ray * NI + 4 * lane);
      io[0] = make_longlong2(ind[0], ind[1]);
      io[1] = make_longlong2(ind[2], ind[3]);
    }
    if (z_std) {                                                                      // handler:267
      double s1 = ((double)smp[0] + (double)smp[1]) + ((double)smp[2] + (double)smp[3]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(kFull, s1, o);
      const double mean = s1 / NI;
      double s2 = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) { const double dlt = smp[j] - mean; s2 += dlt * dlt; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(kFull, s2, o);
      if (lane == 0) z_std[ray] = (float)sqrt(s2 / NI);
    }

    // ---- merge = torch.sort(cat([z_c, z_samples])) values (handler:243) ----
    if (kFromCoarse && z_fine) {
      float prev = __shfl_up_sync(kFull, smp[3], 1);
      if (lane == 0) prev = smp[0];
      const bool sorted = __all_sync(kFull, prev <= smp[0] && smp[0] <= smp[1] && smp[1] <= smp[2] && smp[2] <= smp[3]);
      __syncwarp();                                              // sm.smp complete
      if (sorted) {
        // a sample drawn from bin [below, above] lies between the mid-points around z_c[above]: its rank among the
        // coarse depths is `above` or `above + 1` -- start there and fix up exactly; coarse first on ties
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v = smp[j];
          int r = min(ind[j], M - 1);
          while (r < Sc && sm.zc[r] <= v) ++r;                   // r = #{k : z_c[k] <= v}
          while (r > 0 && sm.zc[r - 1] > v) --r;
          const int pos = 4 * lane + j + r;
          atomicOr(&sm.occ[pos >> 5], 1u << (pos & 31));
        }
        __syncwarp();
        int before = 0;                                          // samples in the words before this one
        const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
        for (int w = 0; w < TOT / 32; ++w) {
          const uint32_t word = sm.occ[w];
          const int k = before + __popc(word & lt);
          const int slot = 32 * w + lane;
          z_fine[ray * TOT + slot] = ((word >> lane) & 1u) ? sm.smp[k] : sm.zc[slot - k];
          before += __popc(word);
        }
      } else {
        // a rounding inversion between neighbouring samples (possible, never seen): sort, then merge by rank
        warp_bitonic_sort(sm.smp, NI, lane);
        for (int i = lane; i < Sc; i += 32) {
          const float v = sm.zc[i];
          z_fine[ray * TOT + i + lower_bound(sm.smp, NI, v)] = v;   // coarse first on ties
        }
        for (int q = lane; q < NI; q += 32) {
          const float v = sm.smp[q];
          z_fine[ray * TOT + q + upper_bound(sm.zc, Sc, v)] = v;
        }
      }
    }
    __syncwarp();                                                // smem is reused by the next ray
  }
}

template <bool kFromCoarse>
static inline unsigned pdf64_grid(int64_t N) {
  static int per_sm = 0;
  if (per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sample_pdf_det64_kernel<kFromCoarse>, kPdfWarps * 32, 0) != cudaSuccess ||
        per_sm < 1)
      per_sm = 4;
  }
  int64_t blocks = (N + kPdfWarps - 1) / kPdfWarps;
  const int64_t cap = (int64_t)num_sms() * per_sm;
  return (unsigned)(blocks < cap ? blocks : cap);
}

// NWX_SAMPLE_PDF=generic forces the generic kernel (A/B measurements, cross-checks in the tests)
static bool pdf_specialised_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("NWX_SAMPLE_PDF");
    on = (e && strcmp(e, "generic") == 0) ? 0 : 1;
  }
  return on == 1;
}
static inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// One resident wave: the kernel is persistent (warps stride over the rays), and 37 KB of shared memory per block
// means 6 blocks fit on an SM, not 8 -- a grid of 8 per SM ran 1.33 waves with a third of the machine idle at the end.
template <bool kFromCoarse>
static inline unsigned pdf_grid(int64_t N) {
  static int per_sm = 0;
  if (per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sample_pdf_kernel<kFromCoarse>, kPdfWarps * 32, 0) != cudaSuccess ||
        per_sm < 1)
      per_sm = 4;
  }
  int64_t blocks = (N + kPdfWarps - 1) / kPdfWarps;
  const int64_t cap = (int64_t)num_sms() * per_sm;
  return (unsigned)(blocks < cap ? blocks : cap);
}

}  // namespace nwx

int nwx::launch_sample_pdf(const float* z_c, const float* w_c, int Sc, const float* u, const RngSpec& rng,
                           const float* u_lin, int n_imp, int64_t N, float* z_samples, float* z_fine, int64_t* inds,
                           float* z_std, cudaStream_t st, float* z_samples_scratch) {
  NWX_REQUIRE(N >= 0 && Sc >= 11 && Sc <= nwx::kMaxBins && n_imp >= 1 && n_imp <= nwx::kMaxImp);
  if (N == 0) return NWX_OK;
  NWX_REQUIRE(z_c && w_c && (z_samples || z_fine) && (u || u_lin || rng.on));
  if (Sc == 64 && n_imp == 128 && !u && !rng.on && nwx::pdf_specialised_enabled() && nwx::aligned8(z_c) &&
      nwx::aligned8(w_c) && (!z_samples || nwx::aligned16(z_samples)) && (!inds || nwx::aligned16(inds))) {
    nwx::sample_pdf_det64_kernel<true><<<nwx::pdf64_grid<true>(N), nwx::kPdfWarps * 32, 0, st>>>(
        z_c, w_c, u_lin, N, z_samples, z_fine, inds, z_std, nullptr);
    NWX_LAUNCHED();
    return NWX_OK;
  }
  if (!z_samples) z_samples = z_samples_scratch;    // the generic kernel always materialises the samples
  NWX_REQUIRE(z_samples);
  nwx::sample_pdf_kernel<true><<<nwx::pdf_grid<true>(N), nwx::kPdfWarps * 32, 0, st>>>(
      z_c, w_c, Sc, Sc - 1, u, rng, u_lin, n_imp, N, z_samples, z_fine, inds, z_std, nullptr);
  NWX_LAUNCHED();
  return NWX_OK;
}

extern "C" int nwx_sample_pdf(const float* z_c, const float* w_c, int Sc, const float* u, const float* u_lin,
                              int n_imp, int64_t N, float* z_samples, float* z_fine, int64_t* inds,
                              float* z_std, void* stream) {
  NWX_REQUIRE(z_samples || N == 0);
  return nwx::launch_sample_pdf(z_c, w_c, Sc, u, nwx::RngSpec{}, u_lin, n_imp, N, z_samples, z_fine, inds, z_std,
                                (cudaStream_t)stream, nullptr);
}

extern "C" int nwx_sample_pdf_bins(const float* bins, const float* weights, int M, const float* u,
                                   const float* u_lin, int n_imp, int64_t N, float* samples, int64_t* inds,
                                   float* cdf_out, void* stream) {
  NWX_REQUIRE(N >= 0 && M >= 10 && M <= nwx::kMaxBins && n_imp >= 1 && n_imp <= nwx::kMaxImp);
  if (N == 0) return NWX_OK;
  NWX_REQUIRE(bins && weights && samples && (u || u_lin));
  if (M == 63 && n_imp == 128 && !u && nwx::pdf_specialised_enabled() && nwx::aligned16(samples) &&
      (!inds || nwx::aligned16(inds))) {
    nwx::sample_pdf_det64_kernel<false><<<nwx::pdf64_grid<false>(N), nwx::kPdfWarps * 32, 0, (cudaStream_t)stream>>>(
        bins, weights, u_lin, N, samples, nullptr, inds, nullptr, cdf_out);
    NWX_LAUNCHED();
    return NWX_OK;
  }
  nwx::sample_pdf_kernel<false><<<nwx::pdf_grid<false>(N), nwx::kPdfWarps * 32, 0, (cudaStream_t)stream>>>(
      bins, weights, 0, M, u, nwx::RngSpec{}, u_lin, n_imp, N, samples, nullptr, inds, nullptr, cdf_out);
  NWX_LAUNCHED();
  return NWX_OK;
}
