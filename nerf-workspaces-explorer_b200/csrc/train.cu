// Training backward of the fused NeRF MLP (reference: autograd through NeRFModel.forward,
// nerf/models/nerf_model.py:45-83, as driven by training handler:277-315) and the small kernels
// of the optimisation step (MSE loss gradient, head gradients, Adam).
//
// The training forward (mlp.cu, kTrain; the views layer folded like at inference) leaves every tensor-core
// operand in HBM as the same [128 points x 64 features] bf16 swizzled tile images it built in shared memory,
// plus one ReLU' bit word per row and 32-column chunk.  Backward:
//
//   dX kernel  (mlp_bwd_dx_kernel): same persistent CTA-pair / two-tiles-in-flight skeleton as the
//       forward.  Per tile it walks the layers backwards: G_views from the rgb head (CUDA cores),
//       then 8 tcgen05 GEMM steps  dH_in = G_out . W  with transposed-weight K-block images streamed
//       by bulk TMA (step 0: the folded views layer); each epilogue applies the ReLU' bit mask and writes
//       the new G tile to smem (next step's A operand) and to HBM (for dW).
//   dW kernel  (mlp_bwd_dw_kernel): job-major.  dW[out,in] = sum_points G[p,out] X[p,in] has the
//       points as the reduction dimension, which is the ROW dimension of the saved images, so both
//       operands are fed to tcgen05.mma as MN-major tiles (same bytes, a_major = b_major = 1): no
//       transposes anywhere.  The 10 (G, X) pairs are dealt to the CTAs; each CTA accumulates ONE dW in
//       TMEM over its share of the tiles and writes one partial; bias gradients are column sums of the
//       G tiles taken from shared memory by otherwise idle warps.  A small kernel reduces the partials.
//   fold_sum / fold_grads_*: chain rule from d W_fold back to d W_view[:, :256], d W_feature, d b_feature.
//   head_{rgb,dir,sigma}_kernel: rgb / sigma heads and the 27 view-direction columns of the views layer (fp32).
//   No atomics anywhere: gradients and losses are bitwise reproducible.
#include "mlp_device.cuh"

namespace nwx {

// ------------------------------------------------------------------------------------------------
// transposed weight images for dX:  dH_in[p][i] = sum_o G[p][o] W[o][i]  =>  B[n=i][k=o] = W[o][i]
// steps: 0 = folded views layer W_fold = W_view[:, :256] . W_feature (K = 128 -> 2 K-blocks, written by
// pack_fold_kernel), 1..7 = pts layers 7..1 (4 K-blocks each).  The training forward runs the folded layer, so
// the backward differentiates exactly that function; the chain rule back to W_view / W_feature / b_feature is
// applied to the accumulated d W_fold / d b_fold by fold_grads_*_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kDxSteps = 8;
constexpr int kDxKBlocks = 2 + 7 * 4;     // 30 images of [256 x 64]
__host__ __device__ constexpr int dx_step_nkb(int s) { return s == 0 ? 2 : 4; }
__host__ __device__ constexpr int dx_step_kb0(int s) { return s == 0 ? 0 : 2 + 4 * (s - 1); }

struct PackTSrc {
  const float* w[kDxSteps];   // (unused: fold), pts7, pts6, pts5, pts4, pts3, pts2, pts1
};

__global__ void pack_weights_t_kernel(PackTSrc src, uint8_t* __restrict__ wimg_t) {
  const int g = blockIdx.x + 2;             // K-blocks 0, 1 (step 0) come from pack_fold_kernel
  int s = 0, kb = g;
  while (kb >= dx_step_nkb(s)) { kb -= dx_step_nkb(s); ++s; }
  // source tensor [out, in_dim]; dX needs input columns c0 .. c0+255 (the h part) only
  const int in_dim = (s == 3) ? kPeXyz + kHidden : kHidden;   // s == 3 is pts layer 5
  const int c0 = (s == 3) ? kPeXyz : 0;
  const int outs = kHidden;
  __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(wimg_t + (size_t)g * kKBlockBytes);
  for (int e = threadIdx.x; e < kHidden * 64; e += blockDim.x) {
    const int n = e >> 6, c = e & 63;                 // n = input feature (row of B), c = k within the block
    const int o = kb * 64 + c;                        // output feature = reduction index
    const float v = (o < outs) ? src.w[s][(size_t)o * in_dim + c0 + n] : 0.0f;
    img[n * 64 + (((c >> 3) ^ (n & 7)) << 3) + (c & 7)] = __float2bfloat16_rn(v);
  }
}

struct TensorTable {
  const float* t[NWX_NUM_WEIGHT_TENSORS];     // state_dict order (see pack_network_images)
};

__global__ void fill_consts_kernel(TensorTable tab, MlpConsts* __restrict__ c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 9 * kHidden) {
    const int l = i / kHidden, j = i % kHidden;
    c->bias[l][j] = (l < 8) ? tab.t[2 * l + 1][j] : tab.t[19][j];
  }
  if (i < kHidden) c->w_alpha[i] = tab.t[20][i];
  if (i < 3 * kViewHidden) (&c->w_rgb[0][0])[i] = tab.t[22][i];
  if (i == 0) c->b_alpha = tab.t[21][0];
  if (i < 3) c->b_rgb[i] = tab.t[23][i];
}

// Biases and heads of the two networks for the training kernels.  They change every optimiser step,
// so they cannot ride in the launch parameters (host copy) like at inference; a stream-ordered
// device-to-device cudaMemcpyToSymbolAsync refreshes them and the epilogues keep reading them
// through the constant bank (uniform LDC) instead of global loads.
__constant__ MlpConsts c_train_consts[2];

int upload_train_consts(int which, const MlpConsts* dev_src, cudaStream_t st) {
  NWX_CUDA_TRY(cudaMemcpyToSymbolAsync(c_train_consts, dev_src, sizeof(MlpConsts), (size_t)which * sizeof(MlpConsts),
                                       cudaMemcpyDeviceToDevice, st));
  return NWX_OK;
}

// ------------------------------------------------------------------------------------------------
// dX kernel
// ------------------------------------------------------------------------------------------------
struct DxArgs {
  const float* d_raw;        // [P,4] dL/d(raw rgb, raw sigma)
  const float* hv;           // [P,128] views hidden (post-ReLU) saved by the forward
  const uint8_t* acts;       // activation images (forward)
  const uint32_t* masks;     // ReLU' bit masks (forward)
  uint8_t* grads;            // gradient images (output)
  const uint8_t* wimg_t;     // transposed weight images
  const MlpConsts* gconsts;
  uint32_t* diag;
  int64_t P, n_tiles;
  int iters, which, experiment;
};

enum { kBwdLinear = 0, kBwdSigmaMask = 1, kBwdMask = 2 };

// ReLU' of packed word p of a 32-column chunk as an AND mask for the packed bf16 pair: bit p of the forward's mask
// word -> 0x0000FFFF, bit 16 + p -> 0xFFFF0000.  One shift puts the two bits on the sign positions of bytes 0 and 2,
// one PRMT in sign-replicate mode (selector nibble | 8) smears them over the two halves: 2 instructions per PAIR of
// elements instead of a test + select per element (the per-element form was 28 % of the kernel's instructions and
// made the epilogue warps issue-bound).
template <int kP>
__device__ __forceinline__ uint32_t relu_pair_mask(uint32_t m) {
  const uint32_t x = kP <= 7 ? (m << (7 - kP)) : (m >> (kP - 7));
  uint32_t sel;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sel) : "r"(x), "r"(0u), "r"(0xAA88u));
  return sel;
}

template <int kKind, int kJ>
__device__ __forceinline__ void bwd_pack_pairs(const MlpConsts& cst, const uint32_t (&v)[32], int col, uint32_t m, float dsig,
                                               uint32_t (&pk)[16]) {
  if constexpr (kJ < 16) {
    float a = __uint_as_float(v[2 * kJ]), b = __uint_as_float(v[2 * kJ + 1]);
    if (kKind == kBwdSigmaMask) {                   // d h8 += dsigma * w_alpha (sigma head, nerf_model.py:63)
      a = fmaf(dsig, cst.w_alpha[col + 2 * kJ], a);
      b = fmaf(dsig, cst.w_alpha[col + 2 * kJ + 1], b);
    }
    uint32_t w = pack_bf16x2(a, b);
    if (kKind != kBwdLinear) w &= relu_pair_mask<kJ>(m);     // ReLU': the saved activation is > 0
    pk[kJ] = w;
    bwd_pack_pairs<kKind, kJ + 1>(cst, v, col, m, dsig, pk);
  }
}

template <int kKind>
__device__ __forceinline__ void bwd_chunk(const MlpConsts& cst, const uint32_t (&v)[32], int col, uint32_t m, float dsig,
                                          uint32_t hrow, bool to_smem, bool to_gmem, int row, uint8_t* grow) {
  uint32_t pk[16];
  bwd_pack_pairs<kKind, 0>(cst, v, col, m, dsig, pk);
  const uint32_t kbo = (uint32_t)(col >> 6) * kTileImgBytes;
  const int j0 = (col & 63) >> 3, r7 = row & 7;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t off = kbo + (((j0 + q) ^ r7) << 4);
    if (to_smem) st_shared_v4(hrow + off, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    if (to_gmem) *reinterpret_cast<uint4*>(grow + off) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }
}

// One thread = one row (point) over its 128-column half, in four 32-column chunks; the TMEM load of chunk c + 1 is
// in flight while chunk c is processed (like the forward's epilogue_hidden).
template <int kKind>
__device__ __forceinline__ void bwd_epilogue(const MlpConsts& cst, uint32_t d_tmem, uint32_t hrow, bool to_smem, bool to_gmem,
                                             int row, int wg, const uint32_t (&mask)[4] /* ReLU' bit words of this row's four
                                             32-column chunks, loaded one step ahead */, uint8_t* grow, float dsig) {
  const int col0 = wg * 128;
  uint32_t va[32], vb[32];
  tmem_ld32(d_tmem + col0, va);
  tmem_wait_ld_dep(va);
  tmem_ld32(d_tmem + col0 + 32, vb);
  bwd_chunk<kKind>(cst, va, col0, mask[0], dsig, hrow, to_smem, to_gmem, row, grow);
  tmem_wait_ld_dep(vb);
  tmem_ld32(d_tmem + col0 + 64, va);
  bwd_chunk<kKind>(cst, vb, col0 + 32, mask[1], dsig, hrow, to_smem, to_gmem, row, grow);
  tmem_wait_ld_dep(va);
  tmem_ld32(d_tmem + col0 + 96, vb);
  bwd_chunk<kKind>(cst, va, col0 + 64, mask[2], dsig, hrow, to_smem, to_gmem, row, grow);
  tmem_wait_ld_dep(vb);
  bwd_chunk<kKind>(cst, vb, col0 + 96, mask[3], dsig, hrow, to_smem, to_gmem, row, grow);
}

__global__ void __launch_bounds__(kThreads, 1)
mlp_bwd_dx_kernel(const __grid_constant__ DxArgs args) {
  constexpr int kStages = 4;
  using L = SmemLayout<true, kStages>;
  const MlpConsts& cst = c_train_consts[args.which];
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0),   // warp-uniform for the compiler
             lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int units = gridDim.x / 2, unit = blockIdx.x / 2;
  const int64_t P = args.P;
  const int iters = args.iters;
  auto tile_of = [&](int it, int t) -> int64_t { return (((int64_t)it * units + unit) * 2 + t) * 2 + rank; };
  auto leader = [&](uint32_t local) -> uint32_t { return mapa(local, 0); };

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(sbase + L::w_full + 8 * s, 1);
      mbar_init(sbase + L::w_empty + 8 * s, 1);
      mbar_init(sbase + L::w_peer + 8 * s, 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(sbase + L::acc_full + 8 * t, 1);
      mbar_init(sbase + L::a_ready + 8 * t, 16);
      mbar_init(sbase + L::pe_ready + 8 * t, 8);
      mbar_init(sbase + L::pe_free + 8 * t, 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<2>(sbase + L::tmem_slot, 512);
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (sbase - smem_u32(smem_raw)) + L::tmem_slot);

  if (warp == 0) {
    if (lane == 0) {                                           // ---- TMA producer: transposed weights
      const WaitCtx wc{args.diag, 0x1100u};
      uint32_t fill = 0;
      const uint32_t bytes = kKBlockBytes / 2;
      for (int it = 0; it < iters; ++it)
        for (int s = 0; s < kDxSteps; ++s)
          for (int kb = 0; kb < dx_step_nkb(s); ++kb, ++fill) {
            const uint32_t stage = fill % kStages, round = fill / kStages;
            mbar_wait(sbase + L::w_empty + 8 * stage, (round & 1) ^ 1, wc);
            const uint32_t bar = sbase + L::w_full + 8 * stage;
            mbar_arrive_expect_tx(bar, bytes);
            bulk_g2s(sbase + L::w0 + stage * L::kStageBytes,
                     args.wimg_t + (size_t)(dx_step_kb0(s) + kb) * kKBlockBytes + rank * bytes, bytes, bar);
          }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                           // ---- MMA issuer / relay
      const WaitCtx wc{args.diag, 0x1200u};
      uint32_t fill = 0;
      if (rank == 0) {
        const uint32_t idesc = umma_idesc_bf16(256, kHidden);
        for (int it = 0; it < iters; ++it) {
          for (int s = 0; s < kDxSteps; ++s) {
            const int nkb = dx_step_nkb(s);
            for (int t = 0; t < 2; ++t) {
              if (s == 0) mbar_wait(sbase + L::pe_ready + 8 * t, it & 1, wc);
              // a_ready[t] completes once per step epilogue: phase index = kDxSteps*it + s - 1
              if (s != 0 || it != 0) mbar_wait(sbase + L::a_ready + 8 * t, (kDxSteps * it + s + 1) & 1, wc);
              tc_fence_after();
              const uint32_t d_tmem = tmem_base + t * kHidden;
              for (int kb = 0; kb < nkb; ++kb) {
                const uint32_t f = fill + kb;
                const uint32_t stage = f % kStages, round = f / kStages;
                if (t == 0) {
                  mbar_wait(sbase + L::w_full + 8 * stage, round & 1, wc);
                  mbar_wait(sbase + L::w_peer + 8 * stage, round & 1, wc);
                  tc_fence_after();
                }
                const uint64_t adesc = umma_desc_k_sw128(sbase + L::h0 + t * kHBytes + kb * kABlock);
                const uint64_t bdesc = umma_desc_k_sw128(sbase + L::w0 + stage * L::kStageBytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16<2>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                if (t == 1) umma_commit<2>(sbase + L::w_empty + 8 * stage);
              }
              umma_commit<2>(sbase + L::acc_full + 8 * t);
              if (s == kDxSteps - 1) umma_commit<2>(sbase + L::pe_free + 8 * t);   // tile buffer free for the next G_views
            }
            fill += nkb;
          }
        }
      } else {
        for (uint32_t n = 0; n < (uint32_t)kDxKBlocks * (uint32_t)iters; ++n, ++fill) {
          const uint32_t stage = fill % kStages, round = fill / kStages;
          mbar_wait(sbase + L::w_full + 8 * stage, round & 1, wc);
          mbar_arrive_cluster(mapa(sbase + L::w_peer + 8 * stage, 0));
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ---- G_views producers: g_v = (W_rgb^T d_rgb) * [hv > 0]   (rgb head + views ReLU backward)
    const WaitCtx wc{args.diag, 0x1300u};
    const int row = (warp - 4) * 32 + lane;
    for (int it = 0; it < iters; ++it) {
      for (int t = 0; t < 2; ++t) {
        const int64_t tile = tile_of(it, t);
        const int64_t p = tile * kTileM + row;
        const bool live = p < P;
        float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) dr = __ldg(reinterpret_cast<const float4*>(args.d_raw) + p);
        // this row of hv (512 B) was prefetched into L2 one tile ahead; fetch the next tile's row now, so the
        // serialized chunk loop below runs at L2 latency while the tensor pipe waits for this tile
        {
          const int64_t pn = (t == 0 ? tile_of(it, 1) : tile_of(it + 1, 0)) * kTileM + row;
          if (pn < P) {
            const float* nx = args.hv + pn * kViewHidden;
#pragma unroll
            for (int q = 0; q < 4; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + q * 32));
          }
        }
        const uint32_t hrow = sbase + L::h0 + t * kHBytes + row * 128;
        uint8_t* grow = (tile < args.n_tiles)
                            ? args.grads + tile_img_offset(grad_slot_kb0(0), 2, args.n_tiles + 1, tile, 0) + row * 128 : nullptr;
        // The whole row (128 columns = 64 packed words) is computed and saved to HBM BEFORE the wait for the
        // shared-memory tile: the hv loads (L2 latency, 32 per thread) then overlap the previous iteration's last
        // MMAs instead of sitting between them and this tile's first one; after the wait only 16 STS.128 remain.
        uint32_t pk[64];
#pragma unroll
        for (int c8 = 0; c8 < 16; ++c8) {                       // 16 chunks of 8 columns = 128 columns
          float hvv[8];
          if (live) {
            const float4 h0 = __ldg(reinterpret_cast<const float4*>(args.hv + p * kViewHidden + c8 * 8));
            const float4 h1 = __ldg(reinterpret_cast<const float4*>(args.hv + p * kViewHidden + c8 * 8 + 4));
            hvv[0] = h0.x; hvv[1] = h0.y; hvv[2] = h0.z; hvv[3] = h0.w;
            hvv[4] = h1.x; hvv[5] = h1.y; hvv[6] = h1.z; hvv[7] = h1.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) hvv[e] = 0.f;
          }
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            float g[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int j = c8 * 8 + e + h;
              const float v = fmaf(cst.w_rgb[0][j], dr.x, fmaf(cst.w_rgb[1][j], dr.y, cst.w_rgb[2][j] * dr.z));
              g[h] = hvv[e + h] > 0.f ? v : 0.f;
            }
            pk[c8 * 4 + (e >> 1)] = pack_bf16x2(g[0], g[1]);
          }
          if (grow) {
            const uint32_t off = (uint32_t)(c8 >> 3) * kTileImgBytes + (((c8 & 7) ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(grow + off) = make_uint4(pk[c8 * 4], pk[c8 * 4 + 1], pk[c8 * 4 + 2], pk[c8 * 4 + 3]);
          }
        }
        if (it > 0) mbar_wait(sbase + L::pe_free + 8 * t, (it - 1) & 1, wc);
#pragma unroll
        for (int c8 = 0; c8 < 16; ++c8) {
          const uint32_t off = (uint32_t)(c8 >> 3) * kTileImgBytes + (((c8 & 7) ^ (row & 7)) << 4);
          st_shared_v4(hrow + off, pk[c8 * 4], pk[c8 * 4 + 1], pk[c8 * 4 + 2], pk[c8 * 4 + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader(sbase + L::pe_ready + 8 * t));
      }
    }
  } else if (warp >= 8) {
    // ---- epilogues
    const WaitCtx wc{args.diag, 0x1400u};
    const int quad = warp & 3, wg = (warp - 8) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    // ReLU' bit words (forward's masks) of the row's four 32-column chunks, per in-flight tile; always loaded ONE STEP
    // AHEAD (right after the previous step's epilogue of the same tile), so the L2 latency of these four loads is
    // covered by an MMA step instead of opening every epilogue
    uint32_t mreg[2][4];
    auto load_masks = [&](int64_t tile, int layer, uint32_t (&m)[4]) {
      const int64_t mt = tile < args.n_tiles ? tile : 0;
      const uint32_t* mrow = reinterpret_cast<const uint32_t*>(
                                 reinterpret_cast<const uint8_t*>(args.masks) + mask_img_offset(mt, layer)) + row;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) m[cc] = __ldg(mrow + (wg * 4 + cc) * kTileM);
    };
#pragma unroll
    for (int t = 0; t < 2; ++t) load_masks(tile_of(0, t), 7, mreg[t]);
    for (int it = 0; it < iters; ++it) {
      float dsig[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int64_t p = tile_of(it, t) * kTileM + row;
        dsig[t] = (p < P) ? __ldg(args.d_raw + p * 4 + 3) : 0.f;
      }
      for (int s = 0; s < kDxSteps; ++s) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(sbase + L::acc_full + 8 * t, (kDxSteps * it + s) & 1, wc);
          tc_fence_after();
          const int64_t tile = tile_of(it, t);
          // dead tiles (past the end) still run so the barrier protocol stays uniform; they write
          // to a scratch tile at the very end of the gradient image buffer
          const int64_t wt = tile < args.n_tiles ? tile : args.n_tiles;
          const uint32_t d_tmem = lane_addr + t * kHidden;
          const uint32_t hrow = sbase + L::h0 + t * kHBytes + row * 128;
          // step s produces G_{8-s} = dL/d(pre-activation of pts layer 7-s) -> grad slot s + 2 (slot 1 is unused)
          uint8_t* grow = args.grads + tile_img_offset(grad_slot_kb0(s + 2), 4, args.n_tiles + 1, wt, 0) + row * 128;
          // mreg[t] = ReLU' of the activation the produced gradient flows into: step 0 -> h8 (pts layer 7) ... step 7 -> h1
          // All but the last two steps leave their G tile in smem (next step's A operand) and save it with ONE
          // TMA store; the last two store per thread: after the last MMA the G_views producers reuse the
          // buffer, and they cannot wait on another thread's bulk group.
          const bool via_tma = s < kDxSteps - 2;
          if (warp == 8 && lane == 0 && NWX_EXP(args) == 0) bulk_wait_read<1>();   // earlier store of this buffer has finished reading it
          if (NWX_EXP(args) != 14) named_bar_sync(3, 256);
          if (s == 0) bwd_epilogue<kBwdSigmaMask>(cst, d_tmem, hrow, true, false, row, wg, mreg[t], grow, dsig[t]);
          else bwd_epilogue<kBwdMask>(cst, d_tmem, hrow, s != kDxSteps - 1, !via_tma, row, wg, mreg[t], grow, 0.f);
          // next use of this tile slot: step s + 1 of the same tile, or step 0 of the next iteration's tile
          if (s + 1 < kDxSteps) load_masks(tile, 6 - s, mreg[t]);
          else if (it + 1 < iters) load_masks(tile_of(it + 1, t), 7, mreg[t]);
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(leader(sbase + L::a_ready + 8 * t));
          if (via_tma) {
            if (NWX_EXP(args) != 14) named_bar_sync(3, 256);                             // every warp's tile writes are fenced
            if (warp == 8 && lane == 0 && NWX_EXP(args) != 12)
              bulk_s2g(args.grads + tile_img_offset(grad_slot_kb0(s + 2), 4, args.n_tiles + 1, wt, 0),
                       sbase + L::h0 + t * kHBytes, kHBytes);
          }
        }
      }
    }
    if (warp == 8 && lane == 0) bulk_wait_read<0>();            // smem must outlive the last store's reads
  }
  tc_fence_before();
  cluster_sync();
  if (warp == 2) tmem_dealloc<2>(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// dW kernel
// ------------------------------------------------------------------------------------------------
// Job-major: every CTA works on ONE (G, X) pair for its whole life -- a share of the tiles, one dW in TMEM, one
// flush at the end -- instead of walking all 11 jobs with a TMEM flush (and an idle tensor pipe) between them.
constexpr int kDwJobs = 10;
constexpr int kDwMaxParts = 16;    // most CTAs one job is split over
struct DwJob {
  int g_kb0, g_nkb;          // gradient image slot (K-block offset, count): out features = 64 * g_nkb
  int x_kb0, x_nkb;          // activation image slot: in features = 64 * x_nkb (n_valid of them real)
  int w_off, in_stride, col0, n_valid;   // where dW[out][col0 + c] goes in the flat gradient buffer
  int b_off;                 // bias gradient offset, or -1 (second job on the same G)
  int cta0, ncta;            // the CTAs [cta0, cta0 + ncta) share this job: CTA cta0 + k takes tiles k, k + ncta, ...
};
struct DwArgs {
  const uint8_t* acts;
  const uint8_t* grads;
  float* partial;            // [kDwMaxParts][NWX_PARAMS_PER_NET]: row k = the k-th CTA of each job
  uint32_t* diag;
  int64_t n_tiles;
  DwJob job[kDwJobs];
};

constexpr int kDwStages = 3;
constexpr uint32_t kDwStageBytes = 65536;        // 64 points: G 4 x 8 KB | X 4 x 8 KB
struct DwSmem {
  static constexpr uint32_t stage0 = 0;
  static constexpr uint32_t full = kDwStages * kDwStageBytes;      // barriers
  static constexpr uint32_t empty = full + 8 * kDwStages;
  static constexpr uint32_t acc_full = empty + 8 * kDwStages;
  static constexpr uint32_t acc_free = acc_full + 8;
  static constexpr uint32_t tmem_slot = acc_free + 8;
  static constexpr uint32_t total = tmem_slot + 16;
  static constexpr uint32_t alloc_bytes = total + 1024;
};

// MN-major operand descriptor, SWIZZLE_128B: 64 contiguous MN elements (128 B) x 8 K rows per atom;
// LBO = byte distance between 64-element MN blocks, SBO = byte distance between 8-row K groups.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int M, int N) {
  return umma_idesc_bf16(M, N) | (1u << 15) | (1u << 16);     // a_major = b_major = MN
}

__global__ void __launch_bounds__(512, 1)
mlp_bwd_dw_kernel(const __grid_constant__ DwArgs args) {
  using S = DwSmem;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0),   // warp-uniform for the compiler
             lane = threadIdx.x & 31;
  const int64_t n_tiles = args.n_tiles;
  int jidx = 0;
  while (jidx + 1 < kDwJobs && (int)blockIdx.x >= args.job[jidx].cta0 + args.job[jidx].ncta) ++jidx;
  const DwJob jb = args.job[jidx];
  const int part = (int)blockIdx.x - jb.cta0;      // my tiles: part, part + ncta, ...
  const int my_tiles = (int)((n_tiles - part + jb.ncta - 1) / jb.ncta);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kDwStages; ++s) {
      mbar_init(sbase + S::full + 8 * s, 1);
      mbar_init(sbase + S::empty + 8 * s, 2);        // MMA commit + bias reducers
    }
    mbar_init(sbase + S::acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<1>(sbase + S::tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sgen + S::tmem_slot);
  float* partial = args.partial + (size_t)part * NWX_PARAMS_PER_NET;

  if (warp == 0) {
    if (lane == 0) {                                           // ---- producer: 64-point half tiles of G and X
      const WaitCtx wc{args.diag, 0x2100u};
      uint32_t fill = 0;
      {
        const uint32_t bytes = (uint32_t)(jb.g_nkb + jb.x_nkb) * 8192u;
        for (int i = 0; i < my_tiles; ++i) {
          const int64_t tile = part + (int64_t)i * jb.ncta;
          for (int half = 0; half < 2; ++half, ++fill) {
            const uint32_t stage = fill % kDwStages, round = fill / kDwStages;
            mbar_wait(sbase + S::empty + 8 * stage, (round & 1) ^ 1, wc);
            const uint32_t bar = sbase + S::full + 8 * stage;
            mbar_arrive_expect_tx(bar, bytes);
            const uint32_t dst = sbase + S::stage0 + stage * kDwStageBytes;
            for (int kb = 0; kb < jb.g_nkb; ++kb)
              bulk_g2s(dst + kb * 8192, args.grads + tile_img_offset(jb.g_kb0, jb.g_nkb, n_tiles + 1, tile, kb) + half * 8192,
                       8192, bar);
            for (int kb = 0; kb < jb.x_nkb; ++kb)
              bulk_g2s(dst + 32768 + kb * 8192, args.acts + tile_img_offset(jb.x_kb0, jb.x_nkb, n_tiles, tile, kb) + half * 8192,
                       8192, bar);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                           // ---- MMA issuer
      const WaitCtx wc{args.diag, 0x2200u};
      uint32_t fill = 0;
      {
        const int mhalves = jb.g_nkb / 2;
        const uint32_t idesc = umma_idesc_bf16_mn(128, 64 * jb.x_nkb);
        for (int i = 0; i < 2 * my_tiles; ++i, ++fill) {
          const uint32_t stage = fill % kDwStages, round = fill / kDwStages;
          mbar_wait(sbase + S::full + 8 * stage, round & 1, wc);
          tc_fence_after();
          const uint32_t st = sbase + S::stage0 + stage * kDwStageBytes;
          for (int mh = 0; mh < mhalves; ++mh) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {                      // 4 x 16 points; 16 rows = 2048 B
              const uint64_t adesc = umma_desc_mn_sw128(st + mh * 16384 + k * 2048, 8192, 1024);
              const uint64_t bdesc = umma_desc_mn_sw128(st + 32768 + k * 2048, 8192, 1024);
              umma_bf16<1>(tmem_base + mh * 256, adesc, bdesc, idesc, (i | k) != 0);
            }
          }
          umma_commit<1>(sbase + S::empty + 8 * stage);
        }
        umma_commit<1>(sbase + S::acc_full);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ---- flush: TMEM dW (lane = output feature) -> this CTA's partial
    const WaitCtx wc{args.diag, 0x2300u};
    const int quad = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    {
      mbar_wait(sbase + S::acc_full, 0, wc);
      tc_fence_after();
      for (int mh = 0; mh < jb.g_nkb / 2; ++mh) {
        const int o = mh * 128 + quad * 32 + lane;
        float* dst = partial + jb.w_off + (size_t)o * jb.in_stride + jb.col0;
        for (int c0 = 0; c0 < 64 * jb.x_nkb; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(lane_addr + mh * 256 + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c0 + c < jb.n_valid) dst[c0 + c] = __uint_as_float(v[c]);
        }
      }
    }
  } else if (warp >= 8) {
    // ---- bias gradients: column sums of the G half-tiles, straight from shared memory
    const WaitCtx wc{args.diag, 0x2400u};
    const int c = threadIdx.x - 256;                 // column 0..255
    uint32_t fill = 0;
    {
      const bool mine = jb.b_off >= 0 && c < 64 * jb.g_nkb;
      float acc = 0.f;
      for (int i = 0; i < 2 * my_tiles; ++i, ++fill) {
        const uint32_t stage = fill % kDwStages, round = fill / kDwStages;
        mbar_wait(sbase + S::full + 8 * stage, round & 1, wc);
        if (mine) {
          const uint8_t* blk = sgen + S::stage0 + stage * kDwStageBytes + (c >> 6) * 8192 + (c & 7) * 2;
          const int ch = (c & 63) >> 3;
#pragma unroll 8
          for (int r = 0; r < 64; ++r)
            acc += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(blk + r * 128 + ((ch ^ (r & 7)) << 4)));
        }
        named_bar_sync(2, 256);                      // all 8 reducer warps done with this stage
        if (threadIdx.x == 256) mbar_arrive(sbase + S::empty + 8 * stage);
      }
      if (mine) partial[jb.b_off + c] = acc;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}

// grad[i] += sum over the CTAs of the job that owns parameter i of partial[k][i], k < that job's ncta
struct DwJobTable { DwJob job[kDwJobs]; };
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partial, const __grid_constant__ DwJobTable tab, float* __restrict__ grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NWX_PARAMS_PER_NET) return;
  int n_parts = 0;                                   // 0: not produced by the dW kernel (heads, view-direction columns)
#pragma unroll 1
  for (int j = 0; j < kDwJobs; ++j) {
    const DwJob& jb = tab.job[j];
    const int rows = 64 * jb.g_nkb;
    // the last job's weight region holds d W_fold, which is not a parameter: fold_grads_*_kernel consume it
    if (j != kDwJobs - 1 && i >= jb.w_off && i < jb.w_off + rows * jb.in_stride) {
      const int col = (i - jb.w_off) % jb.in_stride;
      if (col >= jb.col0 && col < jb.col0 + jb.n_valid) { n_parts = jb.ncta; break; }
    }
    if (jb.b_off >= 0 && i >= jb.b_off && i < jb.b_off + rows) { n_parts = jb.ncta; break; }
  }
  if (n_parts == 0) return;
  float s = 0.f;
  for (int c = 0; c < n_parts; ++c) s += partial[(size_t)c * NWX_PARAMS_PER_NET + i];
  grad[i] += s;
}

// ------------------------------------------------------------------------------------------------
// chain rule through the fold W_fold = V . F, b_fold = b_view + V . b_F   (V = W_view[:, :256] [128 x 256],
// F = W_feature [256 x 256]):   dV = dW_fold . F^T + d b_fold (x) b_F,   dF = V^T . dW_fold,   d b_F = V^T . d b_fold,
// d b_view = d b_fold
// ------------------------------------------------------------------------------------------------
// fold[o][k] = sum over the job's CTAs of partial[part][w_off + o * 283 + k]; fold[128*256 + o] likewise for d b_fold
__global__ void __launch_bounds__(256)
fold_sum_kernel(const float* __restrict__ partial, int n_parts, int w_off, int b_off, float* __restrict__ fold) {
  const int o = blockIdx.x, k = threadIdx.x;
  float s = 0.f;
  for (int c = 0; c < n_parts; ++c) s += partial[(size_t)c * NWX_PARAMS_PER_NET + w_off + o * (kHidden + kPeDir) + k];
  fold[o * kHidden + k] = s;
  if (k == 0) {
    float b = 0.f;
    for (int c = 0; c < n_parts; ++c) b += partial[(size_t)c * NWX_PARAMS_PER_NET + b_off + o];
    fold[kViewHidden * kHidden + o] = b;
  }
}

// dF[m][k] += sum_o V[o][m] fold[o][k]  (4 rows m per block);  d b_F[m] += sum_o V[o][m] dbfold[o]
__global__ void __launch_bounds__(256)
fold_grads_feature_kernel(const float* __restrict__ wv, const float* __restrict__ fold, float* __restrict__ d_wf,
                          float* __restrict__ d_bf) {
  __shared__ float v[kViewHidden][4];
  const int m0 = blockIdx.x * 4, k = threadIdx.x;
  for (int e = threadIdx.x; e < kViewHidden * 4; e += blockDim.x) v[e >> 2][e & 3] = wv[(size_t)(e >> 2) * (kHidden + kPeDir) + m0 + (e & 3)];
  __syncthreads();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int o = 0; o < kViewHidden; ++o) {
    const float f = fold[o * kHidden + k];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = fmaf(v[o][q], f, acc[q]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) d_wf[(size_t)(m0 + q) * kHidden + k] += acc[q];
  if (k < 4) {
    float b = 0.f;
    for (int o = 0; o < kViewHidden; ++o) b = fmaf(v[o][k], fold[kViewHidden * kHidden + o], b);
    d_bf[m0 + k] += b;
  }
}

// dV[o][m] += sum_k fold[o][k] F[m][k] + dbfold[o] b_F[m]  (b_fold = b_view + V b_F depends on V too):
// 32 x 32 output tile per block, both operands staged through smem
__global__ void __launch_bounds__(256)
fold_grads_view_kernel(const float* __restrict__ wf, const float* __restrict__ bf, const float* __restrict__ fold,
                       float* __restrict__ d_wv) {
  __shared__ float sa[32][33], sb[32][33];
  const int o0 = blockIdx.y * 32, m0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 8 rows of threads: 4 outputs each
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < kHidden; k0 += 32) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      sa[ty + 8 * q][tx] = fold[(o0 + ty + 8 * q) * kHidden + k0 + tx];
      sb[ty + 8 * q][tx] = wf[(size_t)(m0 + ty + 8 * q) * kHidden + k0 + tx];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const float b = sb[tx][kk];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(sa[ty + 8 * q][kk], b, acc[q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
    d_wv[(size_t)(o0 + ty + 8 * q) * (kHidden + kPeDir) + m0 + tx] +=
        fmaf(fold[kViewHidden * kHidden + o0 + ty + 8 * q], bf[m0 + tx], acc[q]);
}

// ------------------------------------------------------------------------------------------------
// heads: rgb (3x128), sigma (1x256) and the 27 view-direction columns of the views layer -- fp32.
// Three streaming passes with wide loads (the data is read once; every thread keeps 8 x 16 B in flight),
// per-block partial sums instead of atomics (deterministic), then one small reduction.
// ------------------------------------------------------------------------------------------------
struct HeadArgs {
  const float* d_raw;        // [P,4]
  const float* hv;           // [P,128]
  const uint8_t* acts;       // h8 images (act slot 8)
  const float* pe_dir;       // [n_rays, 27]
  const MlpConsts* gconsts;
  float* head_partial;       // [rows][kHeadOut] per-block sums
  float* gsum;               // [n_rays * segs][128]: per ray segment, sum over its points of g_v = (W_rgb^T d_rgb) * [hv > 0]
  int64_t P, n_tiles, n_rays;
  int S, segs;               // segs = ceil(S / kSegPoints) segments per ray
};
// per-block output layout: w_rgb [3][128] | view-direction columns [128][27] | w_alpha [256] | b_rgb [3], b_alpha
constexpr int kHeadOffDir = 3 * kViewHidden, kHeadOffAlpha = kHeadOffDir + kViewHidden * kPeDir,
              kHeadOffBias = kHeadOffAlpha + kHidden, kHeadValid = kHeadOffBias + 4, kHeadOut = 4128;
constexpr int kSegPoints = 32;           // head_rgb work unit: one warp, 32 consecutive points of one ray
constexpr int kDirUnitsPerBlock = 64;
constexpr int kSigmaTilesPerBlock = 4;

// rgb head: warp = ray segment, lane = 4 views-hidden columns.  d W_rgb[c][j] = sum_p d_rgb[p][c] hv[p][j],
// d b_rgb / d b_alpha = sum_p d_raw[p], and the per-segment sum of g_v for the view-direction columns.
// Persistent: the grid is sized to the machine and the warps stride over the segments.
__global__ void __launch_bounds__(256, 3)
head_rgb_kernel(const HeadArgs a) {
  __shared__ float4 red[8][3][32];
  __shared__ float4 redb[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const MlpConsts& cst = *a.gconsts;
  float wr[3][4], acc[3][4];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int e = 0; e < 4; ++e) { wr[c][e] = cst.w_rgb[c][4 * lane + e]; acc[c][e] = 0.f; }
  float4 db = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t n_units = a.n_rays * a.segs;
  for (int64_t u = (int64_t)blockIdx.x * 8 + warp; u < n_units; u += (int64_t)gridDim.x * 8) {
    const int64_t ray = u / a.segs;
    const int r_begin = (int)(u - ray * a.segs) * kSegPoints;
    const int n = (a.S - r_begin) < kSegPoints ? (a.S - r_begin) : kSegPoints;
    const int64_t p0 = ray * a.S + r_begin;
    float gs[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r0 = 0; r0 < n; r0 += 8) {
      float4 h[8], d[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {                             // all loads first
        const int64_t p = p0 + (r0 + q < n ? r0 + q : n - 1);
        h[q] = ldg_stream4(reinterpret_cast<const float4*>(a.hv + p * kViewHidden) + lane);
        d[q] = __ldg(reinterpret_cast<const float4*>(a.d_raw) + p);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (r0 + q >= n) break;
        const float hh[4] = {h[q].x, h[q].y, h[q].z, h[q].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[0][e] = fmaf(d[q].x, hh[e], acc[0][e]);
          acc[1][e] = fmaf(d[q].y, hh[e], acc[1][e]);
          acc[2][e] = fmaf(d[q].z, hh[e], acc[2][e]);
          if (hh[e] > 0.f) gs[e] += fmaf(wr[0][e], d[q].x, fmaf(wr[1][e], d[q].y, wr[2][e] * d[q].z));
        }
        db.x += d[q].x; db.y += d[q].y; db.z += d[q].z; db.w += d[q].w;
      }
    }
    reinterpret_cast<float4*>(a.gsum + u * kViewHidden)[lane] = make_float4(gs[0], gs[1], gs[2], gs[3]);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) red[warp][c][lane] = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
  if (lane == 0) redb[warp] = db;
  __syncthreads();
  float* hp = a.head_partial + (size_t)blockIdx.x * kHeadOut;
  if (threadIdx.x < 96) {                                       // (c, lane) -> 4 columns, summed over the 8 warps in order
    const int c = threadIdx.x >> 5, l = threadIdx.x & 31;
    float4 t = red[0][c][l];
#pragma unroll
    for (int w = 1; w < 8; ++w) { const float4 v = red[w][c][l]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
    reinterpret_cast<float4*>(hp + c * kViewHidden)[l] = t;
  } else if (threadIdx.x == 96) {
    float4 t = redb[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) { t.x += redb[w].x; t.y += redb[w].y; t.z += redb[w].z; t.w += redb[w].w; }
    hp[kHeadOffBias + 0] = t.x; hp[kHeadOffBias + 1] = t.y; hp[kHeadOffBias + 2] = t.z; hp[kHeadOffBias + 3] = t.w;
  }
}

// view-direction columns of the views layer: d W_view[j][256 + i] = sum_segments gsum[seg][j] * pe_dir[ray(seg)][i]
__global__ void __launch_bounds__(kViewHidden)
head_dir_kernel(const HeadArgs a) {
  __shared__ float pe[kDirUnitsPerBlock][kPeDir];
  const int64_t n_units = a.n_rays * a.segs;
  const int64_t u0 = (int64_t)blockIdx.x * kDirUnitsPerBlock;
  const int n = (int)((n_units - u0) < kDirUnitsPerBlock ? (n_units - u0) : kDirUnitsPerBlock);
  for (int e = threadIdx.x; e < n * kPeDir; e += blockDim.x) {
    const int r = e / kPeDir, i = e - r * kPeDir;
    pe[r][i] = __ldg(a.pe_dir + ((u0 + r) / a.segs) * kPeDir + i);
  }
  __syncthreads();
  const int j = threadIdx.x;
  float acc[kPeDir];
#pragma unroll
  for (int i = 0; i < kPeDir; ++i) acc[i] = 0.f;
  for (int r = 0; r < n; r += 4) {
    float g[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) g[q] = (r + q < n) ? __ldg(a.gsum + (u0 + r + q) * kViewHidden + j) : 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int i = 0; i < kPeDir; ++i) acc[i] = fmaf(g[q], pe[(r + q < n) ? r + q : 0][i], acc[i]);
  }
  float* hp = a.head_partial + (size_t)blockIdx.x * kHeadOut + kHeadOffDir + j * kPeDir;
#pragma unroll
  for (int i = 0; i < kPeDir; ++i) hp[i] = acc[i];
}

// sigma head: d w_alpha[j] = sum_p dsigma[p] h8[p][j] over the saved (swizzled, bf16) h8 tile images.
// warp = one K-block (64 columns) of a tile at a time; lane = (row & 3, 16-byte chunk): 512 B per load instruction.
__global__ void __launch_bounds__(256, 4)
head_sigma_kernel(const HeadArgs a) {
  __shared__ float red[8][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb = warp & 3, lc = lane & 7, rs = lane >> 3;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  const int64_t tile0 = (int64_t)blockIdx.x * kSigmaTilesPerBlock;
  for (int64_t tile = tile0 + (warp >> 2); tile < tile0 + kSigmaTilesPerBlock && tile < a.n_tiles; tile += 2) {
    const uint8_t* img = a.acts + tile_img_offset(act_slot_kb0(8), 4, a.n_tiles, tile, kb);
    const int64_t p0 = tile * kTileM;
#pragma unroll 1
    for (int i0 = 0; i0 < 32; i0 += 8) {
      uint4 v[8];
      float ds[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int r = rs + 4 * (i0 + q);
        v[q] = *reinterpret_cast<const uint4*>(img + r * 128 + ((lc ^ (r & 7)) << 4));
        ds[q] = (p0 + r < a.P) ? __ldg(a.d_raw + (p0 + r) * 4 + 3) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[2 * e] = fmaf(ds[q], __uint_as_float(w[e] << 16), acc[2 * e]);
          acc[2 * e + 1] = fmaf(ds[q], __uint_as_float(w[e] & 0xFFFF0000u), acc[2 * e + 1]);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {                                 // the four row phases of a warp hold the same columns
    acc[e] += __shfl_xor_sync(kFull, acc[e], 8);
    acc[e] += __shfl_xor_sync(kFull, acc[e], 16);
  }
  if (rs == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) red[warp][lc * 8 + e] = acc[e];
  }
  __syncthreads();
  const int c = threadIdx.x;                                    // column 0..255 = K-block c / 64: warps kb and kb + 4
  a.head_partial[(size_t)blockIdx.x * kHeadOut + kHeadOffAlpha + c] = red[c >> 6][c & 63] + red[4 + (c >> 6)][c & 63];
}

// grad[dst(i)] += sum over the rows of head_partial that hold entry i (the rgb / dir / sigma passes use different
// numbers of blocks).  One block per 32 entries; its 8 warps split the rows, then combine in a fixed order.
__global__ void __launch_bounds__(256)
reduce_heads_kernel(const float* __restrict__ hp, int n_rgb, int n_dir, int n_sigma, float* __restrict__ grad, int off_wv,
                    int off_walpha, int off_balpha, int off_wrgb, int off_brgb) {
  __shared__ float red[8][32];
  const int i = blockIdx.x * 32 + (threadIdx.x & 31), part = threadIdx.x >> 5;
  float s = 0.f;
  if (i < kHeadValid) {
    const int n_blocks = (i < kHeadOffDir || i >= kHeadOffBias) ? n_rgb : (i < kHeadOffAlpha ? n_dir : n_sigma);
    for (int b = part; b < n_blocks; b += 8) s += hp[(size_t)b * kHeadOut + i];
  }
  red[part][threadIdx.x & 31] = s;
  __syncthreads();
  if (part != 0 || i >= kHeadValid) return;
#pragma unroll
  for (int w = 1; w < 8; ++w) s += red[w][threadIdx.x];
  int dst;
  if (i < kHeadOffDir) dst = off_wrgb + i;
  else if (i < kHeadOffAlpha) { const int k = i - kHeadOffDir; dst = off_wv + (k / kPeDir) * (kHidden + kPeDir) + kHidden + k % kPeDir; }
  else if (i < kHeadOffBias) dst = off_walpha + (i - kHeadOffAlpha);
  else dst = (i - kHeadOffBias < 3) ? off_brgb + (i - kHeadOffBias) : off_balpha;
  grad[dst] += s;
}

// d(loss)/d(rgb) for loss = mean((rgb_c - gt)^2) + mean((rgb_f - gt)^2)  (training handler:291-305);
// loss_out[0..1] = the two MSE terms (double, like the reference's fp64 loss).  Summed in a fixed order
// (thread -> warp -> block -> last block over the per-block partials), so two runs agree bit for bit.
__global__ void __launch_bounds__(256)
mse_grad_kernel(const float* __restrict__ rgb_c, const float* __restrict__ rgb_f, const float* __restrict__ gt,
                int64_t n3, float* __restrict__ d_c, float* __restrict__ d_f, double* __restrict__ part,
                unsigned int* __restrict__ ticket, double* __restrict__ loss_out) {
  __shared__ double sw[8][2];
  __shared__ bool last;
  double lc = 0.0, lf = 0.0;
  const float scale = 2.0f / (float)n3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = gt[i], ec = rgb_c[i] - g, ef = rgb_f[i] - g;
    d_c[i] = scale * ec; d_f[i] = scale * ef;
    lc += (double)ec * ec; lf += (double)ef * ef;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lc += __shfl_xor_sync(kFull, lc, o); lf += __shfl_xor_sync(kFull, lf, o); }
  if ((threadIdx.x & 31) == 0) { sw[threadIdx.x >> 5][0] = lc; sw[threadIdx.x >> 5][1] = lf; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double bc = 0.0, bf = 0.0;
    for (int w = 0; w < 8; ++w) { bc += sw[w][0]; bf += sw[w][1]; }
    part[2 * blockIdx.x] = bc; part[2 * blockIdx.x + 1] = bf;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double tc = 0.0, tf = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) { tc += __ldcg(part + 2 * b); tf += __ldcg(part + 2 * b + 1); }
    loss_out[0] = tc / (double)n3; loss_out[1] = tf / (double)n3;
    *ticket = 0u;                                               // ready for the next launch on this stream
  }
}

// torch.optim.Adam, default betas/eps, no weight decay (training handler:234); g is pre-scaled by grad_scale.
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            float lr, float b1, float b2, float eps, float bc1, float bc2, float grad_scale) {
  const float step = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= step * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

// ------------------------------------------------------------------------------------------------
// Adam fused with the re-pack (both networks, ONE launch).  The optimiser step used to be 1 + 2 x 6 launches:
// adam_kernel, then per network pack_weights / pack_dir / a D2D copy / pack_weights_t / fill_consts (+ pack_fold),
// all tiny and latency-bound (0.25 ms of a 5 ms step).  Every derived image is a pure function of ONE parameter per
// element, so the thread that updates a parameter also writes its bf16 copy into the forward K-block image and into
// the transposed image of the dX kernel, or its fp32 copy into the device-side constants / view-direction table.
// Only the folded views layer (a 256-long dot product per element) keeps its own kernel, one launch for both networks.
// Same bytes as train_pack() of the updated parameters (tests/test_gpu_train.py compares the buffers).
// ------------------------------------------------------------------------------------------------
struct AdamPackNet {
  float* params; const float* grads; float* m; float* v;     // flat [NWX_PARAMS_PER_NET], state_dict order
  uint8_t* wimg; uint8_t* wimg_t; MlpConsts* gconsts; float* wdir_t; float* bview;
};
struct AdamPackArgs {
  AdamPackNet net[2];
  int off[NWX_NUM_WEIGHT_TENSORS + 1];
  float step, inv_sqrt_bc2, bc2, b1, b2, eps, grad_scale;
};

__device__ __forceinline__ void put_bf16_swizzled(uint8_t* img, int n, int c, float v) {
  reinterpret_cast<__nv_bfloat16*>(img)[n * 64 + ((((c >> 3) ^ (n & 7))) << 3) + (c & 7)] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(256)
adam_pack_kernel(const __grid_constant__ AdamPackArgs a) {
  const int64_t total = 2 * (int64_t)NWX_PARAMS_PER_NET;
  const float inv_sqrt_bc2 = rsqrtf(a.bc2);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int which = idx >= NWX_PARAMS_PER_NET ? 1 : 0;
    const int i = (int)(idx - (int64_t)which * NWX_PARAMS_PER_NET);
    const AdamPackNet& nt = a.net[which];
    // torch.optim.Adam (same arithmetic as adam_kernel)
    const float gi = nt.grads[i] * a.grad_scale;
    const float mi = a.b1 * nt.m[i] + (1.f - a.b1) * gi;
    const float vi = a.b2 * nt.v[i] + (1.f - a.b2) * gi * gi;
    nt.m[i] = mi; nt.v[i] = vi;
    const float p = nt.params[i] - a.step * mi / (sqrtf(vi) * inv_sqrt_bc2 + a.eps);
    nt.params[i] = p;
    // which tensor (24 sorted offsets: 5-step binary search)
    int t = 0;
#pragma unroll
    for (int stp = 16; stp > 0; stp >>= 1)
      if (t + stp < NWX_NUM_WEIGHT_TENSORS && i >= a.off[t + stp]) t += stp;
    const int r = i - a.off[t];
    if (t < 16) {
      const int l = t >> 1;
      if (t & 1) { nt.gconsts->bias[l][r] = p; continue; }                  // _pts_linears.l.bias
      const int in_dim = l == 0 ? kPeXyz : (l == 5 ? kPeXyz + kHidden : kHidden);
      const int o = r / in_dim, k = r - o * in_dim;                          // W[o][k]
      // forward image: K-block kb, column c (layer 5 = [pe(63) | h(256)], nerf_model.py:59)
      int kb, c;
      if (l == 5 && k >= kPeXyz) { kb = 1 + ((k - kPeXyz) >> 6); c = (k - kPeXyz) & 63; }
      else if (l == 0 || l == 5) { kb = 0; c = k; }
      else { kb = k >> 6; c = k & 63; }
      const int g = (l == 0 ? 0 : (l <= 5 ? 1 + 4 * (l - 1) : 22 + 4 * (l - 6))) + kb;
      put_bf16_swizzled(nt.wimg + kblock_offset(g), o, c, p);
      // transposed image of the dX kernel: step s = 8 - l, B[n = input feature][k = output feature]
      if (l >= 1) {
        const int c0 = l == 5 ? kPeXyz : 0;
        if (k >= c0) {
          const int s = 8 - l;
          put_bf16_swizzled(nt.wimg_t + (size_t)(dx_step_kb0(s) + (o >> 6)) * kKBlockBytes, k - c0, o & 63, p);
        }
      }
    } else if (t == 16) {                                                    // _views_linears.0.weight [128][283]
      const int j = r / (kHidden + kPeDir), mcol = r - j * (kHidden + kPeDir);
      if (mcol >= kHidden) nt.wdir_t[(mcol - kHidden) * kViewHidden + j] = p;  // [:, :256] goes through the fold
    } else if (t == 17) nt.bview[r] = p;
    else if (t == 19) nt.gconsts->bias[8][r] = p;                            // _feature_linear.bias (18: weight, fold only)
    else if (t == 20) nt.gconsts->w_alpha[r] = p;
    else if (t == 21) nt.gconsts->b_alpha = p;
    else if (t == 22) (&nt.gconsts->w_rgb[0][0])[r] = p;
    else if (t == 23) nt.gconsts->b_rgb[r] = p;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// offsets of the 24 tensors in the flat state_dict-ordered parameter / gradient buffer
struct FlatLayout {
  int off[NWX_NUM_WEIGHT_TENSORS];
  FlatLayout() {
    const int sizes[NWX_NUM_WEIGHT_TENSORS] = {
        256 * 63, 256, 256 * 256, 256, 256 * 256, 256, 256 * 256, 256, 256 * 256, 256, 256 * 319, 256,
        256 * 256, 256, 256 * 256, 256, 128 * 283, 128, 256 * 256, 256, 256, 1, 3 * 128, 3};
    int o = 0;
    for (int i = 0; i < NWX_NUM_WEIGHT_TENSORS; ++i) { off[i] = o; o += sizes[i]; }
  }
};
static const FlatLayout g_flat;

const int* flat_offsets() { return g_flat.off; }
size_t packed_transposed_bytes() { return (size_t)kDxKBlocks * kKBlockBytes; }

// Re-pack one network from its flat fp32 master parameters (state_dict order): forward images,
// transposed images for dX, device-side biases/heads.  Stream-ordered, no host synchronisation,
// so it can follow the optimiser step directly.
int train_pack(PackedNet& net, const float* params_flat, cudaStream_t st) {
  if (!net.wimg_t) NWX_CUDA_TRY(cudaMalloc(&net.wimg_t, (size_t)kDxKBlocks * kKBlockBytes));
  if (!net.gconsts) NWX_CUDA_TRY(cudaMalloc(&net.gconsts, sizeof(MlpConsts)));
  TensorTable tab;
  for (int i = 0; i < NWX_NUM_WEIGHT_TENSORS; ++i) tab.t[i] = params_flat + g_flat.off[i];
  int rc = pack_network_images(net, tab.t, true, st, net.wimg_t);   // + folded views layer, forward and transposed
  if (rc) return rc;
  net.master = params_flat;
  PackTSrc ts;
  ts.w[0] = nullptr;
  for (int s = 1; s < kDxSteps; ++s) ts.w[s] = tab.t[2 * (8 - s)];     // pts layers 7..1
  pack_weights_t_kernel<<<kDxKBlocks - 2, 256, 0, st>>>(ts, net.wimg_t);
  NWX_LAUNCHED();
  fill_consts_kernel<<<(9 * kHidden + 255) / 256, 256, 0, st>>>(tab, net.gconsts);
  NWX_LAUNCHED();
  net.loaded = true;       // the training kernels are usable; the host copy of the consts is NOT refreshed
  net.consts_stale = true;
  return NWX_OK;
}

// ---- launchers ------------------------------------------------------------------------------------
size_t act_image_bytes(int64_t n_tiles) { return (size_t)kActKBlocksPerTile * n_tiles * kTileImgBytes; }
int dw_partial_rows() { return kDwMaxParts; }
// rows of head_partial: the most blocks any of the three head passes launches, plus the per-ray gsum [n_rays][128]
static int head_rgb_blocks(int64_t n_units) {
  const int64_t want = (n_units + 7) / 8, cap = 3 * (int64_t)num_sms();      // 3 resident blocks per SM
  return (int)(want < cap ? want : cap);
}
static int64_t head_rows(int64_t n_tiles, int64_t n_units) {
  int64_t r = head_rgb_blocks(n_units);
  const int64_t b = (n_tiles + kSigmaTilesPerBlock - 1) / kSigmaTilesPerBlock, c = (n_units + kDirUnitsPerBlock - 1) / kDirUnitsPerBlock;
  if (b > r) r = b;
  if (c > r) r = c;
  return r;
}
// worst case over the sample counts nwx_render_opts allows (S >= 11): segments per ray = ceil(S / 32) <= S / 11 + 1
size_t head_partial_bytes(int64_t n_tiles, int64_t n_rays, int S) {
  const int64_t n_units = n_rays * ((S + kSegPoints - 1) / kSegPoints);
  return (size_t)(head_rows(n_tiles, n_units) * kHeadOut + n_units * kViewHidden) * sizeof(float);
}
size_t grad_image_bytes(int64_t n_tiles) { return (size_t)kGradKBlocksPerTile * (n_tiles + 1) * kTileImgBytes; }

// Backward of one network: d_raw [P,4] -> flat gradient buffer `grad` (state_dict order, += into it).
int launch_mlp_backward(const PackedNet& net, const TrainBwdArgs& a, cudaStream_t st, cudaStream_t heads_st,
                        cudaEvent_t fork, cudaEvent_t join) {
  if (a.P <= 0) return NWX_OK;
  const int64_t tiles = (a.P + kTileM - 1) / kTileM;
  // ---- dX ----
  {
    using Lay = SmemLayout<true, 4>;
    static PerDeviceOnce configured;               // the attribute is per device
    const int dev = current_device();
    if (configured.need(dev)) {
      NWX_CUDA_TRY(cudaFuncSetAttribute(mlp_bwd_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay::alloc_bytes));
      configured.done(dev);
    }
    DxArgs d{};
    d.d_raw = a.d_raw; d.hv = a.hv; d.acts = a.acts; d.masks = a.masks; d.grads = a.gimg; d.wimg_t = net.wimg_t; d.gconsts = net.gconsts;
    d.diag = a.diag; d.P = a.P; d.n_tiles = tiles; d.which = a.which; d.experiment = a.experiment;
    const int units = (num_sms() & ~1) / 2;
    int64_t need = (tiles + 3) / 4;
    const int use = (int)(need < units ? need : units);
    d.iters = (int)((tiles + (int64_t)use * 4 - 1) / ((int64_t)use * 4));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(use * 2); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = Lay::alloc_bytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    NWX_CUDA_TRY(cudaLaunchKernelEx(&cfg, mlp_bwd_dx_kernel, d));
    g_nwx_launches.fetch_add(1, std::memory_order_relaxed);
  }
  // ---- heads (fp32 CUDA-core passes over d_raw / hv / h8): small, latency-bound kernels that write their own,
  // disjoint part of the gradient buffer.  On their own stream, forked BEHIND the dX kernel (whose CTAs fill the
  // shared memory of every SM), they run underneath the dW kernel -- its CTAs leave 30 KB of shared memory and 1 500
  // threads per SM free -- instead of after it.
  {
    const cudaStream_t hs = heads_st ? heads_st : st;
    if (heads_st) {
      NWX_CUDA_TRY(cudaEventRecord(fork, st));                 // behind dX: inputs ready, gradients zeroed, SMs about to run dW
      NWX_CUDA_TRY(cudaStreamWaitEvent(heads_st, fork, 0));
    }
    HeadArgs h{};
    const int64_t n_rays = a.P / a.S;
    const int segs = (a.S + kSegPoints - 1) / kSegPoints;
    const int64_t n_units = n_rays * segs;
    h.d_raw = a.d_raw; h.hv = a.hv; h.acts = a.acts; h.pe_dir = a.pe_dir; h.gconsts = net.gconsts;
    h.head_partial = a.head_partial; h.gsum = a.head_partial + head_rows(tiles, n_units) * kHeadOut;
    h.P = a.P; h.n_tiles = tiles; h.n_rays = n_rays; h.S = a.S; h.segs = segs;
    const int n_rgb = head_rgb_blocks(n_units);
    const int n_dir = (int)((n_units + kDirUnitsPerBlock - 1) / kDirUnitsPerBlock);
    const int n_sigma = (int)((tiles + kSigmaTilesPerBlock - 1) / kSigmaTilesPerBlock);
    head_rgb_kernel<<<n_rgb, 256, 0, hs>>>(h);
    NWX_LAUNCHED();
    head_dir_kernel<<<n_dir, kViewHidden, 0, hs>>>(h);
    NWX_LAUNCHED();
    head_sigma_kernel<<<n_sigma, 256, 0, hs>>>(h);
    NWX_LAUNCHED();
    reduce_heads_kernel<<<(kHeadValid + 31) / 32, 256, 0, hs>>>(a.head_partial, n_rgb, n_dir, n_sigma, a.grad, g_flat.off[16],
                                                               g_flat.off[20], g_flat.off[21], g_flat.off[22], g_flat.off[23]);
    NWX_LAUNCHED();
    if (heads_st) NWX_CUDA_TRY(cudaEventRecord(join, heads_st));
  }
  // ---- dW ----
  int grid = 0;
  {
    static PerDeviceOnce configured;
    const int dev = current_device();
    if (configured.need(dev)) {
      NWX_CUDA_TRY(cudaFuncSetAttribute(mlp_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DwSmem::alloc_bytes));
      configured.done(dev);
    }
    DwArgs w{};
    w.acts = a.acts; w.grads = a.gimg; w.partial = a.partial; w.diag = a.diag; w.n_tiles = tiles;
    const int* off = g_flat.off;
    auto job = [&](int j, int gslot, int aslot, int wt, int stride, int col0, int nvalid, int bt) {
      w.job[j] = DwJob{grad_slot_kb0(gslot), grad_slot_nkb(gslot), act_slot_kb0(aslot), act_slot_nkb(aslot),
                       off[wt], stride, col0, nvalid, bt >= 0 ? off[bt] : -1, 0, 0};
    };
    job(0, 9, 0, 0, 63, 0, 63, 1);                 // pts0: G1 x PE
    for (int l = 1; l <= 4; ++l) job(l, 9 - l, l, 2 * l, 256, 0, 256, 2 * l + 1);   // pts1..4: G_{l+1} x h_l
    job(5, 4, 0, 10, 319, 0, 63, 11);              // pts5, PE columns
    job(6, 4, 5, 10, 319, 63, 256, -1);            // pts5, h5 columns
    job(7, 3, 6, 12, 256, 0, 256, 13);             // pts6: G7 x h6
    job(8, 2, 7, 14, 256, 0, 256, 15);             // pts7: G8 x h7
    // folded views layer: d W_fold = G_v^T h8 lands where d W_view[:, :256] lives (same shape and stride);
    // fold_grads_*_kernel then turns it into d W_view[:, :256], d W_feature, d b_feature
    job(9, 0, 8, 16, 283, 0, 256, 17);
    // CTAs per job in proportion to the bytes a job streams per tile (largest remainder), at most
    // kDwMaxParts and never more than there are tiles
    int units = 0, given = 0, parts[kDwJobs];
    for (int j = 0; j < kDwJobs; ++j) units += w.job[j].g_nkb + w.job[j].x_nkb;
    const int sms = num_sms();
    for (int j = 0; j < kDwJobs; ++j) {
      parts[j] = sms * (w.job[j].g_nkb + w.job[j].x_nkb) / units;
      if (parts[j] < 1) parts[j] = 1;
      given += parts[j];
    }
    for (int left = sms - given; left > 0;) {        // hand the remainder to the jobs with the most work per CTA
      int best = -1;
      double worst = 0.0;
      for (int j = 0; j < kDwJobs; ++j) {
        const double load = (double)(w.job[j].g_nkb + w.job[j].x_nkb) / parts[j];
        if (parts[j] < kDwMaxParts && load > worst) { worst = load; best = j; }
      }
      if (best < 0) break;
      ++parts[best]; --left;
    }
    for (int j = 0; j < kDwJobs; ++j) {
      int n = parts[j] > kDwMaxParts ? kDwMaxParts : parts[j];
      if (n > tiles) n = (int)tiles;
      w.job[j].cta0 = grid; w.job[j].ncta = n;
      grid += n;
    }
    mlp_bwd_dw_kernel<<<grid, 512, DwSmem::alloc_bytes, st>>>(w);
    NWX_LAUNCHED();
    DwJobTable tab;
    for (int j = 0; j < kDwJobs; ++j) tab.job[j] = w.job[j];
    reduce_partials_kernel<<<(NWX_PARAMS_PER_NET + 255) / 256, 256, 0, st>>>(a.partial, tab, a.grad);
    NWX_LAUNCHED();
    // folded views layer -> d W_view[:, :256], d W_feature, d b_feature (fp32, against the master weights)
    if (!net.master) return NWX_E_NO_WEIGHTS;
    const DwJob& fj = w.job[kDwJobs - 1];
    fold_sum_kernel<<<kViewHidden, kHidden, 0, st>>>(a.partial, fj.ncta, fj.w_off, fj.b_off, a.fold_scratch);
    NWX_LAUNCHED();
    fold_grads_feature_kernel<<<kHidden / 4, 256, 0, st>>>(net.master + off[16], a.fold_scratch, a.grad + off[18], a.grad + off[19]);
    NWX_LAUNCHED();
    fold_grads_view_kernel<<<dim3(kHidden / 32, kViewHidden / 32), 256, 0, st>>>(net.master + off[18], net.master + off[19], a.fold_scratch,
                                                                                   a.grad + off[16]);
    NWX_LAUNCHED();
  }
  if (heads_st) NWX_CUDA_TRY(cudaStreamWaitEvent(st, join, 0));      // the network's gradient is final on `st` from here on
  return NWX_OK;
}

int launch_mse_grad(const float* rgb_c, const float* rgb_f, const float* gt, int64_t n_rays, float* d_c, float* d_f,
                    double* loss_scratch, double* loss_out, cudaStream_t st) {
  // loss_scratch: kMseScratchBytes, zero before the first use (the kernel leaves the ticket at zero)
  int64_t blocks = (n_rays * 3 + 255) / 256;
  if (blocks > kMseMaxBlocks) blocks = kMseMaxBlocks;
  mse_grad_kernel<<<(unsigned)blocks, 256, 0, st>>>(rgb_c, rgb_f, gt, n_rays * 3, d_c, d_f, loss_scratch + 1,
                                                    reinterpret_cast<unsigned int*>(loss_scratch), loss_out);
  NWX_LAUNCHED();
  return NWX_OK;
}

// Adam + re-pack of both networks: params / grads / m / v are [2][NWX_PARAMS_PER_NET] (coarse, then fine).
int launch_adam_pack(PackedNet (&nets)[2], float* params, const float* grads, float* m, float* v, float lr, float b1,
                     float b2, float eps, int step, float grad_scale, cudaStream_t st) {
  AdamPackArgs a{};
  for (int w = 0; w < 2; ++w) {
    PackedNet& n = nets[w];
    if (!n.wimg || !n.wimg_t || !n.gconsts || !n.wdir_t || !n.bview || !n.bview_fold) return NWX_E_NO_WEIGHTS;   // train_pack first
    const size_t o = (size_t)w * NWX_PARAMS_PER_NET;
    a.net[w] = AdamPackNet{params + o, grads + o, m + o, v + o, n.wimg, n.wimg_t, n.gconsts, n.wdir_t, n.bview};
  }
  for (int i = 0; i < NWX_NUM_WEIGHT_TENSORS; ++i) a.off[i] = g_flat.off[i];
  a.off[NWX_NUM_WEIGHT_TENSORS] = NWX_PARAMS_PER_NET;
  const float bc1 = 1.0f - powf(b1, (float)step), bc2 = 1.0f - powf(b2, (float)step);
  a.step = lr / bc1; a.inv_sqrt_bc2 = 0.0f /* filled on the device: rsqrtf(bc2) like adam_kernel */; a.bc2 = bc2; a.b1 = b1; a.b2 = b2; a.eps = eps; a.grad_scale = grad_scale;
  int64_t blocks = (2 * (int64_t)NWX_PARAMS_PER_NET + 255) / 256;
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  adam_pack_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  NWX_LAUNCHED();
  // folded views layer of both networks (reads the updated master parameters): one launch
  PackFoldPair f{};
  for (int w = 0; w < 2; ++w) {
    const float* p = params + (size_t)w * NWX_PARAMS_PER_NET;
    f.net[w] = PackFoldNet{p + g_flat.off[16], p + g_flat.off[18], p + g_flat.off[17], p + g_flat.off[19], nets[w].wimg,
                           nets[w].bview_fold, nets[w].wimg_t};
    nets[w].master = p;
    nets[w].loaded = true;
    nets[w].consts_stale = true;      // the host copy of the biases (inference launches) is now behind
  }
  launch_pack_fold_pair(f, st);
  NWX_LAUNCHED();
  return NWX_OK;
}

int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                int step, float grad_scale, cudaStream_t st) {
  const float bc1 = 1.0f - powf(b1, (float)step), bc2 = 1.0f - powf(b2, (float)step);
  int64_t blocks = (n + 255) / 256;
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, bc1, bc2, grad_scale);
  NWX_LAUNCHED();
  return NWX_OK;
}

}  // namespace nwx
