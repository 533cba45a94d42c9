// K3: fused positional encoding + 8x256 NeRF MLP on tcgen05 tensor cores.
//
// Replaces, in ONE persistent kernel, the whole of run_network (reference
// nerf/models/model_utils.py:13-30): pts = o + d*z, Embedding.embed of points (embedding.py:44-48),
// the per-point view-direction broadcast, batchify (utils/batch_utils.py:28-39) and
// NeRFModel.forward (nerf_model.py:45-83).  Nothing per-point is materialised in HBM: in = 4 B
// (z) per point, out = 16 B (raw rgb+sigma).
//
// Organisation (one CTA, or a CTA pair, per SM; 16 warps):
//   warp 0      TMA producer  : streams pre-swizzled bf16 weight K-blocks L2 -> smem ring
//                               (cp.async.bulk + mbarrier complete_tx)
//   warp 1      MMA issuer    : one elected lane issues tcgen05.mma (M=128|256, N=256|128, K=16),
//                               accumulators in TMEM; tcgen05.commit signals mbarriers
//   warp 2      TMEM allocator
//   warps 4-7   PE producers  : thread = point; position, 63 sin/cos features -> bf16, written
//                               straight into the swizzled A-operand tile in smem
//   warps 8-15  epilogue      : tcgen05.ld accumulator -> +bias, ReLU -> bf16 -> next layer's
//                               A-operand tile in smem (in place); fp32 sigma and rgb heads
// Two 128-point tiles are in flight per CTA (TMEM: 2 x 256 fp32 columns): while the tensor core
// runs layer l of tile B the epilogue warps drain layer l of tile A, so the tensor pipe only
// idles when an epilogue is slower than an MMA layer.
//
// Variants (template): kPair = cta_group::2 (CTA pair shares each weight K-block: CTA r holds
// output rows [r*N/2, (r+1)*N/2) of B, UMMA M = 256), kResident = a weight K-block stays in
// smem for both in-flight tiles (halves L2->smem traffic).  See DESIGN.md for the roofline.
//
// Numerics: bf16 operands (PE features, weights, hidden activations), fp32 accumulation and
// biases; the sigma head, the 27-d view-direction term of the views layer and the rgb head are
// evaluated in fp32 on CUDA cores (oracle model: mlp_forward_bf16_emul).
#include "mlp_device.cuh"

namespace nwx {

// chunk schedule: layer 5 is split so that a chunk never needs more than 4 resident K-blocks.
// kFold (inference): the feature layer (l = 8) is folded into the views layer (l = 9) at load time, so a
// tile runs 9 tensor-core layers = 10 chunks instead of 10 layers = 11 chunks (-11 % MMA work).
template <bool kFold> __device__ __forceinline__ constexpr int num_chunks() { return kFold ? 10 : 11; }
template <bool kFold> __device__ __forceinline__ constexpr int layers_per_tile() { return kFold ? 9 : 10; }
template <bool kFold> __device__ __forceinline__ int chunk_layer(int c) {
  const int l = c <= 5 ? c : c - 1;
  return (kFold && l == 8) ? 9 : l;
}
// position of a layer in a tile's sequence of accumulator hand-offs (mbarrier phase bookkeeping)
template <bool kFold> __device__ __forceinline__ int layer_seq(int l) { return (kFold && l == 9) ? 8 : l; }
__device__ __forceinline__ int chunk_kb0(int c) { return c == 6 ? 1 : 0; }
__device__ __forceinline__ int chunk_nkb(int c) { return (c == 0 || c == 5) ? 1 : 4; }
template <bool kFold>
__device__ __forceinline__ int layer_gkb0(int l) {     // global K-block index of a layer's first K-block
  if (kFold && l == 9) return kFoldKBlock0;
  return l == 0 ? 0 : (l <= 5 ? 1 + 4 * (l - 1) : 22 + 4 * (l - 6));
}

// sin / cos of 2^k * x for k = 0..9 without ten full-range sincosf calls.  With t = x / (2 pi) held as
// an unevaluated float pair (hi + lo), 2^k * t_hi is exact (power-of-two scaling), its fractional part
// f is exact, and adding 2^k * t_lo leaves the phase accurate to ~1e-7 turns even at k = 9; the
// reduced angle 2 pi f in [-pi, pi] then goes to MUFU.SIN/COS (abs error 2^-21.4, a tenth of a bf16
// half-ulp at |v| >= 2^-4).  ~7 instructions per sin/cos pair instead of ~40.
__device__ __forceinline__ void sincos_pow2(float t_hi, float t_lo, float scale, float& s, float& c) {
  const float a = t_hi * scale;                                   // exact
  const float f = (a - rintf(a)) + t_lo * scale;                  // fractional turns, |f| <= 0.5 + eps
  const float r = f * 6.283185307179586f;
  s = __sinf(r);
  c = __cosf(r);
}

// 63 positional-encoding features of one point (+1 zero pad) as 32 packed bf16 pairs.
// Column order of Embedding.embed (embedding.py:31-38,48): x/s, then per k: sin(x/s*2^k)(3), cos(..)(3).
__device__ __forceinline__ void encode_point(float px, float py, float pz, uint32_t (&pk)[32]) {
  float f[64];
  const float x[3] = {__fdiv_rn(px, 10.0f), __fdiv_rn(py, 10.0f), __fdiv_rn(pz, 10.0f)};   // scalar_factor 10
  float t_hi[3], t_lo[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {                                   // t = x / (2 pi) as hi + lo
    constexpr float kInv2PiHi = 0.15915494309189535f, kInv2PiLo = 6.4206382e-09f;   // 1/(2 pi) = hi + lo in fp32
    t_hi[a] = x[a] * kInv2PiHi;
    t_lo[a] = fmaf(x[a], kInv2PiHi, -t_hi[a]) + x[a] * kInv2PiLo;
  }
  f[0] = x[0]; f[1] = x[1]; f[2] = x[2];
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    const float sc = (float)(1 << k);            // exact power of two: x*2^k is exact
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float s, c;
      sincos_pow2(t_hi[a], t_lo[a], sc, s, c);
      f[3 + 6 * k + a] = s;
      f[3 + 6 * k + 3 + a] = c;
    }
  }
  f[63] = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; ++i) pk[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
}

// Training: biases/heads change every step, so they are refreshed by a stream-ordered
// device-to-device copy into the constant bank (no host round trip) instead of riding in the
// launch parameters.  (train.cu keeps its own copy for the backward kernels: __constant__
// symbols are per translation unit without relocatable device code.)
__constant__ MlpConsts c_fwd_train_consts[2];

int upload_fwd_train_consts(int which, const MlpConsts* dev_src, cudaStream_t st) {
  NWX_CUDA_TRY(cudaMemcpyToSymbolAsync(c_fwd_train_consts, dev_src, sizeof(MlpConsts), (size_t)which * sizeof(MlpConsts),
                                       cudaMemcpyDeviceToDevice, st));
  return NWX_OK;
}

enum { kEpiRelu = 0, kEpiReluSigma = 1, kEpiLinear = 2 };

// Hidden-layer epilogue of one thread (= one point / TMEM lane) over its 128-column half:
// accumulator -> +bias -> (ReLU) -> bf16 -> the next layer's swizzled A-operand tile.
// Returns this half's partial of the fp32 sigma head when kMode == kEpiReluSigma.
// One 32-column chunk of the hidden-layer epilogue: +bias -> (ReLU) -> bf16 -> swizzled A tile.
template <int kMode, bool kTap, bool kSave, int W>
__device__ __forceinline__ void epilogue_chunk(const MlpConsts& cst, int l, const uint32_t (&v)[W], int col, uint32_t hrow,
                                               int row, float* tap_row, uint8_t* grow /* mask words of (tile, layer) */,
                                               uint64_t& sig2 /* sigma head: (even, odd) column partial sums */) {
  static_assert(W == 32 || W == 16, "chunk of 32 columns (8 epilogue warps) or 16 (16 epilogue warps)");
  static_assert(!kSave || W == 32, "the training masks are one word per 32 columns");
  uint32_t pk[W / 2];
  const bool no_sts = kTap && hrow == 0;          // timing experiment (debug instantiation only)
  // `col` is warp-uniform (the warp index comes from a shuffle), so biases and head weights are read from the
  // constant bank through uniform registers (LDCU.128), four columns per load, and used as packed operands
  const uint4* bias4 = reinterpret_cast<const uint4*>(&cst.bias[l][col]);
  const uint4* wa4 = reinterpret_cast<const uint4*>(&cst.w_alpha[col]);
#pragma unroll
  for (int j = 0; j < W; j += 2) {
    // two columns per instruction: packed fp32x2 add (sm_100 FADD2)
    const uint4 bq = bias4[j >> 2];
    const uint64_t bias2 = (j & 2) ? ((uint64_t)bq.z | ((uint64_t)bq.w << 32)) : ((uint64_t)bq.x | ((uint64_t)bq.y << 32));
    const uint64_t ab = fadd2((uint64_t)v[j] | ((uint64_t)v[j + 1] << 32), bias2);
    const float a = f32x2_lo(ab), b = f32x2_hi(ab);
    if (kMode == kEpiReluSigma) {                  // sigma head on fp32 relu(h7), nerf_model.py:63
      const uint4 wq = wa4[j >> 2];
      const uint64_t w2 = (j & 2) ? ((uint64_t)wq.z | ((uint64_t)wq.w << 32)) : ((uint64_t)wq.x | ((uint64_t)wq.y << 32));
      sig2 = ffma2(f32x2(fmaxf(a, 0.f), fmaxf(b, 0.f)), w2, sig2);
    }
    if (kTap && tap_row) {
      tap_row[col + j] = kMode == kEpiLinear ? a : fmaxf(a, 0.f);
      tap_row[col + j + 1] = kMode == kEpiLinear ? b : fmaxf(b, 0.f);
    }
    pk[j >> 1] = (kMode == kEpiLinear) ? pack_bf16x2(a, b) : pack_bf16x2_relu(a, b);   // feature: no ReLU (:64)
  }
  const uint32_t kbase = hrow + (col >> 6) * kABlock;
  const int j0 = (col & 63) >> 3;
  if (!no_sts) {
#pragma unroll
    for (int q = 0; q < W / 8; ++q)
      st_shared_v4(kbase + (((j0 + q) ^ (row & 7)) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  } else if (pk[0] == 0x12345678u && pk[7] == 0x9abcdef0u) {
    st_shared_v4(kbase, pk[0], pk[3], pk[5], pk[7]);        // keeps the math alive
  }
  if constexpr (kSave && kMode != kEpiLinear && W == 32) {
   if (grow != nullptr) {
    // training: ReLU' of this row's 32 columns as one word for the dX kernel (bit j / 16 + j = low / high
    // half of packed word j is non-zero), so the backward reads 4 B instead of 64 B of activations here.
    // h >= +0 after the ReLU, so h + 0x7FFF carries into bit 15 exactly when h != 0.
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) m |= ((pk[j] + 0x7FFF7FFFu) >> (15 - j)) & (0x00010001u << j);
    reinterpret_cast<uint32_t*>(grow)[(col >> 5) * kTileM + row] = m;
   }
  }
}

// Hidden-layer epilogue of one thread (= one point / TMEM lane) over its 4 * W columns (8 epilogue warps: a
// 128-column half in chunks of 32; 16 epilogue warps: a 64-column quarter in chunks of 16), software pipelined
// two deep: the tcgen05.ld of chunk c+1 is in flight while chunk c is processed (the epilogue is latency-bound:
// 2-4 epilogue warps per scheduler), and the bias loads are free to move above the TMEM wait.  Returns this
// thread's partial of the fp32 sigma head (kEpiReluSigma).
template <int kMode, bool kTap, bool kSave, int W>
__device__ __forceinline__ float epilogue_hidden(const MlpConsts& cst, int l, uint32_t d_tmem, uint32_t hrow,
                                                 int row, int wg, float* tap_row, uint8_t* grow, bool dbg_half = false) {
  uint64_t sig = 0;                  // (even, odd) column partial sums of the sigma head, fp32 each
  const int col0 = wg * 4 * W;
  uint32_t va[W], vb[W];
  tmem_ld(d_tmem + col0, va);
  tmem_wait_ld_dep(va);
  tmem_ld(d_tmem + col0 + W, vb);
  epilogue_chunk<kMode, kTap, kSave, W>(cst, l, va, col0, hrow, row, tap_row, grow, sig);
  tmem_wait_ld_dep(vb);
  if (kTap && dbg_half) {            // timing experiment: half the epilogue (what twice the epilogue warps would leave per thread)
    epilogue_chunk<kMode, kTap, kSave, W>(cst, l, vb, col0 + W, hrow, row, tap_row, grow, sig);
    return f32x2_lo(sig) + f32x2_hi(sig);
  }
  tmem_ld(d_tmem + col0 + 2 * W, va);
  epilogue_chunk<kMode, kTap, kSave, W>(cst, l, vb, col0 + W, hrow, row, tap_row, grow, sig);
  tmem_wait_ld_dep(va);
  tmem_ld(d_tmem + col0 + 3 * W, vb);
  epilogue_chunk<kMode, kTap, kSave, W>(cst, l, va, col0 + 2 * W, hrow, row, tap_row, grow, sig);
  tmem_wait_ld_dep(vb);
  epilogue_chunk<kMode, kTap, kSave, W>(cst, l, vb, col0 + 3 * W, hrow, row, tap_row, grow, sig);
  return f32x2_lo(sig) + f32x2_hi(sig);
}

template <bool kPair, bool kResident, int kStages, bool kTap, bool kTrain, bool kFold, int kEpiWarps>
__global__ void __launch_bounds__(256 + 32 * kEpiWarps, 1)
mlp_fused_kernel(const __grid_constant__ MlpArgs args, const __grid_constant__ MlpConsts cst_param) {
  using L = SmemLayout<kPair, kStages>;
  const MlpConsts& cst = kTrain ? c_fwd_train_consts[args.which] : cst_param;
  // Timeline trace (debug instantiation only, nwx_debug_tap(ctx, -2, buf)): CTA 0 records
  // (tag, clock) pairs per role into buf[role][event][2] so the critical path can be read off.
  constexpr int kTraceEvents = 4096;
  const bool tracing = kTap && args.dbg_out != nullptr && args.dbg_layer == -2 && blockIdx.x == 0;
  uint32_t trace_n = 0;
  auto trace = [&](int role, int ev, int it, int l, int t) {
    if (kTap && tracing && trace_n < kTraceEvents) {
      uint32_t* rec = reinterpret_cast<uint32_t*>(args.dbg_out) + ((size_t)role * kTraceEvents + trace_n) * 2;
      rec[0] = (uint32_t)(ev * 100000 + it * 1000 + l * 10 + t);
      rec[1] = (uint32_t)clock64();
      ++trace_n;
    }
  };
  constexpr int kCG = kPair ? 2 : 1;
  constexpr int kEpiGroups = kEpiWarps / 4;              // column groups: 2 x 128 columns or 4 x 64 columns
  constexpr int kW = 256 / kEpiGroups / 4;               // epilogue chunk width (32 or 16 columns)
  static_assert(kEpiWarps == 8 || kEpiWarps == 16, "8 or 16 epilogue warps");
  static_assert(!(kTrain && kEpiWarps != 8), "the training forward (masks, TMA tile stores) is written for 8 epilogue warps");
  constexpr int kNumChunks = num_chunks<kFold>();
  constexpr int kSeqLen = layers_per_tile<kFold>();      // accumulator hand-offs per tile and iteration
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));

  // warp index through a shuffle so that the compiler knows it (and the column group / bias addresses derived
  // from it) is warp-uniform: the epilogue's biases then come straight from the constant bank through uniform
  // registers instead of one LDC per pair of columns
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0;
  const int units = kPair ? gridDim.x / 2 : gridDim.x;
  const int unit = kPair ? blockIdx.x / 2 : blockIdx.x;
  const int64_t P = args.P;
  const int iters = args.iters;
  auto tile_of = [&](int it, int t) -> int64_t {
    const int64_t u = ((int64_t)it * units + unit) * 2 + t;
    return kPair ? u * 2 + rank : u;
  };
  // barriers that the MMA issuer waits on live in the leader CTA (rank 0)
  auto leader = [&](uint32_t local) -> uint32_t { return kPair ? mapa(local, 0) : local; };

  // ------------------------------------------------------------------ setup ----
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(sbase + L::w_full + 8 * s, 1);
      mbar_init(sbase + L::w_empty + 8 * s, 1);
      mbar_init(sbase + L::w_peer + 8 * s, 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(sbase + L::acc_full + 8 * t, 1);
      mbar_init(sbase + L::a_ready + 8 * t, kEpiWarps * kCG);   // one arrive per epilogue warp (per CTA)
      mbar_init(sbase + L::pe_ready + 8 * t, 4 * kCG);    // one arrive per PE warp (per CTA)
      mbar_init(sbase + L::pe_free + 8 * t, 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<kCG>(sbase + L::tmem_slot, 512);
  tc_fence_before();
  if (kPair) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sgen + L::tmem_slot);

  if (warp == 0) {
    // =========================================================== TMA producer ====
    if (lane == 0) {
      const WaitCtx wc{args.diag, 0x100u};
      uint32_t fill = 0;
      for (int it = 0; it < iters; ++it) {
        for (int c = 0; c < kNumChunks; ++c) {
          const int l = chunk_layer<kFold>(c);
          const uint32_t bytes = (l == 9 ? kKBlockBytes / 2 : kKBlockBytes) / kCG;
          for (int t = 0; t < (kResident ? 1 : 2); ++t) {
            for (int kb = 0; kb < chunk_nkb(c); ++kb, ++fill) {
              const uint32_t stage = fill % kStages, round = fill / kStages;
              trace(0, 31, it, l, kb);
              mbar_wait(sbase + L::w_empty + 8 * stage, (round & 1) ^ 1, wc);
              trace(0, 32, it, l, kb);
              const uint32_t bar = sbase + L::w_full + 8 * stage;
              mbar_arrive_expect_tx(bar, bytes);
              const uint8_t* src = args.wimg + kblock_offset(layer_gkb0<kFold>(l) + chunk_kb0(c) + kb) + rank * bytes;
              bulk_g2s(sbase + L::w0 + stage * L::kStageBytes, src, bytes, bar);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================= MMA issuer ====
    if (lane == 0) {
      const WaitCtx wc{args.diag, 0x200u};
      uint32_t fill = 0;
      if (rank == 0) {
        for (int it = 0; it < iters; ++it) {
          for (int c = 0; c < kNumChunks; ++c) {
            const int l = chunk_layer<kFold>(c);
            const int nkb = chunk_nkb(c);
            const uint32_t seq = (uint32_t)it * kSeqLen + layer_seq<kFold>(l);   // index of this accumulator hand-off
            const bool first_chunk = (c != 6), last_chunk = (c != 5);
            const uint32_t idesc = umma_idesc_bf16(kTileM * kCG, l == 9 ? kViewHidden : kHidden);
            for (int t = 0; t < 2; ++t) {
              trace(1, 1, it, l, t);
              if (first_chunk) {
                if (l == 0) mbar_wait(sbase + L::pe_ready + 8 * t, it & 1, wc);
                // a_ready[t] completes once per layer epilogue: phase index = seq - 1
                if (seq != 0) mbar_wait(sbase + L::a_ready + 8 * t, (seq - 1) & 1, wc);
                tc_fence_after();
              }
              trace(1, 2, it, l, t);
              const uint32_t d_tmem = tmem_base + t * kHidden;
              for (int kb = 0; kb < nkb; ++kb) {
                const uint32_t f = kResident ? fill + kb : fill++;
                const uint32_t stage = f % kStages, round = f / kStages;
                // timing experiment (debug instantiation, dbg_layer == -6): never wait for weights after the
                // first iteration -- how much of the frame is the tensor pipe waiting on the weight ring?
                const bool skip_w = kTap && args.dbg_layer == -6 && args.dbg_out != nullptr && it > 0;
                if ((!kResident || t == 0) && !skip_w) {
                  mbar_wait(sbase + L::w_full + 8 * stage, round & 1, wc);
                  if (kPair) mbar_wait(sbase + L::w_peer + 8 * stage, round & 1, wc);
                  tc_fence_after();
                }
                const int akb = chunk_kb0(c) + kb;          // K-block index within the layer
                uint32_t a_addr;
                if (l == 0 || (l == 5 && akb == 0)) a_addr = sbase + L::pe0 + t * kABlock;
                else a_addr = sbase + L::h0 + t * kHBytes + (l == 5 ? akb - 1 : akb) * kABlock;
                const uint64_t adesc = umma_desc_k_sw128(a_addr);
                const uint64_t bdesc = umma_desc_k_sw128(sbase + L::w0 + stage * L::kStageBytes);
#pragma unroll
                for (int k = 0; k < 4; ++k)                 // 4 x (K = 16) per 64-wide K-block: +32 B
                  umma_bf16<kCG>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (akb | k) != 0);
                if (!kResident || t == 1) umma_commit<kCG>(sbase + L::w_empty + 8 * stage);
              }
              if (last_chunk) umma_commit<kCG>(sbase + L::acc_full + 8 * t);
              if (c == 5) umma_commit<kCG>(sbase + L::pe_free + 8 * t);
              trace(1, 3, it, l, t);
            }
            if (kResident) fill += nkb;
          }
        }
      } else {
        // peer CTA of a pair: relay "my half of the weight K-block has landed" to the leader
        constexpr uint32_t kBlocksPerTile = kFold ? kNumKBlocks - 4 : kNumKBlocks;
        const uint32_t per_iter = kResident ? kBlocksPerTile : 2 * kBlocksPerTile;
        for (uint32_t n = 0; n < per_iter * (uint32_t)iters; ++n, ++fill) {
          const uint32_t stage = fill % kStages, round = fill / kStages;
          mbar_wait(sbase + L::w_full + 8 * stage, round & 1, wc);
          mbar_arrive_cluster(mapa(sbase + L::w_peer + 8 * stage, 0));
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ============================================================ PE producers ====
    const WaitCtx wc{args.diag, 0x300u};
    const int row = (warp - 4) * 32 + lane;
    for (int it = 0; it < iters; ++it) {
      for (int t = 0; t < 2; ++t) {
        int64_t p = tile_of(it, t) * kTileM + row;
        if (p >= P) p = P - 1;                                  // tail tile: recompute a valid point
        float px = 0.f, py = 0.f, pz = 0.f;
        if (args.embedded) {
          // handled below: features are read, not computed
        } else if (args.pts) {
          px = __ldg(args.pts + p * 3 + 0); py = __ldg(args.pts + p * 3 + 1); pz = __ldg(args.pts + p * 3 + 2);
        } else {
          const int64_t ray = (p >> 32) == 0 ? (int64_t)((uint32_t)p / (uint32_t)args.S) : p / args.S;   // 32-bit divide when it fits
          const float* r = args.rays + ray * args.ray_dim;
          const float zz = __ldg(args.z + p);
          px = __fadd_rn(__ldg(r + 0), __fmul_rn(__ldg(r + 3), zz));   // inference handler:223
          py = __fadd_rn(__ldg(r + 1), __fmul_rn(__ldg(r + 4), zz));
          pz = __fadd_rn(__ldg(r + 2), __fmul_rn(__ldg(r + 5), zz));
        }
        uint32_t pk[32];
        if (args.embedded) {
          const float* e = args.embedded + p * (kPeXyz + kPeDir);
#pragma unroll
          for (int i = 0; i < 31; ++i) pk[i] = pack_bf16x2(__ldg(e + 2 * i), __ldg(e + 2 * i + 1));
          pk[31] = pack_bf16x2(__ldg(e + 62), 0.0f);
        } else {
          encode_point(px, py, pz, pk);
        }
        if (row == 0) trace(2, 21, it, 0, t);
        if (it > 0) mbar_wait(sbase + L::pe_free + 8 * t, (it - 1) & 1, wc);
        if (row == 0) trace(2, 22, it, 0, t);
        const uint32_t dst = sbase + L::pe0 + t * kABlock + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(dst + ((j ^ (row & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (kPair) mbar_arrive_cluster(leader(sbase + L::pe_ready + 8 * t));
          else mbar_arrive(sbase + L::pe_ready + 8 * t);
        }
        if (kTrain && args.acts != nullptr && tile_of(it, t) < args.n_tiles) {
          uint8_t* g = args.acts + tile_img_offset(act_slot_kb0(0), 1, args.n_tiles, tile_of(it, t), 0) + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(g + ((j ^ (row & 7)) << 4)) =
                make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
      }
    }
  } else if (warp >= 8) {
    // ================================================================ epilogue ====
    const WaitCtx wc{args.diag, 0x400u};
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
    const int wg = (warp - 8) >> 2;               // column group
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    float sig0 = 0.f, sig1 = 0.f;                 // sigma-head partials of tile 0 / tile 1
    for (int it = 0; it < iters; ++it) {
      for (int l = 0; l < kNumLayers; ++l) {
        if (kFold && l == 8) continue;                         // folded into the views layer
        const uint32_t seq = (uint32_t)it * kSeqLen + layer_seq<kFold>(l);
        for (int t = 0; t < 2; ++t) {
          if (lane == 0 && quad == 0) trace(3 + wg, 11, it, l, t);
          if (l == 9) {
            // views layer: this row's per-ray term (two 128 B lines of dirbias) is needed the moment the
            // accumulator is ready -- pull it into L1 while the tensor pipe is still working on it
            const int64_t pp = tile_of(it, t) * kTileM + row, pq = pp < P ? pp : P - 1;
            const int64_t ry = (pq >> 32) == 0 ? (int64_t)((uint32_t)pq / (uint32_t)args.S) : pq / args.S;
            const float* pf = args.dirbias + ry * kViewHidden + wg * (kViewHidden / kEpiGroups);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
            if (kEpiGroups == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(pf + 32));
          }
          mbar_wait(sbase + L::acc_full + 8 * t, seq & 1, wc);   // phase index = seq
          tc_fence_after();
          if (lane == 0 && quad == 0) trace(3 + wg, 12, it, l, t);
          const uint32_t d_tmem = lane_addr + t * kHidden;
          const int64_t p = tile_of(it, t) * kTileM + row;
          if (kTrain) {
            // the TMA store that saved this buffer's previous contents must have finished reading it
            // (groups complete in order: all but the newest one, which belongs to the other tile)
            if (warp == 8 && lane == 0 && NWX_EXP(args) == 0) bulk_wait_read<1>();
            if (NWX_EXP(args) != 14) named_bar_sync(3, 256);
          }
          float* tap_row = nullptr;
          if (kTap && args.dbg_out != nullptr && args.dbg_layer == l && p < P) tap_row = args.dbg_out + p * kHidden;
          if (kTap && args.dbg_layer == -4 && args.dbg_out != nullptr) {
            // timing experiment: no epilogue work at all, only the barrier handshake
          } else if (l < 9) {
            const uint32_t hrow = (kTap && args.dbg_layer == -3 && args.dbg_out != nullptr)
                                      ? 0u : sbase + L::h0 + t * kHBytes + row * 128;
            uint8_t* mask_row = nullptr;         // training: ReLU' bit words of (tile, layer l), see mask_img_offset
            if (kTrain && args.masks != nullptr && l < 8 && NWX_EXP(args) != 13)
              mask_row = reinterpret_cast<uint8_t*>(args.masks) +
                         mask_img_offset(tile_of(it, t) < args.n_tiles ? tile_of(it, t) : args.n_tiles, l);
            const bool dbg_half = kTap && args.dbg_layer == -7 && args.dbg_out != nullptr;
            if (l == 7) {
              const float sig = epilogue_hidden<kEpiReluSigma, kTap, kTrain, kW>(cst, l, d_tmem, hrow, row, wg, tap_row, mask_row, dbg_half);
              if (t == 0) sig0 = sig; else sig1 = sig;
            } else if (l == 8) {
              epilogue_hidden<kEpiLinear, kTap, false, kW>(cst, l, d_tmem, hrow, row, wg, tap_row, nullptr, dbg_half);
            } else {
              epilogue_hidden<kEpiRelu, kTap, kTrain, kW>(cst, l, d_tmem, hrow, row, wg, tap_row, mask_row, dbg_half);
            }
            fence_proxy_async_smem();            // my smem writes -> visible to the next layer's UMMA / TMA store
          } else {
            // views layer (N = 128): + (b_view + W_view[:,256:] . pe(dir)), ReLU, fp32 rgb head
            const int64_t pc = p < P ? p : P - 1;
            const int64_t dray = (pc >> 32) == 0 ? (int64_t)((uint32_t)pc / (uint32_t)args.S) : pc / args.S;
            const float* db = args.dirbias + dray * kViewHidden;
            uint64_t r2 = 0, g2 = 0, b2 = 0;     // rgb head partial sums over (even, odd) columns, packed fp32x2
            constexpr int kViewCols = kViewHidden / kEpiGroups;    // 64 or 32 columns per thread
            auto pair = [](uint32_t lo, uint32_t hi) -> uint64_t { return (uint64_t)lo | ((uint64_t)hi << 32); };
#pragma unroll 1
            for (int cc = 0; cc < kViewCols / 32; ++cc) {
              const int col = wg * kViewCols + cc * 32;
              uint32_t v[32];
              tmem_ld32(d_tmem + col, v);
              tmem_wait_ld_dep(v);               // dependency form: the dirbias loads may be scheduled above the wait
              const uint4* wr4 = reinterpret_cast<const uint4*>(&cst.w_rgb[0][col]);
              const uint4* wg4 = reinterpret_cast<const uint4*>(&cst.w_rgb[1][col]);
              const uint4* wb4 = reinterpret_cast<const uint4*>(&cst.w_rgb[2][col]);
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const uint4 d4 = __ldg(reinterpret_cast<const uint4*>(db + col + j));
                const uint4 wr = wr4[j >> 2], wgq = wg4[j >> 2], wb = wb4[j >> 2];
                float dd[4];
#pragma unroll
                for (int q = 0; q < 4; q += 2) {
                  const uint64_t pre = fadd2(pair(v[j + q], v[j + q + 1]), q ? pair(d4.z, d4.w) : pair(d4.x, d4.y));   // nerf_model.py:68-70
                  dd[q] = fmaxf(f32x2_lo(pre), 0.f);
                  dd[q + 1] = fmaxf(f32x2_hi(pre), 0.f);
                  if (kTap && tap_row) { tap_row[col + j + q] = dd[q]; tap_row[col + j + q + 1] = dd[q + 1]; }
                  const uint64_t hv2 = f32x2(dd[q], dd[q + 1]);
                  r2 = ffma2(hv2, q ? pair(wr.z, wr.w) : pair(wr.x, wr.y), r2);                 // :74
                  g2 = ffma2(hv2, q ? pair(wgq.z, wgq.w) : pair(wgq.x, wgq.y), g2);
                  b2 = ffma2(hv2, q ? pair(wb.z, wb.w) : pair(wb.x, wb.y), b2);
                }
                if (kTrain && args.hv_out != nullptr && p < P && NWX_EXP(args) != 15)      // post-ReLU views hidden, for the backward
                  *reinterpret_cast<float4*>(args.hv_out + p * kViewHidden + col + j) = make_float4(dd[0], dd[1], dd[2], dd[3]);
              }
            }
            const float r = f32x2_lo(r2) + f32x2_hi(r2), g = f32x2_lo(g2) + f32x2_hi(g2), b = f32x2_lo(b2) + f32x2_hi(b2);
            // combine the column groups through the (now dead) activation tile, then store
            float4* scratch = reinterpret_cast<float4*>(sgen + L::h0 + t * kHBytes);
            const float sig = t == 0 ? sig0 : sig1;
            if (wg != 0) scratch[(wg - 1) * kTileM + row] = make_float4(r, g, b, sig);
            named_bar_sync(1, 32 * kEpiWarps);
            if (wg == 0 && p < P) {
              float4 o = scratch[row];
#pragma unroll
              for (int gi = 1; gi < kEpiGroups - 1; ++gi) {
                const float4 o2 = scratch[gi * kTileM + row];
                o.x += o2.x; o.y += o2.y; o.z += o2.z; o.w += o2.w;
              }
              float4 out;
              out.x = r + o.x + cst.b_rgb[0];
              out.y = g + o.y + cst.b_rgb[1];
              out.z = b + o.z + cst.b_rgb[2];
              out.w = sig + o.w + cst.b_alpha;
              reinterpret_cast<float4*>(args.raw_out)[p] = out;                        // (rgb, sigma), :76
            }
          }
          tc_fence_before();                     // my tcgen05.ld's are complete before the MMA overwrites D
          __syncwarp();
          if (lane == 0) {
            if (kPair) mbar_arrive_cluster(leader(sbase + L::a_ready + 8 * t));
            else mbar_arrive(sbase + L::a_ready + 8 * t);
            if (quad == 0) trace(3 + wg, 13, it, l, t);
          }
          if (kTrain && l < 9) {
            // training: the activation tile just written IS the image the backward wants -- one TMA
            // store of the whole 64 KB tile instead of 16 global stores per thread
            if (NWX_EXP(args) != 14) named_bar_sync(3, 256);              // every warp's tile writes are fenced
            if (warp == 8 && lane == 0 && args.acts != nullptr && tile_of(it, t) < args.n_tiles && NWX_EXP(args) != 12)
              bulk_s2g(args.acts + tile_img_offset(act_slot_kb0(l + 1), 4, args.n_tiles, tile_of(it, t), 0),
                       sbase + L::h0 + t * kHBytes, kHBytes);
          }
        }
      }
    }
    if (kTrain && warp == 8 && lane == 0) bulk_wait_read<0>();   // smem must outlive the last store's reads
  }

  // --------------------------------------------------------------- teardown ----
  tc_fence_before();
  if (kPair) cluster_sync(); else __syncthreads();
  if (warp == 2) tmem_dealloc<kCG>(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// Weight packing: fp32 nn.Linear weights [out, in] -> bf16 K-block images in the exact byte
// layout the UMMA B descriptor expects (K-major, SWIZZLE_128B: row n at n*128 B, 16-byte chunk c
// of the row stored at chunk position c ^ (n & 7)), so the kernel moves them with flat bulk copies.
// ------------------------------------------------------------------------------------------------
struct PackSrc {
  const float* w[kNumLayers];     // pts 0..7, feature, views
};

__global__ void pack_weights_kernel(PackSrc src, uint8_t* __restrict__ wimg) {
  const int g = blockIdx.x;                       // global K-block
  int l = 0, kb = g;
  while (kb >= layer_kblocks(l)) { kb -= layer_kblocks(l); ++l; }
  const int rows = (l == 9) ? kViewHidden : kHidden;
  const int in_dim = (l == 0) ? kPeXyz : (l == 5 ? kPeXyz + kHidden : (l == 9 ? kHidden + kPeDir : kHidden));
  __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(wimg + kblock_offset(g));
  for (int e = threadIdx.x; e < rows * 64; e += blockDim.x) {
    const int n = e >> 6, c = e & 63;
    int k;                                        // source input column, -1 = zero padding
    if (l == 0) k = (c < kPeXyz) ? c : -1;
    else if (l == 5) k = (kb == 0) ? ((c < kPeXyz) ? c : -1) : kPeXyz + (kb - 1) * 64 + c;   // [pe | h], nerf_model.py:59
    else k = kb * 64 + c;                         // views: only the 256 feature columns (:66)
    const float v = (k >= 0) ? src.w[l][(size_t)n * in_dim + k] : 0.0f;
    const int chunk = (c >> 3) ^ (n & 7);
    img[n * 64 + chunk * 8 + (c & 7)] = __float2bfloat16_rn(v);
  }
}

__global__ void pack_dir_kernel(const float* __restrict__ wv, float* __restrict__ wdir_t) {
  // wdir_t[i][j] = W_view[j][256 + i]
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < kPeDir * kViewHidden; e += gridDim.x * blockDim.x) {
    const int i = e / kViewHidden, j = e - i * kViewHidden;
    wdir_t[e] = wv[(size_t)j * (kHidden + kPeDir) + kHidden + i];
  }
}

// Folded views layer (inference): W_fold[j][k] = sum_m W_view[j][m] * W_feature[m][k]  (128 x 256, fp64
// accumulation, rounded once to fp32 and then to bf16), written as 4 half-size K-block images; and
// b_fold[j] = b_view[j] + sum_m W_view[j][m] * b_feature[m].  hv = relu(W_fold h8 + b_fold + W_view[:,256:] pe(dir))
// is the same function as nerf_model.py:64-70 with one bf16 rounding fewer (no rounded `feature`).
__device__ __forceinline__ void pack_fold_body(const float* __restrict__ wv, const float* __restrict__ wf,
                                               const float* __restrict__ bv, const float* __restrict__ bf,
                                               uint8_t* __restrict__ wimg, float* __restrict__ bview_fold,
                                               uint8_t* __restrict__ wimg_t /* training: transposed image for dX, or nullptr */) {
  const int kb = blockIdx.x;                       // K-block of the folded layer (64 input columns)
  const int n0 = blockIdx.y * 2;                   // 2 output rows per block: 256 blocks, one output per thread
  constexpr int kIn = kHidden + kPeDir;
  __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(wimg + kblock_offset(kFoldKBlock0 + kb));
  for (int e = threadIdx.x; e < 2 * 64; e += blockDim.x) {
    const int n = n0 + (e >> 6), c = e & 63, k = kb * 64 + c;
    double a4[4] = {0.0, 0.0, 0.0, 0.0};          // four independent chains: the loop is latency-bound
#pragma unroll 4
    for (int m = 0; m < kHidden; m += 4)
#pragma unroll
      for (int q = 0; q < 4; ++q) a4[q] += (double)wv[(size_t)n * kIn + m + q] * (double)wf[(size_t)(m + q) * kHidden + k];
    const double acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
    const int chunk = (c >> 3) ^ (n & 7);
    const __nv_bfloat16 w = __float2bfloat16_rn((float)acc);
    img[n * 64 + chunk * 8 + (c & 7)] = w;
    if (wimg_t) {
      // dX step 0: d h8[p][i] = sum_o G_v[p][o] W_fold[o][i]  =>  B[row i][k = o]; K-block o / 64 of [256 x 64]
      __nv_bfloat16* timg = reinterpret_cast<__nv_bfloat16*>(wimg_t + (size_t)(n >> 6) * kKBlockBytes);
      const int o = n & 63;
      timg[k * 64 + (((o >> 3) ^ (k & 7)) << 3) + (o & 7)] = w;
    }
  }
  if (kb == 0 && threadIdx.x < 2) {
    const int n = n0 + threadIdx.x;
    double acc = (double)bv[n];
    for (int m = 0; m < kHidden; ++m) acc += (double)wv[(size_t)n * kIn + m] * (double)bf[m];
    bview_fold[n] = (float)acc;
  }
}
__global__ void pack_fold_kernel(const float* __restrict__ wv, const float* __restrict__ wf,
                                 const float* __restrict__ bv, const float* __restrict__ bf,
                                 uint8_t* __restrict__ wimg, float* __restrict__ bview_fold, uint8_t* __restrict__ wimg_t) {
  pack_fold_body(wv, wf, bv, bf, wimg, bview_fold, wimg_t);
}

__global__ void pack_fold_pair_kernel(const __grid_constant__ PackFoldPair f) {
  const PackFoldNet& n = f.net[blockIdx.z];
  pack_fold_body(n.wv, n.wf, n.bv, n.bf, n.wimg, n.bview_fold, n.wimg_t);
}
void launch_pack_fold_pair(const PackFoldPair& f, cudaStream_t st) {
  pack_fold_pair_kernel<<<dim3(4, kViewHidden / 2, 2), 128, 0, st>>>(f);
}

// Device-side part of packing (no host synchronisation): swizzled bf16 K-block images, the transposed
// fp32 view-direction weights and the views bias.  t: 24 device pointers in state_dict order:
// pts.{0..7}.{w,b} (0..15), views.{w,b} (16,17), feature (18,19), alpha (20,21), rgb (22,23).
int pack_network_images(PackedNet& net, const float* const* t, bool with_fold, cudaStream_t st, uint8_t* fold_t) {
  if (!net.wimg) NWX_CUDA_TRY(cudaMalloc(&net.wimg, kWeightImageBytes));
  if (!net.wdir_t) NWX_CUDA_TRY(cudaMalloc(&net.wdir_t, sizeof(float) * kPeDir * kViewHidden));
  if (!net.bview) NWX_CUDA_TRY(cudaMalloc(&net.bview, sizeof(float) * kViewHidden));
  if (!net.bview_fold) NWX_CUDA_TRY(cudaMalloc(&net.bview_fold, sizeof(float) * kViewHidden));
  PackSrc src;
  for (int i = 0; i < 8; ++i) src.w[i] = t[2 * i];
  src.w[8] = t[18];
  src.w[9] = t[16];
  pack_weights_kernel<<<kNumKBlocks, 256, 0, st>>>(src, net.wimg);
  NWX_LAUNCHED();
  pack_dir_kernel<<<4, 256, 0, st>>>(t[16], net.wdir_t);
  NWX_LAUNCHED();
  NWX_CUDA_TRY(cudaMemcpyAsync(net.bview, t[17], sizeof(float) * kViewHidden, cudaMemcpyDeviceToDevice, st));
  if (with_fold) {
    pack_fold_kernel<<<dim3(4, kViewHidden / 2), 128, 0, st>>>(t[16], t[18], t[17], t[19], net.wimg, net.bview_fold, fold_t);
    NWX_LAUNCHED();
  }
  return NWX_OK;
}

int pack_network(PackedNet& net, const float* const* t, cudaStream_t st) {
  int rc = pack_network_images(net, t, true, st);
  if (rc) return rc;
  MlpConsts& c = net.consts;
  for (int i = 0; i < 8; ++i)
    NWX_CUDA_TRY(cudaMemcpyAsync(c.bias[i], t[2 * i + 1], sizeof(float) * kHidden, cudaMemcpyDeviceToHost, st));
  NWX_CUDA_TRY(cudaMemcpyAsync(c.bias[8], t[19], sizeof(float) * kHidden, cudaMemcpyDeviceToHost, st));
  NWX_CUDA_TRY(cudaMemcpyAsync(c.w_alpha, t[20], sizeof(float) * kHidden, cudaMemcpyDeviceToHost, st));
  NWX_CUDA_TRY(cudaMemcpyAsync(&c.b_alpha, t[21], sizeof(float), cudaMemcpyDeviceToHost, st));
  NWX_CUDA_TRY(cudaMemcpyAsync(c.w_rgb, t[22], sizeof(float) * 3 * kViewHidden, cudaMemcpyDeviceToHost, st));
  NWX_CUDA_TRY(cudaMemcpyAsync(c.b_rgb, t[23], sizeof(float) * 3, cudaMemcpyDeviceToHost, st));
  NWX_CUDA_TRY(cudaStreamSynchronize(st));       // load-time only: the consts are a host-side launch argument
  net.loaded = true;
  net.consts_stale = false;
  return NWX_OK;
}

// dirbias[n][j] = b_view[j] + sum_i W_view[j][256+i] * pe_dir(dir_n)[i]   (fp32)
// pe_dir = Embedding(num_freqs=4, scalar_factor=1).embed  (embedding.py:44-48)
// Register-tiled: a warp owns 32 rays and all 128 columns; lane = four adjacent columns with their 4 x 27 weights in
// registers.  The warp's embeddings sit feature-major in its own shared-memory slice ([feature][ray]); per feature ONE
// broadcast LDS.128 brings four rays and feeds 16 FMAs (8 packed fp32x2, the weight as the broadcast scalar), so the
// shared-memory return path (the limit of a thread-per-column layout: one LDS.128 per 4 FMAs) is no longer the bound.
// Per (ray, column) the sum runs i = 0 .. 26 from the bias, each term one IEEE fma; one STG.128 per ray and lane.
constexpr int kDirbiasWarps = 4, kDirbiasRays = 32 * kDirbiasWarps;
__global__ void __launch_bounds__(32 * kDirbiasWarps)
dirbias_kernel(const float* __restrict__ dirs, int stride, int64_t n, int pre_embedded,
               const float* __restrict__ wdir_t, const float* __restrict__ bview, float* __restrict__ out) {
  constexpr int kF = 27;
  __shared__ float4 pe4[kDirbiasWarps][kF][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t base = (int64_t)blockIdx.x * kDirbiasRays + warp * 32;
  if (base >= n) return;                                         // no block-wide barrier below
  float* pe = reinterpret_cast<float*>(&pe4[warp][0][0]);       // [feature][ray of the warp]
  {
    const int64_t ray = base + lane < n ? base + lane : n - 1;
    const float* d = dirs + ray * stride;
    if (pre_embedded) {
#pragma unroll
      for (int i = 0; i < kF; ++i) pe[i * 32 + lane] = __ldg(d + i);
    } else {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float x = __ldg(d + a);
        pe[a * 32 + lane] = x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                            // sin and cos of 2^k x from one range reduction
          float sn, cs;
          sincosf(x * (float)(1 << k), &sn, &cs);
          pe[(3 + 6 * k + a) * 32 + lane] = sn;
          pe[(6 + 6 * k + a) * 32 + lane] = cs;
        }
      }
    }
  }
  float4 w[kF];
#pragma unroll
  for (int i = 0; i < kF; ++i) w[i] = __ldg(reinterpret_cast<const float4*>(wdir_t + i * kViewHidden) + lane);
  const float4 b = __ldg(reinterpret_cast<const float4*>(bview) + lane);
  const uint32_t pe_addr = smem_u32(pe);
  __syncwarp();
#pragma unroll 1
  for (int g = 0; g < 8; ++g) {                                  // four rays at a time
    uint64_t acc[4][2];                                          // [column][ray pair]
    acc[0][0] = acc[0][1] = f32x2(b.x, b.x);
    acc[1][0] = acc[1][1] = f32x2(b.y, b.y);
    acc[2][0] = acc[2][1] = f32x2(b.z, b.z);
    acc[3][0] = acc[3][1] = f32x2(b.w, b.w);
#pragma unroll
    for (int i = 0; i < kF; ++i) {
      uint64_t v0, v1;                                           // feature i of rays (4g, 4g + 1), (4g + 2, 4g + 3)
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v0), "=l"(v1) : "r"(pe_addr + (uint32_t)((i * 8 + g) * 16)) : "memory");
      const float wc[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        acc[c][0] = ffma2(f32x2(wc[c], wc[c]), v0, acc[c][0]);
        acc[c][1] = ffma2(f32x2(wc[c], wc[c]), v1, acc[c][1]);
      }
    }
    float4* o = reinterpret_cast<float4*>(out + (base + 4 * g) * kViewHidden) + lane;
    const int64_t left = n - (base + 4 * g);
    if (left > 0) o[0] = make_float4(f32x2_lo(acc[0][0]), f32x2_lo(acc[1][0]), f32x2_lo(acc[2][0]), f32x2_lo(acc[3][0]));
    if (left > 1) o[32] = make_float4(f32x2_hi(acc[0][0]), f32x2_hi(acc[1][0]), f32x2_hi(acc[2][0]), f32x2_hi(acc[3][0]));
    if (left > 2) o[64] = make_float4(f32x2_lo(acc[0][1]), f32x2_lo(acc[1][1]), f32x2_lo(acc[2][1]), f32x2_lo(acc[3][1]));
    if (left > 3) o[96] = make_float4(f32x2_hi(acc[0][1]), f32x2_hi(acc[1][1]), f32x2_hi(acc[2][1]), f32x2_hi(acc[3][1]));
  }
}

int launch_dirbias(const PackedNet& net, const float* dirs, int stride, int64_t n, bool pre_embedded, bool fold,
                   float* out, cudaStream_t st) {
  if (n == 0) return NWX_OK;
  dirbias_kernel<<<(unsigned)((n + kDirbiasRays - 1) / kDirbiasRays), kViewHidden, 0, st>>>(dirs, stride, n, pre_embedded ? 1 : 0, net.wdir_t,
                                                                 fold ? net.bview_fold : net.bview, out);
  NWX_LAUNCHED();
  return NWX_OK;
}

template <bool kPair, bool kResident, int kStages, bool kTap, bool kTrain = false, bool kFold = false, int kEpiWarps = 8>
static int launch_variant(const PackedNet& net, MlpArgs args, cudaStream_t st) {
  using Lay = SmemLayout<kPair, kStages>;
  auto kern = mlp_fused_kernel<kPair, kResident, kStages, kTap, kTrain, kFold, kEpiWarps>;
  static PerDeviceOnce configured;                 // the attribute is per device
  const int dev = current_device();
  if (configured.need(dev)) {
    NWX_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay::alloc_bytes));
    configured.done(dev);
  }
  const int64_t tiles = (args.P + kTileM - 1) / kTileM;
  args.n_tiles = tiles;
  int ctas = num_sms();
  if (kPair) ctas &= ~1;
  const int units = kPair ? ctas / 2 : ctas;
  const int tiles_per_unit_iter = kPair ? 4 : 2;
  // do not launch more units than there is work for
  int64_t need_units = (tiles + tiles_per_unit_iter - 1) / tiles_per_unit_iter;
  int use_units = (int)(need_units < units ? need_units : units);
  if (use_units < 1) use_units = 1;
  args.iters = (int)((tiles + (int64_t)use_units * tiles_per_unit_iter - 1) / ((int64_t)use_units * tiles_per_unit_iter));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kPair ? use_units * 2 : use_units);
  cfg.blockDim = dim3(256 + 32 * kEpiWarps);
  cfg.dynamicSmemBytes = Lay::alloc_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NWX_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, args, net.consts));
  g_nwx_launches.fetch_add(1, std::memory_order_relaxed);
  return NWX_OK;
}

// variant: 0/1 = CTA pair + resident weights + folded feature layer (production), 2 = CTA pair streaming,
//          3 = single CTA streaming (cta_group::1), 4 = as 1 with the reference's layer structure (no fold)
int launch_mlp(const PackedNet& net, MlpArgs args, int variant, cudaStream_t st) {
  if (args.P <= 0) return NWX_OK;
  const bool tap = args.dbg_out != nullptr;
  switch (variant) {
    case 0:
    case 1: return tap ? launch_variant<true, true, 4, true, false, true>(net, args, st)
                       : launch_variant<true, true, 4, false, false, true>(net, args, st);
    case 4: return tap ? launch_variant<true, true, 4, true>(net, args, st) : launch_variant<true, true, 4, false>(net, args, st);
    case 2: return tap ? launch_variant<true, false, 4, true>(net, args, st) : launch_variant<true, false, 4, false>(net, args, st);
    case 3: return tap ? launch_variant<false, false, 2, true>(net, args, st) : launch_variant<false, false, 2, false>(net, args, st);
    default: return NWX_E_INVALID;
  }
}

// Training forward: same kernel (CTA pair, resident weights, folded feature layer -- re-folded from the master
// weights by every train_pack), biases/heads read from device memory (they change every step), tensor-core
// operands saved as tile images for the backward.
int launch_mlp_train_forward(const PackedNet& net, MlpArgs args, cudaStream_t st) {
  if (args.P <= 0) return NWX_OK;
  args.gconsts = net.gconsts;
  return launch_variant<true, true, 4, false, true, true>(net, args, st);
}

}  // namespace nwx
