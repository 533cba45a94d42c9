// Shared declarations for the fused PE + MLP kernel (mlp.cu) and the context (context.cu).
#pragma once
#include "nwx_common.cuh"

namespace nwx {

// Architecture of NeRFModel(D=8, W=256, input_ch=63, input_ch_views=27, skips=(4,),
// use_view_dirs=True) -- reference nerf/models/nerf_model.py:12-43.
constexpr int kHidden = 256;
constexpr int kPeXyz = 63;        // 3 + 3*2*10  (embedding.py:24-38, num_freqs_3d = 10)
constexpr int kPeDir = 27;        // 3 + 3*2*4   (num_freqs_2d = 4)
constexpr int kViewHidden = 128;
constexpr int kNumLayers = 10;    // tensor-core layers: pts 0..7, feature (8), views (9)

// K-blocks (64 input columns each) per tensor-core layer.  Layer 5 = [pe(63->64) | h(256)].
__host__ __device__ constexpr int layer_kblocks(int l) { return l == 0 ? 1 : (l == 5 ? 5 : 4); }
constexpr int kNumKBlocks = 1 + 4 * 4 + 5 + 4 * 4;           // 38
constexpr int kFullKBlocks = kNumKBlocks - 4;                // 34 with N = 256, then 4 with N = 128
constexpr int kKBlockBytes = 256 * 64 * 2;                   // one [256 x 64] bf16 image
constexpr size_t kWeightImageBytes = (size_t)kFullKBlocks * kKBlockBytes + 4 * (kKBlockBytes / 2);

__host__ __device__ constexpr uint32_t kblock_offset(int g) {
  return g < kFullKBlocks ? (uint32_t)g * kKBlockBytes
                          : (uint32_t)kFullKBlocks * kKBlockBytes + (uint32_t)(g - kFullKBlocks) * (kKBlockBytes / 2);
}

// fp32 side data of one network, passed to the kernel by value (constant bank).
struct MlpConsts {
  float bias[9][kHidden];            // _pts_linears.{0..7}.bias, _feature_linear.bias
  float w_alpha[kHidden];            // _alpha_linear.weight
  float w_rgb[3][kViewHidden];       // _rgb_linear.weight
  float b_alpha;
  float b_rgb[3];
};

struct PackedNet {
  uint8_t* wimg = nullptr;           // device: swizzled bf16 K-block images (kWeightImageBytes)
  float* wdir_t = nullptr;           // device: [27][128] fp32 = _views_linears.0.weight[:, 256:]^T
  float* bview = nullptr;            // device: [128] fp32 _views_linears.0.bias
  MlpConsts consts;                  // host copy
  bool loaded = false;
};

struct MlpArgs {
  const float* rays;                 // [N, ray_dim] or nullptr (points mode)
  const float* z;                    // [N, S]
  const float* pts;                  // [P, 3] (points mode, S == 1)
  const float* embedded;             // [P, 90] caller-embedded input (NeRFModel.forward signature, S == 1)
  const uint8_t* wimg;
  const float* dirbias;              // [N (or P), 128]: b_view + W_view[:, 256:] . pe(viewdir)
  float* raw_out;                    // [P, 4]
  float* dbg_out;                    // optional tap [P, 256] (post-activation fp32 of dbg_layer)
  uint32_t* diag;                    // optional host-mapped diagnostics word(s)
  int64_t P;                         // total points
  int ray_dim, S, iters, dbg_layer;
};

int pack_network(PackedNet& net, const float* const* tensors, cudaStream_t st);
int launch_dirbias(const PackedNet& net, const float* dirs, int stride, int64_t n, bool pre_embedded, float* out,
                   cudaStream_t st);
int launch_mlp(const PackedNet& net, MlpArgs args, int variant, cudaStream_t st);

}  // namespace nwx
