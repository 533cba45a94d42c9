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
// Inference folds _feature_linear into the views layer (no non-linearity between them, nerf_model.py:64-68):
// W_fold = W_view[:, :256] . W_feature (128 x 256), kept as 4 more half-size K-blocks after the plain image.
constexpr int kFoldKBlock0 = kNumKBlocks;                    // global K-block index of the folded views layer
constexpr size_t kWeightImageBytes = (size_t)kFullKBlocks * kKBlockBytes + 8 * (kKBlockBytes / 2);

__host__ __device__ constexpr uint32_t kblock_offset(int g) {
  return g < kFullKBlocks ? (uint32_t)g * kKBlockBytes
                          : (uint32_t)kFullKBlocks * kKBlockBytes + (uint32_t)(g - kFullKBlocks) * (kKBlockBytes / 2);
}

// fp32 side data of one network, passed to the kernel by value (constant bank).
struct alignas(16) MlpConsts {     // 16-byte aligned: the epilogues read it with 128-bit uniform constant loads
  float bias[9][kHidden];            // _pts_linears.{0..7}.bias, _feature_linear.bias
  float w_alpha[kHidden];            // _alpha_linear.weight
  float w_rgb[3][kViewHidden];       // _rgb_linear.weight
  float b_alpha;
  float b_rgb[3];
};

// ---- training: activation / gradient tile images in HBM -------------------------------------
// The training forward saves every tensor-core operand as the very [128 points x 64 features]
// bf16 swizzled tile images it builds in shared memory (16 KB each), slot-major:
//   address(slot, tile, kb) = base + ((slot_kb0(slot) * n_tiles + tile * slot_nkb(slot) + kb) << 14)
// act slots: 0 = PE (1 K-block), 1..8 = h1..h8 (post-ReLU); the feature output (old slot 9) is not
// materialised since the views layer is folded.
// grad slots (written by the dX kernel): 0 = G_views (2 K-blocks), 1 = unused (was d_feature), 2..9 = G8..G1
// (G_l = dL/d(pre-activation of the layer that produced h_l)).
constexpr int kTileImgBytes = 16384;
__host__ __device__ constexpr int act_slot_nkb(int s) { return s == 0 ? 1 : 4; }
__host__ __device__ constexpr int act_slot_kb0(int s) { return s == 0 ? 0 : 1 + 4 * (s - 1); }
constexpr int kActKBlocksPerTile = 33;      // slots 0..8 (slot 9, the feature output, is not materialised: folded)
__host__ __device__ constexpr int grad_slot_nkb(int s) { return s == 0 ? 2 : 4; }
__host__ __device__ constexpr int grad_slot_kb0(int s) { return s == 0 ? 0 : 2 + 4 * (s - 1); }
constexpr int kGradKBlocksPerTile = 38;
__host__ __device__ inline size_t tile_img_offset(int kb0, int nkb, int64_t n_tiles, int64_t tile, int kb) {
  return ((size_t)kb0 * n_tiles + (size_t)tile * nkb + kb) * kTileImgBytes;
}

// ReLU' bit masks for the dX kernel: per tile, per pts layer l = 0..7 (the layer that produced h_{l+1}), per
// 32-column chunk c = 0..7, one uint32 per row: [tile][l][c][128 rows]; 32 KB per tile (256 B per point instead
// of re-reading the 4 KB of activations).  One extra tile at the end takes the writes of tail tiles.
constexpr size_t kMaskTileBytes = 8 * 8 * 128 * 4;
__host__ __device__ inline size_t mask_img_offset(int64_t tile, int l) { return (size_t)tile * kMaskTileBytes + (size_t)l * 4096; }

struct PackedNet {
  uint8_t* wimg = nullptr;           // device: swizzled bf16 K-block images (kWeightImageBytes)
  float* wdir_t = nullptr;           // device: [27][128] fp32 = _views_linears.0.weight[:, 256:]^T
  float* bview = nullptr;            // device: [128] fp32 _views_linears.0.bias
  float* bview_fold = nullptr;       // device: [128] fp32 b_view + W_view[:, :256] . b_feature (folded inference path)
  MlpConsts consts;                  // host copy
  MlpConsts* gconsts = nullptr;      // device copy (training kernels read it through a pointer)
  uint8_t* wimg_t = nullptr;         // device: transposed-weight K-block images for the dX kernel
  const float* master = nullptr;     // device: flat fp32 master parameters given to the last train_pack (caller-owned)
  bool loaded = false;
  bool consts_stale = false;         // host consts older than the device master weights (training)
};

// Timing experiments (tools/train_experiments.py; results wrong on purpose) exist only in builds made with
// `make EXPERIMENTS=1`; in the product build the switch is the constant 0 and the branches disappear.
#ifdef NWX_EXPERIMENTS
#define NWX_EXP(args) ((args).experiment)
#else
#define NWX_EXP(args) 0
#endif

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
  const MlpConsts* gconsts;          // training: biases / heads read from device memory
  uint8_t* acts;                     // training: activation tile images (see act slots), or nullptr
  float* hv_out;                     // training: views-layer hidden [P,128] fp32 (post-ReLU), or nullptr
  uint32_t* masks;                   // training: ReLU' bit masks (mask_img_offset), or nullptr
  int64_t P;                         // total points
  int64_t n_tiles;                   // ceil(P / 128)
  int ray_dim, S, iters, dbg_layer, which;
  int experiment;                    // training forward timing experiments (nwx_debug_experiment), 0 = none
};

int pack_network(PackedNet& net, const float* const* tensors, cudaStream_t st);
int pack_network_images(PackedNet& net, const float* const* tensors, bool with_fold, cudaStream_t st,
                        uint8_t* fold_t = nullptr);
inline bool variant_folds(int variant) { return variant <= 1; }
int launch_dirbias(const PackedNet& net, const float* dirs, int stride, int64_t n, bool pre_embedded, bool fold,
                   float* out, cudaStream_t st);
int launch_mlp(const PackedNet& net, MlpArgs args, int variant, cudaStream_t st);
int launch_mlp_train_forward(const PackedNet& net, MlpArgs args, cudaStream_t st);

// ---- training (train.cu) ----
struct TrainBwdArgs {
  const float* d_raw;        // [P,4]
  const float* hv;           // [P,128]
  const uint8_t* acts;       // activation images of this network's forward
  const uint32_t* masks;     // ReLU' bit masks of this network's forward (mask_image_bytes)
  uint8_t* gimg;             // gradient image scratch (grad_image_bytes)
  float* partial;            // [max_partials][NWX_PARAMS_PER_NET] dW partials (zero-initialised once)
  float* head_partial;       // head_partial_bytes(n_tiles): per-block sums of the head gradients
  float* fold_scratch;       // kFoldScratchFloats: d W_fold [128 x 256] and d b_fold [128]
  const float* pe_dir;       // [n_rays,27] embedded view directions
  float* grad;               // flat gradient buffer of this network (accumulated into)
  uint32_t* diag;
  int64_t P;
  int S, max_partials, which;
  int experiment;            // dX timing experiments (nwx_debug_experiment), 0 = none
};
int upload_train_consts(int which, const MlpConsts* dev_src, cudaStream_t st);       // backward kernels (train.cu)
int upload_fwd_train_consts(int which, const MlpConsts* dev_src, cudaStream_t st);   // training forward (mlp.cu)
size_t act_image_bytes(int64_t n_tiles);
size_t grad_image_bytes(int64_t n_tiles);
size_t head_partial_bytes(int64_t n_tiles, int64_t n_rays, int S);
int dw_partial_rows();
constexpr size_t kFoldScratchFloats = (size_t)kViewHidden * kHidden + kViewHidden;
inline size_t mask_image_bytes(int64_t n_tiles) { return (size_t)(n_tiles + 1) * kMaskTileBytes; }
const int* flat_offsets();
size_t packed_transposed_bytes();
int train_pack(PackedNet& net, const float* params_flat, cudaStream_t st);
int launch_mlp_backward(const PackedNet& net, const TrainBwdArgs& a, cudaStream_t st, cudaStream_t heads_st = nullptr,
                        cudaEvent_t fork = nullptr, cudaEvent_t join = nullptr);
constexpr int kMseMaxBlocks = 512;
constexpr size_t kMseScratchBytes = (1 + 2 * kMseMaxBlocks) * sizeof(double);   // ticket word + per-block partial sums
int launch_mse_grad(const float* rgb_c, const float* rgb_f, const float* gt, int64_t n_rays, float* d_c, float* d_f,
                    double* loss_scratch, double* loss_out, cudaStream_t st);
// folded views layer of two networks in one launch (mlp.cu)
struct PackFoldNet {
  const float *wv, *wf, *bv, *bf;    // fp32 master: W_view [128][283], W_feature [256][256], b_view, b_feature
  uint8_t* wimg; float* bview_fold; uint8_t* wimg_t;
};
struct PackFoldPair { PackFoldNet net[2]; };
void launch_pack_fold_pair(const PackFoldPair& f, cudaStream_t st);
int launch_adam_pack(PackedNet (&nets)[2], float* params, const float* grads, float* m, float* v, float lr, float b1,
                     float b2, float eps, int step, float grad_scale, cudaStream_t st);
int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                int step, float grad_scale, cudaStream_t st);

}  // namespace nwx
