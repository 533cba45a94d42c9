// Context (packed weights + scratch) and the fused launch sequence of _volumetric_rendering
// (reference nerf/inference/nerf_replica_inference_handler.py:203-277 and
// nerf/training/nerf_replica_training_handler.py:534-618).
//
// The reference walks a frame in 38 Python ray chunks x 64 network chunks, with 22 host syncs
// per chunk.  Here one chunk of any size is 8 kernel launches on one
// stream, no host synchronisation, nothing per-point materialised except raw [N,S,4]:
//   coarse_z -> dirbias(coarse) -> MLP(coarse) -> composite -> sample_pdf+merge
//            -> dirbias(fine)   -> MLP(fine)   -> composite (writes rgb and/or the uint8 pixels)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "mlp.cuh"

std::atomic<int64_t> g_nwx_launches{0};

struct nwx_ctx {
  int device = 0;
  nwx::PackedNet net[2];
  int mlp_variant = 0;
  // scratch, grown on demand (nwx_ctx_reserve pre-sizes it)
  float* scratch = nullptr;
  size_t scratch_floats = 0;
  int64_t scratch_generation = 0;   // bumped whenever the scratch is re-allocated (captured CUDA graphs hold its address)
  float* dbg_out = nullptr;      // optional tap target set by nwx_debug_tap
  int dbg_layer = -1;
  int experiment = 0;            // timing experiments in the training kernels (results wrong on purpose), 0 = none
  uint32_t* diag = nullptr;      // host-mapped diagnostics the kernels write before aborting (own_diag unless overridden)
  uint32_t* own_diag = nullptr;  // 4 words of mapped pinned host memory, allocated with the context
  // optional per-stage device timing of nwx_render_rays (bench.py: roofline of the dominant kernel)
  // training scratch (activation / gradient tile images, dW partials, ...), grown on demand
  uint8_t* tscratch = nullptr;
  size_t tscratch_bytes = 0;
  float* partial = nullptr;      // [n_partials][NWX_PARAMS_PER_NET], zero-initialised once
  double* loss_scratch = nullptr;   // ticket + per-block partial sums of the MSE kernel, zero-initialised once
  int n_partials = 0;
  // training: the head kernels of each network's backward run on this stream, underneath its dW kernel
  cudaStream_t heads_stream = nullptr;
  cudaEvent_t ev_heads_fork[2] = {}, ev_heads_join[2] = {};
  bool heads_on_side_stream = true;          // NWX_TRAIN_HEADS_STREAM=0: on the caller's stream (A/B)
  bool profiling = false;
  bool ev_recorded = false;
  cudaEvent_t ev[NWX_NUM_STAGES + 1] = {};
};

namespace {

// The random source of one draw: the caller's tensor if given, else the library's counter-based
// generator when the options ask for it (rng_* fields), else none.
nwx::RngSpec rng_for(const nwx_render_opts* o, uint32_t stream_id, const float* tensor, bool wanted) {
  nwx::RngSpec r;
  r.seed = o->rng_seed; r.offset = o->rng_offset; r.stream = stream_id;
  r.scale = stream_id >= 2 ? o->raw_noise_std : 1.0f;
  r.on = (tensor == nullptr && wanted) ? 1 : 0;
  return r;
}

struct ScratchPlan {
  size_t z_c, raw_c, w_c, z_s, z_f, raw_f, dirbias, rgb_c, rgb_f, total;
};

ScratchPlan plan_scratch(int64_t N, int Sc, int Ni) {
  ScratchPlan p{};
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 63) & ~(size_t)63; return o; };   // 256 B aligned
  const int Sf = Sc + Ni;
  p.z_c = take((size_t)N * Sc);
  p.raw_c = take((size_t)N * Sc * 4);
  p.w_c = take((size_t)N * Sc);
  p.z_s = take((size_t)N * Ni);
  p.z_f = take((size_t)N * Sf);
  p.raw_f = take((size_t)N * Sf * 4);
  p.dirbias = take((size_t)N * nwx::kViewHidden);
  p.rgb_c = take((size_t)N * 3);
  p.rgb_f = take((size_t)N * 3);
  p.total = off;
  return p;
}

int ensure_scratch(nwx_ctx* ctx, size_t floats) {
  if (floats <= ctx->scratch_floats) return NWX_OK;
  if (ctx->scratch) NWX_CUDA_TRY(cudaFree(ctx->scratch));
  ctx->scratch = nullptr;
  ctx->scratch_floats = 0;
  NWX_CUDA_TRY(cudaMalloc(&ctx->scratch, floats * sizeof(float)));
  ctx->scratch_floats = floats;
  ++ctx->scratch_generation;
  return NWX_OK;
}

}  // namespace

extern "C" int nwx_version(void) { return NWX_VERSION; }

extern "C" const char* nwx_error_string(int code) {
  switch (code) {
    case NWX_OK: return "ok";
    case NWX_E_INVALID: return "invalid argument";
    case NWX_E_NO_WEIGHTS: return "weights not loaded (call nwx_load_weights)";
    case NWX_E_UNSUPPORTED: return "device is not sm_100 (B200); there is no fallback path";
    case NWX_E_STALE: return "weights were trained since the last nwx_load_weights: reload them before inference";
    default:
      if (code >= NWX_E_CUDA) return cudaGetErrorString((cudaError_t)(code - NWX_E_CUDA));
      return "unknown error";
  }
}

extern "C" int64_t nwx_launch_count(void) { return g_nwx_launches.load(std::memory_order_relaxed); }

extern "C" int nwx_ctx_create(int device, nwx_ctx** out) {
  NWX_REQUIRE(out);
  *out = nullptr;
  NWX_CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  NWX_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return NWX_E_UNSUPPORTED;     // tcgen05/TMEM only; the product has no fallback
  nwx_ctx* c = new (std::nothrow) nwx_ctx();
  if (!c) return NWX_E_INVALID;
  c->device = device;
  if (const char* v = getenv("NWX_TRAIN_HEADS_STREAM")) c->heads_on_side_stream = atoi(v) != 0;
  if (const char* v = getenv("NWX_MLP_VARIANT")) {       // A/B measurements of the MLP kernel variants (nwx_set_mlp_variant)
    const int iv = atoi(v);
    if (iv >= 0 && iv <= 4) c->mlp_variant = iv;
  }
  // always-on diagnostics: survives a poisoned context because it lives in mapped host memory
  if (cudaHostAlloc(reinterpret_cast<void**>(&c->own_diag), 4 * sizeof(uint32_t), cudaHostAllocMapped) == cudaSuccess) {
    memset(c->own_diag, 0, 4 * sizeof(uint32_t));
    c->diag = c->own_diag;
  } else {
    (void)cudaGetLastError();
    c->own_diag = nullptr;
  }
  *out = c;
  return NWX_OK;
}

extern "C" int nwx_ctx_destroy(nwx_ctx* ctx) {
  if (!ctx) return NWX_OK;
  cudaSetDevice(ctx->device);
  for (auto& n : ctx->net) {
    if (n.wimg) cudaFree(n.wimg);
    if (n.wdir_t) cudaFree(n.wdir_t);
    if (n.bview) cudaFree(n.bview);
    if (n.bview_fold) cudaFree(n.bview_fold);
  }
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->tscratch) cudaFree(ctx->tscratch);
  if (ctx->partial) cudaFree(ctx->partial);
  if (ctx->loss_scratch) cudaFree(ctx->loss_scratch);
  for (auto& n : ctx->net) {
    if (n.wimg_t) cudaFree(n.wimg_t);
    if (n.gconsts) cudaFree(n.gconsts);
  }
  for (auto e : ctx->ev)
    if (e) cudaEventDestroy(e);
  if (ctx->heads_stream) cudaStreamDestroy(ctx->heads_stream);
  for (int w = 0; w < 2; ++w) {
    if (ctx->ev_heads_fork[w]) cudaEventDestroy(ctx->ev_heads_fork[w]);
    if (ctx->ev_heads_join[w]) cudaEventDestroy(ctx->ev_heads_join[w]);
  }
  if (ctx->own_diag) cudaFreeHost(ctx->own_diag);
  delete ctx;
  return NWX_OK;
}

extern "C" int nwx_load_weights(nwx_ctx* ctx, int which, const float* const* tensors, void* stream) {
  NWX_REQUIRE(ctx && tensors && (which == NWX_NET_COARSE || which == NWX_NET_FINE));
  for (int i = 0; i < NWX_NUM_WEIGHT_TENSORS; ++i) NWX_REQUIRE(tensors[i] != nullptr);
  return nwx::pack_network(ctx->net[which], tensors, (cudaStream_t)stream);
}

extern "C" int nwx_set_mlp_variant(nwx_ctx* ctx, int variant) {
  NWX_REQUIRE(ctx && variant >= 0 && variant <= 4);
  ctx->mlp_variant = variant;
  return NWX_OK;
}

// Test/diagnostic hooks (declared here, not part of the reference-facing surface in nwx.h).
extern "C" int nwx_debug_tap(nwx_ctx* ctx, int layer, float* out) {
  NWX_REQUIRE(ctx);
  ctx->dbg_layer = layer;
  ctx->dbg_out = out;
  return NWX_OK;
}
extern "C" int nwx_debug_diag(nwx_ctx* ctx, uint32_t* host_mapped) {
  NWX_REQUIRE(ctx);
  ctx->diag = host_mapped ? host_mapped : ctx->own_diag;
  return NWX_OK;
}
// Timing experiments in the training forward / dX kernels (tools/train_experiments.py); results are WRONG on purpose:
// 11 = the epilogues do not wait for the previous TMA store of their tile, 12 = no TMA stores of the tile images at all,
// 13 = the forward neither builds nor stores the ReLU' bit masks, 14 = no named barriers around the tile writes (and 11),
// 15 = the forward does not store the views hidden, 18 = the training forward saves nothing at all (the kTrain
// instantiation with null save pointers: what the operands' way to HBM costs in total).
// Only a library built with `make EXPERIMENTS=1` contains them; the product build accepts code 0 alone.
extern "C" int nwx_debug_experiment(nwx_ctx* ctx, int code) {
#ifdef NWX_EXPERIMENTS
  NWX_REQUIRE(ctx && (code == 0 || (code >= 11 && code <= 15) || code == 18));
#else
  NWX_REQUIRE(ctx && code == 0);
#endif
  ctx->experiment = code;
  return NWX_OK;
}
extern "C" int nwx_ctx_last_diag(nwx_ctx* ctx, uint32_t* out4) {
  NWX_REQUIRE(ctx && out4);
  for (int i = 0; i < 4; ++i) out4[i] = ctx->diag ? ((volatile uint32_t*)ctx->diag)[i] : 0u;
  return NWX_OK;
}

extern "C" int nwx_ctx_set_profiling(nwx_ctx* ctx, int on) {
  NWX_REQUIRE(ctx);
  if (on && !ctx->ev[0])
    for (auto& e : ctx->ev) NWX_CUDA_TRY(cudaEventCreate(&e));
  ctx->profiling = on != 0;
  ctx->ev_recorded = false;
  return NWX_OK;
}

extern "C" int nwx_ctx_stage_ms(nwx_ctx* ctx, float* ms_out) {
  NWX_REQUIRE(ctx && ms_out && ctx->ev_recorded);
  NWX_CUDA_TRY(cudaEventSynchronize(ctx->ev[NWX_NUM_STAGES]));
  for (int i = 0; i < NWX_NUM_STAGES; ++i) NWX_CUDA_TRY(cudaEventElapsedTime(&ms_out[i], ctx->ev[i], ctx->ev[i + 1]));
  return NWX_OK;
}

extern "C" int nwx_ctx_scratch_state(nwx_ctx* ctx, int64_t* bytes, int64_t* generation) {
  NWX_REQUIRE(ctx && bytes && generation);
  *bytes = (int64_t)(ctx->scratch_floats * sizeof(float));
  *generation = ctx->scratch_generation;
  return NWX_OK;
}

extern "C" int nwx_ctx_reserve(nwx_ctx* ctx, int64_t max_rays, int n_samples, int n_importance) {
  NWX_REQUIRE(ctx && max_rays > 0 && n_samples > 0 && n_importance >= 0);
  return ensure_scratch(ctx, plan_scratch(max_rays, n_samples, n_importance).total);
}

static int run_mlp(nwx_ctx* ctx, int which, const float* rays, int ray_dim, const float* z, const float* pts,
                   const float* dirs, int dir_stride, int64_t n_dir, int64_t P, int S, float* dirbias,
                   float* raw_out, cudaStream_t st, const float* embedded = nullptr, cudaEvent_t mid = nullptr) {
  const nwx::PackedNet& net = ctx->net[which];
  if (!net.loaded) return NWX_E_NO_WEIGHTS;
  if (net.consts_stale) return NWX_E_STALE;      // trained since the last nwx_load_weights: biases on the host are old
  int rc = nwx::launch_dirbias(net, dirs, dir_stride, n_dir, embedded != nullptr, nwx::variant_folds(ctx->mlp_variant),
                               dirbias, st);
  if (rc) return rc;
  if (mid) NWX_CUDA_TRY(cudaEventRecord(mid, st));
  nwx::MlpArgs a{};
  a.rays = rays; a.z = z; a.pts = pts; a.embedded = embedded; a.wimg = net.wimg; a.dirbias = dirbias; a.raw_out = raw_out;
  a.dbg_out = ctx->dbg_out; a.dbg_layer = ctx->dbg_layer; a.diag = ctx->diag;
  a.P = P; a.ray_dim = ray_dim; a.S = S;
  return nwx::launch_mlp(net, a, ctx->mlp_variant, st);
}

extern "C" int nwx_mlp_forward(nwx_ctx* ctx, int which, const float* rays, int ray_dim, const float* z,
                               int64_t N, int S, float* raw_out, void* stream) {
  NWX_REQUIRE(ctx && (which == 0 || which == 1) && ray_dim >= NWX_RAY_DIM && S >= 1 && N >= 0);
  if (N == 0) return NWX_OK;
  NWX_REQUIRE(rays && z && raw_out);
  int rc = ensure_scratch(ctx, (size_t)N * nwx::kViewHidden);
  if (rc) return rc;
  return run_mlp(ctx, which, rays, ray_dim, z, nullptr, rays + 8, ray_dim, N, N * S, S, ctx->scratch, raw_out,
                 (cudaStream_t)stream);
}

extern "C" int nwx_mlp_forward_points(nwx_ctx* ctx, int which, const float* pts, const float* dirs, int64_t P,
                                      int pts_per_dir, float* raw_out, void* stream) {
  NWX_REQUIRE(ctx && (which == 0 || which == 1) && P >= 0 && pts_per_dir >= 1);
  if (P == 0) return NWX_OK;
  NWX_REQUIRE(pts && dirs && raw_out);
  const int64_t n_dir = (P + pts_per_dir - 1) / pts_per_dir;
  int rc = ensure_scratch(ctx, (size_t)n_dir * nwx::kViewHidden);
  if (rc) return rc;
  return run_mlp(ctx, which, nullptr, 0, nullptr, pts, dirs, 3, n_dir, P, pts_per_dir, ctx->scratch, raw_out,
                 (cudaStream_t)stream);
}

extern "C" int nwx_mlp_forward_embedded(nwx_ctx* ctx, int which, const float* x, int64_t P, float* raw_out,
                                        void* stream) {
  NWX_REQUIRE(ctx && (which == 0 || which == 1) && P >= 0);
  if (P == 0) return NWX_OK;
  NWX_REQUIRE(x && raw_out);
  int rc = ensure_scratch(ctx, (size_t)P * nwx::kViewHidden);
  if (rc) return rc;
  return run_mlp(ctx, which, nullptr, 0, nullptr, nullptr, x + nwx::kPeXyz, nwx::kPeXyz + nwx::kPeDir, P, P, 1,
                 ctx->scratch, raw_out, (cudaStream_t)stream, x);
}

extern "C" int nwx_render_rays(nwx_ctx* ctx, const float* rays, int64_t N, const nwx_render_opts* o,
                               const nwx_render_out* out, void* stream) {
  NWX_REQUIRE(ctx && o && out && N >= 0);
  if (N == 0) return NWX_OK;
  NWX_REQUIRE(rays && (out->rgb_fine || out->rgb8_fine));
  NWX_REQUIRE(o->n_samples >= 11 && o->n_samples <= 128 && o->n_importance >= 1 && o->n_importance <= 128);
  NWX_REQUIRE(o->ray_dim >= NWX_RAY_DIM && o->t_vals && (o->u || o->u_lin || o->rng_u));
  if (N == 0) return NWX_OK;
  if (!ctx->net[0].loaded || !ctx->net[1].loaded) return NWX_E_NO_WEIGHTS;
  if (ctx->net[0].consts_stale || ctx->net[1].consts_stale) return NWX_E_STALE;
  auto st = (cudaStream_t)stream;
  const int Sc = o->n_samples, Ni = o->n_importance, Sf = Sc + Ni, rd = o->ray_dim;
  const ScratchPlan pl = plan_scratch(N, Sc, Ni);
  int rc = ensure_scratch(ctx, pl.total);
  if (rc) return rc;
  float* s = ctx->scratch;
  // caller-provided outputs double as the working buffers where they exist
  float* z_c = out->z_vals_coarse ? out->z_vals_coarse : s + pl.z_c;
  float* raw_c = out->raw_coarse ? out->raw_coarse : s + pl.raw_c;
  float* w_c = out->weights_coarse ? out->weights_coarse : s + pl.w_c;
  float* z_s = out->z_samples;                       // NULL = not wanted: the resampling kernel then skips the store
  float* z_f = out->z_vals_fine ? out->z_vals_fine : s + pl.z_f;
  float* raw_f = out->raw_fine ? out->raw_fine : s + pl.raw_f;
  float* rgb_c = out->rgb_coarse ? out->rgb_coarse : s + pl.rgb_c;
  float* dirb = s + pl.dirbias;
  if (out->flags) NWX_CUDA_TRY(cudaMemsetAsync(out->flags, 0, sizeof(int32_t), st));
  const bool prof = ctx->profiling;
  auto mark = [&](int i) -> int {
    if (prof) NWX_CUDA_TRY(cudaEventRecord(ctx->ev[i], st));
    return NWX_OK;
  };

  if ((rc = mark(0))) return rc;
  const nwx::RngSpec rj = rng_for(o, 0, o->t_rand, o->rng_jitter != 0), ru = rng_for(o, 1, o->u, o->rng_u != 0);
  const nwx::RngSpec rnc = rng_for(o, 2, o->noise_coarse, o->raw_noise_std > 0.f);
  const nwx::RngSpec rnf = rng_for(o, 3, o->noise_fine, o->raw_noise_std > 0.f);
  if ((rc = nwx::launch_coarse_z(rays, rd, N, Sc, o->t_vals, o->t_rand, rj, z_c, st))) return rc;
  if ((rc = mark(1))) return rc;
  if ((rc = run_mlp(ctx, NWX_NET_COARSE, rays, rd, z_c, nullptr, rays + 8, rd, N, N * Sc, Sc, dirb, raw_c, st, nullptr,
                    prof ? ctx->ev[2] : nullptr))) return rc;
  if ((rc = mark(3))) return rc;
  if ((rc = nwx::launch_composite_fwd(raw_c, z_c, rays + 3, rd, o->noise_coarse, rnc, N, Sc, o->white_bkgd, rgb_c,
                                      out->disp_coarse, out->acc_coarse, out->depth_coarse, w_c, out->flags, st))) return rc;
  if ((rc = mark(4))) return rc;
  if ((rc = nwx::launch_sample_pdf(z_c, w_c, Sc, o->u, ru, o->u_lin, Ni, N, z_s, z_f, out->inds, out->z_std, st,
                                   s + pl.z_s))) return rc;
  if ((rc = mark(5))) return rc;
  if ((rc = run_mlp(ctx, NWX_NET_FINE, rays, rd, z_f, nullptr, rays + 8, rd, N, N * Sf, Sf, dirb, raw_f, st, nullptr,
                    prof ? ctx->ev[6] : nullptr))) return rc;
  if ((rc = mark(7))) return rc;
  if ((rc = nwx::launch_composite_fwd(raw_f, z_f, rays + 3, rd, o->noise_fine, rnf, N, Sf, o->white_bkgd, out->rgb_fine,
                                      out->disp_fine, out->acc_fine, out->depth_fine, out->weights_fine, out->flags, st,
                                      out->rgb8_fine))) return rc;
  if ((rc = mark(8))) return rc;
  ctx->ev_recorded = prof;
  return NWX_OK;
}

// ------------------------------------------------------------------------------------------------
// training: forward + backward of one ray batch (the autograd part of
// NeRFReplicaTrainingHandler.step, training handler:277-308), optimiser step, re-pack
// ------------------------------------------------------------------------------------------------
extern "C" int nwx_param_offsets(int* offsets24) {
  NWX_REQUIRE(offsets24);
  for (int i = 0; i < NWX_NUM_WEIGHT_TENSORS; ++i) offsets24[i] = nwx::flat_offsets()[i];
  return NWX_OK;
}

extern "C" int nwx_train_pack(nwx_ctx* ctx, int which, const float* params_flat, void* stream) {
  NWX_REQUIRE(ctx && params_flat && (which == 0 || which == 1));
  return nwx::train_pack(ctx->net[which], params_flat, (cudaStream_t)stream);
}

extern "C" int nwx_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1,
                             float beta2, float eps, int step, float grad_scale, void* stream) {
  NWX_REQUIRE(params && grads && m && v && n >= 0 && step >= 1);
  if (n == 0) return NWX_OK;
  return nwx::launch_adam(params, grads, m, v, n, lr, beta1, beta2, eps, step, grad_scale, (cudaStream_t)stream);
}

extern "C" int nwx_adam_pack_step(nwx_ctx* ctx, float* params, const float* grads, float* m, float* v, float lr,
                                  float beta1, float beta2, float eps, int step, float grad_scale, void* stream) {
  NWX_REQUIRE(ctx && params && grads && m && v && step >= 1);
  return nwx::launch_adam_pack(ctx->net, params, grads, m, v, lr, beta1, beta2, eps, step, grad_scale, (cudaStream_t)stream);
}

// Test hook: copy one packed buffer of a network to `dst` (device): 0 forward weight image, 1 transposed image (dX),
// 2 device-side constants (biases / heads), 3 view-direction table, 4 views bias, 5 folded views bias.
extern "C" int nwx_debug_copy_packed(nwx_ctx* ctx, int which, int what, void* dst, int64_t bytes, void* stream) {
  NWX_REQUIRE(ctx && (which == 0 || which == 1) && dst && bytes >= 0);
  const nwx::PackedNet& n = ctx->net[which];
  const void* src = nullptr;
  int64_t size = 0;
  switch (what) {
    case 0: src = n.wimg; size = (int64_t)nwx::kWeightImageBytes; break;
    case 1: src = n.wimg_t; size = (int64_t)nwx::packed_transposed_bytes(); break;
    case 2: src = n.gconsts; size = (int64_t)sizeof(nwx::MlpConsts); break;
    case 3: src = n.wdir_t; size = (int64_t)sizeof(float) * nwx::kPeDir * nwx::kViewHidden; break;
    case 4: src = n.bview; size = (int64_t)sizeof(float) * nwx::kViewHidden; break;
    case 5: src = n.bview_fold; size = (int64_t)sizeof(float) * nwx::kViewHidden; break;
    default: return NWX_E_INVALID;
  }
  if (!src) return NWX_E_NO_WEIGHTS;
  NWX_REQUIRE(bytes == size);
  NWX_CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)size, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return NWX_OK;
}

namespace {
// The training kernels read biases / heads from process-global __constant__ banks (mlp.cu c_fwd_train_consts,
// train.cu c_train_consts) that every nwx_train_fwd_bwd refreshes on its own stream.  Two contexts (or one
// context on two streams) training on the same device would race on them, so consecutive training calls on
// a device are chained: a call that is not on the previous call's (ctx, stream) first waits -- on the device,
// no host synchronisation -- for the event the previous call recorded behind its last kernel.  The host-side
// mutex keeps two host threads from interleaving their enqueues.
struct TrainGate {
  std::mutex mu;
  cudaEvent_t done = nullptr;
  const nwx_ctx* owner = nullptr;
  cudaStream_t stream = nullptr;
};
TrainGate g_train_gate[nwx::kMaxDevices];

struct TrainPlan {
  size_t acts_c, acts_f, masks_c, masks_f, gimg, head_partial, fold, hv_c, hv_f, d_raw_c, d_raw_f, pe_dir, d_rgb_c, d_rgb_f, rgb_c, rgb_f, total;
};
TrainPlan plan_train(int64_t N, int Sc, int Ni) {
  TrainPlan p{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 1023) & ~(size_t)1023; return o; };
  const int Sf = Sc + Ni;
  const int64_t tc = (N * Sc + 127) / 128, tf = (N * Sf + 127) / 128;
  p.acts_c = take(nwx::act_image_bytes(tc));
  p.acts_f = take(nwx::act_image_bytes(tf));
  p.masks_c = take(nwx::mask_image_bytes(tc));
  p.masks_f = take(nwx::mask_image_bytes(tf));
  p.gimg = take(nwx::grad_image_bytes(tf));            // reused: coarse backward, then fine backward
  p.head_partial = take(nwx::head_partial_bytes(tf, N, Sf));  // reused like gimg
  p.fold = take(nwx::kFoldScratchFloats * sizeof(float));
  p.hv_c = take((size_t)N * Sc * nwx::kViewHidden * 4);
  p.hv_f = take((size_t)N * Sf * nwx::kViewHidden * 4);
  p.d_raw_c = take((size_t)N * Sc * 16);
  p.d_raw_f = take((size_t)N * Sf * 16);
  p.pe_dir = take((size_t)N * nwx::kPeDir * 4);
  p.d_rgb_c = take((size_t)N * 12);
  p.d_rgb_f = take((size_t)N * 12);
  p.rgb_c = take((size_t)N * 12);
  p.rgb_f = take((size_t)N * 12);
  p.total = off;
  return p;
}
}  // namespace

extern "C" int nwx_train_fwd_bwd(nwx_ctx* ctx, const nwx_train_io* io, int64_t N, const nwx_render_opts* o, void* stream) {
  NWX_REQUIRE(ctx && io && o && io->rays && io->gt_rgb && io->grad_coarse && io->grad_fine && io->loss && N > 0);
  NWX_REQUIRE(o->n_samples >= 11 && o->n_samples <= 128 && o->n_importance >= 1 && o->n_importance <= 128);
  NWX_REQUIRE(o->ray_dim >= NWX_RAY_DIM && o->t_vals && (o->u || o->u_lin || o->rng_u));
  for (int w = 0; w < 2; ++w)
    if (!ctx->net[w].loaded || !ctx->net[w].gconsts || !ctx->net[w].wimg_t) return NWX_E_NO_WEIGHTS;   // nwx_train_pack first
  auto st = (cudaStream_t)stream;
  TrainGate& gate = g_train_gate[ctx->device >= 0 && ctx->device < nwx::kMaxDevices ? ctx->device : 0];
  std::lock_guard<std::mutex> gate_lock(gate.mu);
  if (gate.owner && (gate.owner != ctx || gate.stream != st)) NWX_CUDA_TRY(cudaStreamWaitEvent(st, gate.done, 0));
  struct GateRelease {       // record "this call's kernels are done" on every exit path after the wait
    TrainGate& g; const nwx_ctx* c; cudaStream_t s;
    ~GateRelease() {
      if (!g.done && cudaEventCreateWithFlags(&g.done, cudaEventDisableTiming) != cudaSuccess) { g.done = nullptr; return; }
      if (cudaEventRecord(g.done, s) == cudaSuccess) { g.owner = c; g.stream = s; }
    }
  } gate_release{gate, ctx, st};
  const int Sc = o->n_samples, Ni = o->n_importance, Sf = Sc + Ni, rd = o->ray_dim;
  const ScratchPlan pl = plan_scratch(N, Sc, Ni);
  int rc = ensure_scratch(ctx, pl.total);
  if (rc) return rc;
  const TrainPlan tp = plan_train(N, Sc, Ni);
  if (tp.total > ctx->tscratch_bytes) {
    if (ctx->tscratch) NWX_CUDA_TRY(cudaFree(ctx->tscratch));
    ctx->tscratch = nullptr; ctx->tscratch_bytes = 0;
    NWX_CUDA_TRY(cudaMalloc(&ctx->tscratch, tp.total));
    ctx->tscratch_bytes = tp.total;
  }
  if (!ctx->partial) {
    ctx->n_partials = nwx::dw_partial_rows();
    const size_t bytes = (size_t)ctx->n_partials * NWX_PARAMS_PER_NET * sizeof(float);
    NWX_CUDA_TRY(cudaMalloc(&ctx->partial, bytes));
    NWX_CUDA_TRY(cudaMemsetAsync(ctx->partial, 0, bytes, st));
  }
  if (ctx->heads_on_side_stream && !ctx->heads_stream) {
    NWX_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->heads_stream, cudaStreamNonBlocking));
    for (int w = 0; w < 2; ++w) {
      NWX_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_heads_fork[w], cudaEventDisableTiming));
      NWX_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_heads_join[w], cudaEventDisableTiming));
    }
  }
  if (!ctx->loss_scratch) {
    NWX_CUDA_TRY(cudaMalloc(&ctx->loss_scratch, nwx::kMseScratchBytes));
    NWX_CUDA_TRY(cudaMemsetAsync(ctx->loss_scratch, 0, nwx::kMseScratchBytes, st));
  }
  float* s = ctx->scratch;
  uint8_t* ts = ctx->tscratch;
  float *z_c = s + pl.z_c, *raw_c = s + pl.raw_c, *w_c = s + pl.w_c, *z_s = s + pl.z_s, *z_f = s + pl.z_f,
        *raw_f = s + pl.raw_f, *dirb = s + pl.dirbias;
  float* rgb_c = io->rgb_coarse ? io->rgb_coarse : reinterpret_cast<float*>(ts + tp.rgb_c);
  float* rgb_f = io->rgb_fine ? io->rgb_fine : reinterpret_cast<float*>(ts + tp.rgb_f);
  float* hv[2] = {reinterpret_cast<float*>(ts + tp.hv_c), reinterpret_cast<float*>(ts + tp.hv_f)};
  float* d_raw[2] = {reinterpret_cast<float*>(ts + tp.d_raw_c), reinterpret_cast<float*>(ts + tp.d_raw_f)};
  float* d_rgb[2] = {reinterpret_cast<float*>(ts + tp.d_rgb_c), reinterpret_cast<float*>(ts + tp.d_rgb_f)};
  uint8_t* acts[2] = {ts + tp.acts_c, ts + tp.acts_f};
  uint32_t* masks[2] = {reinterpret_cast<uint32_t*>(ts + tp.masks_c), reinterpret_cast<uint32_t*>(ts + tp.masks_f)};
  float* pe_dir = reinterpret_cast<float*>(ts + tp.pe_dir);

  // ---- forward (training handler:534-618), operands saved for the backward ----
  for (int w = 0; w < 2; ++w) {        // this step's biases / heads -> constant bank (stream-ordered D2D)
    if ((rc = nwx::upload_fwd_train_consts(w, ctx->net[w].gconsts, st))) return rc;
    if ((rc = nwx::upload_train_consts(w, ctx->net[w].gconsts, st))) return rc;
  }
  auto fwd = [&](int which, const float* z, int S, float* raw) -> int {
    const nwx::PackedNet& net = ctx->net[which];
    int r = nwx::launch_dirbias(net, io->rays + 8, rd, N, false, true, dirb, st);    // folded views bias
    if (r) return r;
    nwx::MlpArgs a{};
    a.which = which;
    a.rays = io->rays; a.z = z; a.wimg = net.wimg; a.dirbias = dirb; a.raw_out = raw; a.diag = ctx->diag;
    a.acts = acts[which]; a.masks = masks[which]; a.hv_out = hv[which]; a.P = N * S; a.ray_dim = rd; a.S = S;
    a.experiment = ctx->experiment;
#ifdef NWX_EXPERIMENTS
    if (ctx->experiment == 18) { a.acts = nullptr; a.masks = nullptr; a.hv_out = nullptr; }
#endif
    return nwx::launch_mlp_train_forward(net, a, st);
  };
  const nwx::RngSpec rj = rng_for(o, 0, o->t_rand, o->rng_jitter != 0), ru = rng_for(o, 1, o->u, o->rng_u != 0);
  const nwx::RngSpec rnc = rng_for(o, 2, o->noise_coarse, o->raw_noise_std > 0.f);
  const nwx::RngSpec rnf = rng_for(o, 3, o->noise_fine, o->raw_noise_std > 0.f);
  if ((rc = nwx::launch_coarse_z(io->rays, rd, N, Sc, o->t_vals, o->t_rand, rj, z_c, st))) return rc;
  if ((rc = fwd(NWX_NET_COARSE, z_c, Sc, raw_c))) return rc;
  if ((rc = nwx::launch_composite_fwd(raw_c, z_c, io->rays + 3, rd, o->noise_coarse, rnc, N, Sc, o->white_bkgd, rgb_c,
                                      nullptr, nullptr, nullptr, w_c, nullptr, st))) return rc;
  if ((rc = nwx::launch_sample_pdf(z_c, w_c, Sc, o->u, ru, o->u_lin, Ni, N, z_s, z_f, nullptr, nullptr, st))) return rc;
  if ((rc = fwd(NWX_NET_FINE, z_f, Sf, raw_f))) return rc;
  if ((rc = nwx::launch_composite_fwd(raw_f, z_f, io->rays + 3, rd, o->noise_fine, rnf, N, Sf, o->white_bkgd, rgb_f,
                                      nullptr, nullptr, nullptr, nullptr, nullptr, st))) return rc;
  // ---- loss (training handler:291-305) and backward through compositing; z_samples is detached (:580) ----
  if ((rc = nwx::launch_mse_grad(rgb_c, rgb_f, io->gt_rgb, N, d_rgb[0], d_rgb[1], ctx->loss_scratch, io->loss, st))) return rc;
  if ((rc = nwx::launch_composite_bwd(raw_c, z_c, io->rays + 3, rd, o->noise_coarse, rnc, d_rgb[0], N, Sc, o->white_bkgd,
                                      d_raw[0], st))) return rc;
  if ((rc = nwx::launch_composite_bwd(raw_f, z_f, io->rays + 3, rd, o->noise_fine, rnf, d_rgb[1], N, Sf, o->white_bkgd,
                                      d_raw[1], st))) return rc;
  if ((rc = nwx::launch_embed(io->rays + 8, rd, N, 4, 1.0f, pe_dir, st))) return rc;   // pe(viewdir) per ray
  // ---- backward through the two MLPs ----
  float* grads[2] = {io->grad_coarse, io->grad_fine};
  const int S[2] = {Sc, Sf};
  for (int w = 0; w < 2; ++w) {
    NWX_CUDA_TRY(cudaMemsetAsync(grads[w], 0, sizeof(float) * NWX_PARAMS_PER_NET, st));
    nwx::TrainBwdArgs b{};
    b.d_raw = d_raw[w]; b.hv = hv[w]; b.acts = acts[w]; b.masks = masks[w]; b.gimg = ts + tp.gimg; b.partial = ctx->partial;
    b.head_partial = reinterpret_cast<float*>(ts + tp.head_partial);
    b.fold_scratch = reinterpret_cast<float*>(ts + tp.fold);
    b.pe_dir = pe_dir; b.grad = grads[w]; b.diag = ctx->diag; b.P = N * S[w]; b.S = S[w]; b.max_partials = ctx->n_partials;
    b.which = w;
    b.experiment = ctx->experiment;
    if ((rc = nwx::launch_mlp_backward(ctx->net[w], b, st, ctx->heads_on_side_stream ? ctx->heads_stream : nullptr,
                                       ctx->ev_heads_fork[w], ctx->ev_heads_join[w]))) return rc;
    // data-parallel callers all-reduce the coarse network's gradients underneath the fine network's backward
    if (w == 0 && io->ev_coarse_done) NWX_CUDA_TRY(cudaEventRecord((cudaEvent_t)io->ev_coarse_done, st));
  }
  return NWX_OK;
}
