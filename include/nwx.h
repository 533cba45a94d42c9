/*
 * nwx.h -- C ABI of the B200-native NeRF render/train engine (libnwx.so).
 *
 * Drop-in boundary for the one hot path of dmjovan/NeRF-Workspaces-Explorer:
 * nerf/rays + nerf/models + the _volumetric_rendering body of nerf/inference and
 * nerf/training.  The reference has no FFI layer of its own: its
 * boundary is a set of Python callables operating on torch tensors.  Each entry point
 * below names the reference callable (file:line, relative to the reference root) whose
 * arithmetic it replaces; the Python shim `nwx` (nerf-workspaces-explorer_b200/nwx) keeps
 * the reference's Python signatures and only converts tensors to the pointers below.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to a contiguous row-major buffer owned by the
 *    caller (in practice a torch tensor), unless its comment says "host";
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work
 *    is enqueued on it, nothing synchronises, nothing allocates except nwx_ctx_create /
 *    nwx_load_weights / nwx_ctx_reserve;
 *  - every function returns 0 on success, otherwise an NWX_E_* code or (1000 + cudaError_t);
 *    nothing throws or exits.  nwx_error_string() decodes either;
 *  - re-entrant per (ctx, stream); a ctx must not be used from two host threads at once;
 *  - all floating point is fp32 unless stated; indices are int64 like torch's.
 */
#ifndef NWX_H_
#define NWX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NWX_VERSION 100

enum {
  NWX_OK = 0,
  NWX_E_INVALID = 1,      /* bad argument (null pointer, unsupported size)          */
  NWX_E_NO_WEIGHTS = 2,   /* nwx_load_weights was not called for the requested net  */
  NWX_E_UNSUPPORTED = 3,  /* device is not sm_100 (there is no fallback path)        */
  NWX_E_STALE = 4,        /* nwx_train_pack ran after the last nwx_load_weights: the inference
                             entry points need nwx_load_weights again (host-side biases)  */
  NWX_E_CUDA = 1000       /* 1000 + cudaError_t                                      */
};

enum { NWX_NET_COARSE = 0, NWX_NET_FINE = 1 };

/* Ray record layout of create_rays (rays.py:26-30): o(3) d(3) near far viewdir(3). */
#define NWX_RAY_DIM 11
/* Architecture constants of NeRFModel(D=8, W=256, 63, 27, skips=(4,), use_view_dirs=True)
 * (nerf_model.py:12-43) -- the only architecture the reference instantiates. */
#define NWX_NUM_WEIGHT_TENSORS 24
#define NWX_PARAMS_PER_NET 595844

typedef struct nwx_ctx nwx_ctx;

int nwx_version(void);
const char* nwx_error_string(int code);

/* ---- context: owns the bf16-packed weight images of the coarse and fine networks ------ */
int nwx_ctx_create(int device, nwx_ctx** out);
int nwx_ctx_destroy(nwx_ctx* ctx);

/* Pack one network.  `tensors` is a HOST array of 24 DEVICE pointers to fp32 tensors in
 * NeRFModel.state_dict() order: _pts_linears.{0..7}.{weight,bias}, _views_linears.0.{weight,
 * bias}, _feature_linear.*, _alpha_linear.*, _rgb_linear.* (nerf_model.py:32-41; checkpoint
 * layout of training handler:404-407).  Replaces load_state_dict at inference handler:140-141. */
int nwx_load_weights(nwx_ctx* ctx, int which, const float* const* tensors, void* stream);

/* ---- K1: ray generation + coarse depths ------------------------------------------------ */
/* create_rays (rays.py:6-32): rays of global index [ray0, ray0+nrays) out of the B*H*W rays of B
 * poses (ray = b*H*W + row*W + col).  c2w: [B,16].  rays_out: [nrays, use_view_dirs ? 11 : 8]. */
int nwx_raygen(const float* c2w, int B, int H, int W, float fx, float fy, float cx, float cy,
               float near, float far, int use_view_dirs, int64_t ray0, int64_t nrays,
               float* rays_out, void* stream);

/* z = near*(1-t) + far*t (inference handler:216-220); with t_rand != NULL the stratified jitter
 * of training handler:553-562.  rays: [N, ray_dim]; t_vals: [S] (= torch.linspace(0,1,S));
 * t_rand: [N,S] or NULL; z_out: [N,S]. */
int nwx_coarse_z(const float* rays, int ray_dim, int64_t N, int S, const float* t_vals,
                 const float* t_rand, float* z_out, void* stream);

/* ---- K3: fused positional encoding + 8x256 MLP (tcgen05 / TMEM / TMA) -------------------- */
/* run_network + Embedding.embed + NeRFModel.forward (model_utils.py:13-30, embedding.py:44-48,
 * nerf_model.py:45-83) for the points o + d*z of N rays x S samples.  raw_out: [N,S,4] =
 * (rgb_raw, sigma_raw), pre-activation.  bf16 tensor-core operands, fp32 accumulate/heads. */
int nwx_mlp_forward(nwx_ctx* ctx, int which, const float* rays, int ray_dim, const float* z,
                    int64_t N, int S, float* raw_out, void* stream);

/* Same network on caller-supplied points: what run_network(inputs[N,S,3], viewdirs[N,3], ...)
 * (model_utils.py:13-30) computes without materialising the 90-d embedding.  pts: [P,3];
 * dirs: [P/pts_per_dir, 3] unit view directions, shared by pts_per_dir consecutive points
 * (S for run_network, 1 for one direction per point); raw_out: [P,4]. */
int nwx_mlp_forward_points(nwx_ctx* ctx, int which, const float* pts, const float* dirs,
                           int64_t P, int pts_per_dir, float* raw_out, void* stream);

/* NeRFModel.forward on caller-embedded input x: [P,90] = (pe_xyz 63, pe_dir 27), the literal
 * signature of nerf_model.py:45 (show_endpoint=False); raw_out: [P,4]. */
int nwx_mlp_forward_embedded(nwx_ctx* ctx, int which, const float* x, int64_t P, float* raw_out,
                             void* stream);

/* Embedding.embed (embedding.py:44-48): x [P,3] -> out [P, 3+6*num_freqs] =
 * (x/s, sin(x/s*2^k), cos(x/s*2^k))_k. */
int nwx_embed(const float* x, int64_t P, int num_freqs, float scalar_factor, float* out, void* stream);

/* ---- K4: alpha compositing --------------------------------------------------------------- */
/* raw2outputs (model_utils.py:33-100).  raw: [N,S,4]; z: [N,S]; rays_d: ray directions, row n at
 * rays_d + n*d_stride (a [N,3] tensor: d_stride 3; a ray record: rays+3, d_stride 11);
 * noise: [N,S] already scaled by raw_noise_std, or NULL.  Outputs rgb [N,3], disp/acc/depth [N];
 * weights [N,S] may be NULL.  flags (may be NULL): one int32 word, bit0 = NaN seen, bit1 = Inf
 * seen in any output (replaces the 22 host syncs of inference handler:273-275). */
int nwx_composite_fwd(const float* raw, const float* z, const float* rays_d, int d_stride,
                      const float* noise, int64_t N, int S, int white_bkgd, float* rgb, float* disp,
                      float* acc, float* depth, float* weights, int32_t* flags, void* stream);

/* d(loss)/d(raw) given d(loss)/d(rgb_map) [N,3]; analytic backward of the above (what autograd
 * computes for training handler:305-308).  Needs the forward's weights [N,S]. d_raw: [N,S,4]. */
int nwx_composite_bwd(const float* raw, const float* z, const float* rays_d, int d_stride,
                      const float* noise, const float* weights, const float* d_rgb, int64_t N, int S,
                      int white_bkgd, float* d_raw, void* stream);

/* ---- K2: hierarchical resampling + merge --------------------------------------------------- */
/* sample_pdf(z_mid, weights[:,1:-1], n_imp, det) followed by sort(cat(z, z_samples))
 * (rays.py:74-121; inference handler:236-243).  z_c, w_c: [N,Sc] (Sc <= 128); u: [N,n_imp]
 * uniforms or NULL for det (then u_lin [n_imp] = torch.linspace(0,1,n_imp) is used);
 * outputs: z_samples [N,n_imp], z_fine [N,Sc+n_imp] (NULL to skip the merge), inds int64
 * [N,n_imp] (NULL to skip) = the searchsorted result of rays.py:103, z_std [N] (NULL to skip)
 * = std(z_samples, unbiased=False) of inference handler:267. n_imp <= 256. */
int nwx_sample_pdf(const float* z_c, const float* w_c, int Sc, const float* u, const float* u_lin,
                   int n_imp, int64_t N, float* z_samples, float* z_fine, int64_t* inds,
                   float* z_std, void* stream);

/* The literal signature of sample_pdf (rays.py:74): bins [N,M], weights [N,M-1], no merge. */
int nwx_sample_pdf_bins(const float* bins, const float* weights, int M, const float* u,
                        const float* u_lin, int n_imp, int64_t N, float* samples, int64_t* inds,
                        float* cdf_out, void* stream);

/* ---- whole chunk: the body of _volumetric_rendering ---------------------------------------- */
typedef struct nwx_render_opts {
  int n_samples;          /* 64  (yaml rendering.n_samples)    */
  int n_importance;       /* 128 (yaml rendering.n_importance) */
  int white_bkgd;         /* yaml rendering.white_background   */
  int ray_dim;            /* 11                                */
  const float* t_vals;    /* [n_samples]  linspace(0,1)        */
  const float* u_lin;     /* [n_importance] linspace(0,1)      */
  const float* t_rand;    /* [N,n_samples] or NULL  (training handler:560) */
  const float* u;         /* [N,n_importance] or NULL = det (rays.py:95-98) */
  const float* noise_coarse; /* [N,n_samples] scaled, or NULL  (model_utils.py:65) */
  const float* noise_fine;   /* [N,n_samples+n_importance] scaled, or NULL */
  /* In-kernel random source (counter-based Philox, csrc/rng.cuh) for the draws above whose pointer
   * is NULL -- nothing is stored in HBM and the backward regenerates the forward's noise: */
  uint64_t rng_seed;         /* key                                                             */
  uint64_t rng_offset;       /* e.g. the step number: fresh numbers every step                  */
  float raw_noise_std;       /* > 0 and noise_* == NULL: sigma noise N(0,1)*std generated in place */
  int rng_jitter;            /* != 0 and t_rand == NULL: stratified jitter generated in place     */
  int rng_u;                 /* != 0 and u == NULL: random importance uniforms (else deterministic) */
} nwx_render_opts;

typedef struct nwx_render_out {       /* any pointer may be NULL = not wanted; one of rgb_fine / rgb8_fine is required */
  float *rgb_coarse, *disp_coarse, *acc_coarse, *depth_coarse, *raw_coarse;
  float *rgb_fine, *disp_fine, *acc_fine, *depth_fine, *raw_fine, *z_std;
  float *z_vals_coarse, *weights_coarse, *z_samples, *z_vals_fine, *weights_fine;
  int64_t* inds;
  int32_t* flags;
  uint8_t* rgb8_fine;                 /* to8b(rgb_fine) (model_utils.py:9), [N,3]; written by the compositing kernel itself */
} nwx_render_out;

/* Reserve the context's scratch for chunks of up to max_rays rays (otherwise grown on demand,
 * which allocates). */
int nwx_ctx_reserve(nwx_ctx* ctx, int64_t max_rays, int n_samples, int n_importance);

/* Size of the context's scratch and a generation counter that changes whenever it is re-allocated.  A caller that
 * captures nwx_render_rays into a CUDA graph (the graph bakes in the scratch address and the biases, which ride in
 * the kernel parameters) must re-capture when the generation changed or nwx_load_weights ran. */
int nwx_ctx_scratch_state(nwx_ctx* ctx, int64_t* bytes /* host */, int64_t* generation /* host */);

/* inference handler:203-277 / training handler:534-618 for N rays [N,ray_dim]. */
int nwx_render_rays(nwx_ctx* ctx, const float* rays, int64_t N, const nwx_render_opts* opts,
                    const nwx_render_out* out, void* stream);

/* out[i] = the library's counter-based draw for element i of (seed, offset, rng_stream): kind 0 =
 * U[0,1), kind 1 = N(0,1)*scale.  rng_stream: 0 jitter, 1 importance u, 2 / 3 sigma noise of the
 * coarse / fine pass.  Exactly what the kernels generate in place (tests inject it back as tensors). */
int nwx_rng_fill(int kind, uint64_t seed, uint64_t offset, uint32_t rng_stream, float scale, int64_t n,
                 float* out, void* stream);

/* (255*clip(x,0,1)).astype(uint8) (model_utils.py:9) over n floats. */
int nwx_to8b(const float* x, int64_t n, uint8_t* out, void* stream);

/* ---- training: the autograd part of NeRFReplicaTrainingHandler.step (training handler:277-315) -- */
/* Master parameters, gradients and Adam moments of one network are flat fp32 device buffers of
 * NWX_PARAMS_PER_NET floats in state_dict order; nwx_param_offsets gives the 24 tensor offsets. */
int nwx_param_offsets(int* offsets24 /* host, 24 ints */);

/* Re-pack one network (bf16 forward images incl. the folded views layer W_view[:, :256] . W_feature,
 * transposed images for the backward, device-side biases/heads) from its flat master parameters.
 * Stream-ordered, no host synchronisation; call after every optimiser step.  (nwx_load_weights is the
 * inference-time equivalent.)  params_flat must stay allocated until the next nwx_train_pack of this
 * network: the backward reads the fp32 W_view / W_feature / b_feature from it for the chain rule
 * through the fold (d W_fold -> d W_view, d W_feature, d b_feature). */
int nwx_train_pack(nwx_ctx* ctx, int which, const float* params_flat, void* stream);

typedef struct nwx_train_io {
  const float* rays;       /* [N, ray_dim] sampled rays (training handler:341-370)              */
  const float* gt_rgb;     /* [N,3] ground-truth pixels                                         */
  float* grad_coarse;      /* out: [NWX_PARAMS_PER_NET] d(loss)/d(params), overwritten          */
  float* grad_fine;        /* out                                                               */
  double* loss;            /* out: [2] = mse(rgb_coarse, gt), mse(rgb_fine, gt) (handler:291-298) */
  float* rgb_coarse;       /* optional out [N,3]                                                */
  float* rgb_fine;         /* optional out [N,3]                                                */
  void* ev_coarse_done;    /* optional cudaEvent_t, recorded on `stream` as soon as grad_coarse is final
                              (before the fine network's backward is enqueued): a data-parallel caller
                              all-reduces the coarse gradients on a side stream underneath it */
} nwx_train_io;

/* _sample_training_data (training handler:341-370) on the device: ONE random image of the bank and n random
 * pixels of it (with replacement) from the library's counter-based generator (rng_stream 4: image index,
 * element 0; rng_stream 5: pixel index of sample r, element r), then the gather.  rays_bank: [num_img, num_ray,
 * ray_dim] (what initialize_rays builds, :243-263); rgb_bank: [num_img, num_ray, 3]; rays_out: [n, ray_dim];
 * gt_out: [n,3]; idx_out (may be NULL): int64 [1 + n] = the image index, then the n pixel indices.  No host
 * synchronisation (the reference draws on the CPU and fancy-indexes). */
int nwx_sample_training_batch(const float* rays_bank, const float* rgb_bank, int num_img, int64_t num_ray,
                              int ray_dim, int64_t n, uint64_t seed, uint64_t offset, float* rays_out,
                              float* gt_out, int64_t* idx_out, void* stream);

/* Forward in training mode (jitter t_rand, noise, random u from opts; training handler:534-618),
 * loss = mse(rgb_coarse) + mse(rgb_fine), backward to both networks' parameters.  What
 * total_loss.backward() (handler:305-308) computes; no gradient flows through z_samples (:580). */
int nwx_train_fwd_bwd(nwx_ctx* ctx, const nwx_train_io* io, int64_t N, const nwx_render_opts* opts,
                      void* stream);

/* torch.optim.Adam step (default betas/eps are the caller's to pass; handler:234) on flat buffers;
 * grads are multiplied by grad_scale first (1/world_size after a sum all-reduce). step is 1-based. */
int nwx_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1,
                  float beta2, float eps, int step, float grad_scale, void* stream);

/* The optimiser step of the training loop in two launches: Adam (as nwx_adam_step) over BOTH networks' flat buffers
 * -- params / grads / m / v are [2][NWX_PARAMS_PER_NET], coarse then fine -- fused with the re-pack of nwx_train_pack
 * (the thread that updates a parameter also writes its bf16 / fp32 copies into the kernels' images), then the folded
 * views layer of both networks.  Needs one nwx_train_pack per network beforehand (allocation, zero padding).  `params`
 * must stay allocated like params_flat of nwx_train_pack. */
int nwx_adam_pack_step(nwx_ctx* ctx, float* params, const float* grads, float* m, float* v, float lr, float beta1,
                       float beta2, float eps, int step, float grad_scale, void* stream);

/* Per-stage device timing of nwx_render_rays (CUDA events on the caller's stream).  Stage order:
 * coarse_z, dirbias(coarse), mlp(coarse), composite(coarse), sample_pdf, dirbias(fine), mlp(fine),
 * composite(fine, incl. the uint8 pixels).  nwx_ctx_stage_ms waits for the last recorded call and fills
 * ms_out[NWX_NUM_STAGES] (host). */
#define NWX_NUM_STAGES 8
int nwx_ctx_set_profiling(nwx_ctx* ctx, int on);
int nwx_ctx_stage_ms(nwx_ctx* ctx, float* ms_out);

/* ---- introspection for bench/tests --------------------------------------------------------- */
/* Number of kernels this library has launched since load (all contexts). */
int64_t nwx_launch_count(void);
/* MLP kernel variant: 0/1 = production (CTA pair, resident weights, _feature_linear folded into the
 * views layer at weight-load time -- exact algebra, nerf_model.py:64-68 has no non-linearity between
 * them), 2 = CTA pair streaming, 3 = single CTA (cta_group::1), 4 = as 1 with the reference's layer
 * structure (no fold).  Tests use it to cross-check variants; see DESIGN.md. */
int nwx_set_mlp_variant(nwx_ctx* ctx, int variant);
/* Diagnostics: tap the post-activation fp32 output of tensor-core layer `layer` (0..9) of
 * subsequent MLP launches into out [P,256] (NULL = off); register a host-mapped uint32[4] that
 * a barrier wait that exceeds its wall-clock bound (10 s) fills before the kernel traps instead of hanging. */
int nwx_debug_tap(nwx_ctx* ctx, int layer, float* out);
int nwx_debug_diag(nwx_ctx* ctx, uint32_t* host_mapped);
/* Test hook: copy one packed buffer of network `which` to dst (device, exactly `bytes` long): what = 0 forward weight
 * image, 1 transposed image (dX), 2 device-side constants, 3 view-direction table, 4 views bias, 5 folded views bias. */
int nwx_debug_copy_packed(nwx_ctx* ctx, int which, int what, void* dst, int64_t bytes, void* stream);
/* Timing experiments in the training kernels (results wrong on purpose; tools/train_experiments.py): 0 = none,
 * 11 = epilogues do not wait for their tile's previous TMA store, 12 = no TMA stores of the tile images, 13-15 see
 * csrc/context.cu.  Compiled in only by `make EXPERIMENTS=1`; the product library returns NWX_E_INVALID for any
 * code but 0. */
int nwx_debug_experiment(nwx_ctx* ctx, int code);
/* The 4 diagnostic words (0xDEADxxxx | waiter code, block, barrier, parity) of the last aborted wait; zeros
 * if none.  Every context owns such a host-mapped word from creation (nwx_debug_diag(ctx, NULL) restores
 * it), so the cause of a trap is readable even though the CUDA context is unusable afterwards. */
int nwx_ctx_last_diag(nwx_ctx* ctx, uint32_t* out4 /* host */);

#ifdef __cplusplus
}
#endif
#endif /* NWX_H_ */
