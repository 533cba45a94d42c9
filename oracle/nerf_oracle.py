"""CPU oracle for the NeRF render/train hot path -- TEST INFRASTRUCTURE ONLY.

This module restates, with plain torch-CPU ops, the algorithm of the reference
hot path (dmjovan/NeRF-Workspaces-Explorer: nerf/rays, nerf/models, nerf/inference,
nerf/training).  It exists to *check* the CUDA engine; nothing under
``nerf-workspaces-explorer_b200/`` may import it.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py``.

Parity pinning: the reference ships NO tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned against *outputs of the reference
itself*: ``tests/golden/make_golden.py`` imports the reference modules from
``/root/reference`` in the build container, runs them and this oracle on the same
seeded inputs, asserts bit-equality, and freezes the results as fixtures under
``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks the oracle against
those fixtures wherever the tests run (the reference itself does not travel).

Every function cites the reference file:line it follows.  The arithmetic is kept
op-for-op identical (same torch ops in the same order, fp32) because bit-exact
ray / searchsorted indices are part of the parity contract.

The render / training functions follow the device of their inputs (constants are created where
the reference creates them -- on the CPU -- and moved like its ``.cuda()`` calls do): with CPU
tensors this is the pinned oracle; with CUDA tensors it is what the reference executes on a GPU
(torch eager, fp32), used as the ``gpu_eager_baseline`` of bench.py and as the fp32 training
reference of tests/test_gpu_trained.py.  CUDA results are NOT bit-pinned (cuBLAS / CUB orders).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------- #
# rays  (reference: nerf/rays/rays.py)
# --------------------------------------------------------------------------- #


def camera_directions(num_images: int, height: int, width: int, fx: float, fy: float,
                      cx: float, cy: float) -> torch.Tensor:
    """Pinhole directions in the camera frame, [B,H,W,3].  rays.py:35-58.

    OpenCV convention (x right, y down, z forward), integer pixel centres, fp32.
    """
    col = torch.arange(width).float()[None, :].expand(height, width)   # i (x)   rays.py:41-42
    row = torch.arange(height).float()[:, None].expand(height, width)  # j (y)   rays.py:41,43
    col_b = col[None].expand(num_images, height, width).contiguous()    # rays.py:47-50
    row_b = row[None].expand(num_images, height, width).contiguous()
    x = (col_b - cx) / fx                                               # rays.py:52
    y = (row_b - cy) / fy                                               # rays.py:53
    z = torch.ones(num_images, height, width)                           # rays.py:54
    return torch.stack((x, y, z), dim=3)                                # rays.py:56


def world_rays(c2w: torch.Tensor, dirs_cam: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Rotate camera directions into the world frame.  rays.py:61-71."""
    rot = c2w[:, :3, :3]
    dirs_w = torch.matmul(rot[:, None, ...], dirs_cam[..., None]).squeeze(-1)   # rays.py:67
    origins = c2w[:, :3, -1]
    origins = torch.broadcast_tensors(origins[:, None, :], dirs_w)[0]          # rays.py:69
    return origins, dirs_w


def create_rays(num_images: int, Ts_c2w: torch.Tensor, height: int, width: int, fx: float, fy: float,
                cx: float, cy: float, near: float, far: float, use_view_dirs: bool = True) -> torch.Tensor:
    """[B,H*W,11] rows (o, d, near, far, d/|d|); ray index = row*W + col.  rays.py:6-32."""
    dirs_cam = camera_directions(num_images, height, width, fx, fy, cx, cy).view(num_images, -1, 3)
    rays_o, rays_d = world_rays(Ts_c2w, dirs_cam)
    near_t = near * torch.ones_like(rays_d[..., :1])                            # rays.py:26
    far_t = far * torch.ones_like(rays_d[..., :1])
    parts = [rays_o, rays_d, near_t, far_t]
    if use_view_dirs:
        parts.append(rays_d / torch.norm(rays_d, dim=-1, keepdim=True).float())  # rays.py:24
    return torch.cat(parts, -1)


def pdf_to_cdf(weights: torch.Tensor) -> torch.Tensor:
    """weights [N,S-2] -> cdf [N,S-1] with a leading 0.  rays.py:87-90."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)


def invert_cdf(bins: torch.Tensor, cdf: torch.Tensor, u: torch.Tensor
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Inverse-CDF lookup: returns (samples [N,Nu], inds int64 [N,Nu]).  rays.py:102-119."""
    u = u.contiguous()
    inds = torch.searchsorted(cdf.detach(), u, right=True)                      # rays.py:103
    below = torch.clamp(inds - 1, min=0)                                        # rays.py:104
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)                            # rays.py:105
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)     # rays.py:110
    bin_b, bin_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)   # rays.py:111
    denom = cdf_a - cdf_b                                                       # rays.py:113
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)            # rays.py:114
    t = (u - cdf_b) / denom                                                     # rays.py:118
    return bin_b + t * (bin_a - bin_b), inds                                    # rays.py:119


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, n_samples: int, det: bool = False,
               u: Optional[torch.Tensor] = None, return_inds: bool = False):
    """Hierarchical resampling.  rays.py:74-121.  ``u`` injects the uniforms the
    reference would draw with torch.rand (rays.py:98) so runs are repeatable."""
    cdf = pdf_to_cdf(weights)
    if u is None:
        if det:
            u = torch.linspace(0., 1., steps=n_samples).to(cdf.device)           # rays.py:95 (+ .cuda() :100)
            u = u.expand(list(cdf.shape[:-1]) + [n_samples])
        else:
            u = torch.rand(list(cdf.shape[:-1]) + [n_samples]).to(cdf.device)    # rays.py:98
    samples, inds = invert_cdf(bins, cdf, u)
    return (samples, inds) if return_inds else samples


# --------------------------------------------------------------------------- #
# model  (reference: nerf/models/embedding.py, nerf_model.py)
# --------------------------------------------------------------------------- #


def positional_encoding(x: torch.Tensor, num_freqs: int, scalar_factor: float = 1.0) -> torch.Tensor:
    """[P,3] -> [P,3+6L]: (x/s, sin(x/s*2^k), cos(x/s*2^k))_k.  embedding.py:24-48."""
    xs = x / scalar_factor                                                      # embedding.py:48
    freqs = 2. ** torch.linspace(0., num_freqs - 1, steps=num_freqs)            # embedding.py:32
    out = [xs]
    for f in freqs:
        out.append(torch.sin(xs * f))                                           # embedding.py:36
        out.append(torch.cos(xs * f))
    return torch.cat(out, -1)


STATE_KEYS: Tuple[str, ...] = tuple(
    [f"_pts_linears.{i}.{p}" for i in range(8) for p in ("weight", "bias")]
    + [f"_views_linears.0.{p}" for p in ("weight", "bias")]
    + [f"_{n}_linear.{p}" for n in ("feature", "alpha", "rgb") for p in ("weight", "bias")]
)
"""state_dict order of NeRFModel(use_view_dirs=True).  nerf_model.py:32-41."""


def init_state_dict(seed: int, alpha_bias: Optional[float] = 0.1, trained_like: bool = False,
                    generator: Optional[torch.Generator] = None) -> Dict[str, torch.Tensor]:
    """Random-init weights of the reference architecture (8x256, skip 4, view branch).

    Draw order and distributions reproduce ``nn.Linear`` default init in the module
    construction order of nerf_model.py:32-41 (kaiming-uniform(a=sqrt(5)) weight, then
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) bias), so ``torch.manual_seed(seed)`` followed by
    the reference constructor gives the same tensors (asserted by make_golden.py).
    SURVEY.md section 7: sigma's sign at the last sample is random under raw default init,
    which makes outputs discontinuous; ``alpha_bias`` pins ``_alpha_linear.bias``.
    """
    g = generator or torch.Generator().manual_seed(seed)
    shapes = ([(256, 63)] + [(256, 256)] * 4 + [(256, 319)] + [(256, 256)] * 2  # pts_linears
              + [(128, 283)]                                                    # views
              + [(256, 256), (1, 256), (3, 128)])                               # feature, alpha, rgb
    names = ([f"_pts_linears.{i}" for i in range(8)] + ["_views_linears.0"]
             + ["_feature_linear", "_alpha_linear", "_rgb_linear"])
    sd: Dict[str, torch.Tensor] = {}
    for name, (fo, fi) in zip(names, shapes):
        bound = 1.0 / math.sqrt(fi)
        # kaiming_uniform_(a=sqrt(5)): gain=sqrt(2/(1+5)), bound=gain*sqrt(3/fan_in)=1/sqrt(fan_in)
        sd[f"{name}.weight"] = torch.empty(fo, fi).uniform_(-bound, bound, generator=g)
        sd[f"{name}.bias"] = torch.empty(fo).uniform_(-bound, bound, generator=g)
    if trained_like:  # SURVEY.md section 8d "trained-like" stress set
        sd["_alpha_linear.weight"] *= 30.0
        sd["_rgb_linear.weight"] *= 10.0
        sd["_alpha_linear.bias"].fill_(1.0)
    elif alpha_bias is not None:
        sd["_alpha_linear.bias"].fill_(alpha_bias)
    return sd


def mlp_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """NeRFModel.forward, use_view_dirs=True, show_endpoint=False.  nerf_model.py:45-83.

    x [P,90] = (pe_xyz 63, pe_dir 27) -> [P,4] = (rgb_raw 3, sigma_raw 1), pre-activation.
    """
    lin = torch.nn.functional.linear
    pts, views = torch.split(x, [63, 27], dim=-1)                                # nerf_model.py:53
    h = pts
    for i in range(8):
        h = torch.relu(lin(h, sd[f"_pts_linears.{i}.weight"], sd[f"_pts_linears.{i}.bias"]))
        if i == 4:
            h = torch.cat([pts, h], -1)                                         # nerf_model.py:58-59
    alpha = lin(h, sd["_alpha_linear.weight"], sd["_alpha_linear.bias"])         # nerf_model.py:63
    feat = lin(h, sd["_feature_linear.weight"], sd["_feature_linear.bias"])      # nerf_model.py:64
    h = torch.cat([feat, views], -1)                                            # nerf_model.py:66
    h = torch.relu(lin(h, sd["_views_linears.0.weight"], sd["_views_linears.0.bias"]))
    rgb = lin(h, sd["_rgb_linear.weight"], sd["_rgb_linear.bias"])               # nerf_model.py:74
    return torch.cat([rgb, alpha], -1)                                          # nerf_model.py:76


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def mlp_forward_bf16_emul(sd: Dict[str, torch.Tensor], pe_xyz: torch.Tensor, pe_dir: torch.Tensor,
                          fold_feature: bool = True) -> torch.Tensor:
    """Numerics model of the CUDA engine's fused MLP (NOT reference behaviour).

    Same graph as :func:`mlp_forward`, with operands rounded to bf16 exactly where the
    tcgen05 kernel rounds them (DESIGN.md "numerics"): PE features, every tensor-core weight
    and every hidden activation fed to a tensor-core layer are bf16; accumulation, biases,
    the 27-d view-direction contribution, the sigma head and the rgb head stay fp32.
    Used by tests to check the kernel far tighter than the 1e-3 reference tolerance.

    fold_feature=True models the production inference kernel, which folds ``_feature_linear`` into the
    views layer at load time (no non-linearity between them, nerf_model.py:64-68): one 128x256 bf16
    weight ``W_view[:, :256] @ W_feature`` (fp64 product, rounded once) applied to bf16 h8, its bias
    ``b_view + W_view[:, :256] @ b_feature`` joining the per-ray fp32 term.  fold_feature=False models
    the kernels that keep the reference's layer structure (training forward, MLP variants 2-4).
    """
    lin = torch.nn.functional.linear
    W = lambda k: _bf16(sd[k])
    pts = _bf16(pe_xyz)
    h = pts
    for i in range(8):
        acc = lin(h, W(f"_pts_linears.{i}.weight")) + sd[f"_pts_linears.{i}.bias"]
        h32 = torch.relu(acc)
        h = _bf16(h32)
        if i == 4:
            h = torch.cat([pts, h], -1)
    alpha = lin(h32, sd["_alpha_linear.weight"], sd["_alpha_linear.bias"])        # fp32 head on fp32 h7
    wv = sd["_views_linears.0.weight"]
    if fold_feature:
        w_fold = (wv[:, :256].double() @ sd["_feature_linear.weight"].double()).float()
        b_fold = (sd["_views_linears.0.bias"].double()
                  + wv[:, :256].double() @ sd["_feature_linear.bias"].double()).float()
        hv = torch.relu(lin(h, _bf16(w_fold)) + lin(pe_dir, wv[:, 256:], b_fold))
    else:
        feat = _bf16(lin(h, W("_feature_linear.weight")) + sd["_feature_linear.bias"])
        dir_bias = lin(pe_dir, wv[:, 256:], sd["_views_linears.0.bias"])         # fp32, per ray
        hv = torch.relu(lin(feat, _bf16(wv[:, :256])) + dir_bias)
    rgb = lin(hv, sd["_rgb_linear.weight"], sd["_rgb_linear.bias"])
    return torch.cat([rgb, alpha], -1)


def batchify(fn: Callable, chunk: Optional[int]) -> Callable:
    """utils/batch_utils.py:28-39."""
    if chunk is None:
        return fn
    return lambda inputs: torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)


def run_network(inputs: torch.Tensor, viewdirs: Optional[torch.Tensor], fn: Callable,
                embed_fn: Callable, embeddirs_fn: Optional[Callable], netchunk: Optional[int] = 1024 * 64
                ) -> torch.Tensor:
    """Embed points (+ per-point broadcast view dirs) and apply ``fn`` in chunks.  model_utils.py:13-30."""
    flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded = embed_fn(flat)
    if viewdirs is not None:
        dirs = viewdirs[:, None].expand(inputs.shape)                            # model_utils.py:23
        embedded = torch.cat([embedded, embeddirs_fn(torch.reshape(dirs, [-1, dirs.shape[-1]]))], -1)
    out = batchify(fn, netchunk)(embedded)
    return torch.reshape(out, list(inputs.shape[:-1]) + [out.shape[-1]])


# --------------------------------------------------------------------------- #
# compositing  (reference: nerf/models/model_utils.py:33-100, cuda_enabled=False branch)
# --------------------------------------------------------------------------- #


def raw2outputs(raw: torch.Tensor, z_vals: torch.Tensor, rays_d: torch.Tensor, raw_noise_std: float = 0,
                white_bkgd: bool = False, noise: Optional[torch.Tensor] = None):
    """Alpha compositing -> (rgb, disp, acc, weights, depth).  model_utils.py:33-100.

    ``noise`` injects the N(0,1)*std draw of model_utils.py:65 (already scaled)."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]                                   # :51
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], -1)        # :56 (Tensor([1e10]).expand)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)                     # :60
    rgb = torch.sigmoid(raw[..., :3])                                            # :62
    if noise is None:
        noise = torch.randn(raw[..., 3].shape, device=raw.device) * raw_noise_std if raw_noise_std > 0. else 0.
    alpha = 1. - torch.exp(-torch.relu(raw[..., 3] + noise) * dists)             # :49,:71
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1. - alpha + 1e-10], -1), -1)[:, :-1]   # :73-77
    weights = alpha * trans                                                      # :79-80
    rgb_map = torch.sum(weights[..., None] * rgb, -2)                            # :84
    depth_map = torch.sum(weights * z_vals, -1)                                  # :93
    disp_map = 1. / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / torch.sum(weights, -1))
    acc_map = torch.sum(weights, -1)                                             # :95
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc_map[..., None])                            # :98
    return rgb_map, disp_map, acc_map, weights, depth_map


# --------------------------------------------------------------------------- #
# volumetric rendering  (reference: nerf/inference/...handler.py:203-277,
#                                   nerf/training/...handler.py:534-618)
# --------------------------------------------------------------------------- #


class RenderConfig:
    """Scalars the handlers read from YAML (office_tokyo_config.yaml:17-30,41)."""

    def __init__(self, n_samples: int = 64, n_importance: int = 128, num_freqs_3d: int = 10,
                 num_freqs_2d: int = 4, white_bkgd: bool = False, perturb: float = 1.0,
                 raw_noise_std: float = 1.0, chunk: int = 1024 * 8, net_chunk: int = 1024 * 32):
        self.n_samples, self.n_importance = n_samples, n_importance
        self.num_freqs_3d, self.num_freqs_2d = num_freqs_3d, num_freqs_2d
        self.white_bkgd, self.perturb, self.raw_noise_std = white_bkgd, perturb, raw_noise_std
        self.chunk, self.net_chunk = chunk, net_chunk


def coarse_z(ray_batch: torch.Tensor, n_samples: int, t_rand: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Linear-in-depth sample positions (+ optional stratified jitter).

    inference handler:213-220; training handler:544-562 (t_rand = the torch.rand of :560)."""
    bounds = torch.reshape(ray_batch[..., 6:8], [-1, 1, 2])
    near, far = bounds[..., 0], bounds[..., 1]
    t_vals = torch.linspace(0., 1., steps=n_samples).to(ray_batch.device)       # computed on the CPU, then .cuda() (:216)
    z = near * (1. - t_vals) + far * t_vals
    z = z.expand([ray_batch.shape[0], n_samples])
    if t_rand is not None:
        mids = .5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    return z


def volumetric_rendering(ray_batch: torch.Tensor, sd_coarse: Dict[str, torch.Tensor],
                         sd_fine: Dict[str, torch.Tensor], cfg: RenderConfig, train_mode: bool = False,
                         t_rand: Optional[torch.Tensor] = None, u: Optional[torch.Tensor] = None,
                         noise_coarse: Optional[torch.Tensor] = None,
                         noise_fine: Optional[torch.Tensor] = None,
                         mlp: Callable = mlp_forward) -> Dict[str, torch.Tensor]:
    """One ray chunk -> the handlers' 11-key output dict (+ ``z_vals_fine``/``z_samples``/``inds`` extras
    used only by tests).  inference handler:203-277 (train_mode=False: no jitter, no noise, det
    fine sampling, :238) and training handler:534-618 (train_mode=True: the three random draws are
    injected through t_rand / u / noise_*)."""
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 8 else None
    embed = lambda x: positional_encoding(x, cfg.num_freqs_3d, 10)
    embed_d = lambda x: positional_encoding(x, cfg.num_freqs_2d, 1)

    jitter = t_rand if (train_mode and cfg.perturb > 0.) else None
    z_vals = coarse_z(ray_batch, cfg.n_samples, jitter)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    noise_std = cfg.raw_noise_std if train_mode else 0
    raw_c = run_network(pts, viewdirs, lambda x: mlp(sd_coarse, x), embed, embed_d, cfg.net_chunk)
    rgb_c, disp_c, acc_c, w_c, depth_c = raw2outputs(raw_c, z_vals, rays_d, noise_std, cfg.white_bkgd,
                                                     noise=noise_coarse if train_mode else None)
    z_mid = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
    det = (cfg.perturb == 0.) or (not train_mode)
    z_samples, inds = sample_pdf(z_mid, w_c[..., 1:-1], cfg.n_importance, det=det,
                                 u=None if det else u, return_inds=True)
    z_samples = z_samples.detach()
    z_fine, _ = torch.sort(torch.cat([z_vals, z_samples], -1), -1)
    pts_f = rays_o[..., None, :] + rays_d[..., None, :] * z_fine[..., :, None]
    raw_f = run_network(pts_f, viewdirs, lambda x: mlp(sd_fine, x), embed, embed_d, cfg.net_chunk)
    rgb_f, disp_f, acc_f, w_f, depth_f = raw2outputs(raw_f, z_fine, rays_d, noise_std, cfg.white_bkgd,
                                                     noise=noise_fine if train_mode else None)
    return {
        "rgb_coarse": rgb_c, "disp_coarse": disp_c, "acc_coarse": acc_c, "depth_coarse": depth_c,
        "raw_coarse": raw_c, "rgb_fine": rgb_f, "disp_fine": disp_f, "acc_fine": acc_f,
        "depth_fine": depth_f, "z_std": torch.std(z_samples, dim=-1, unbiased=False), "raw_fine": raw_f,
        # extras (not in the reference dict) for stage-level parity checks
        "z_vals_coarse": z_vals, "weights_coarse": w_c, "z_samples": z_samples, "inds": inds,
        "z_vals_fine": z_fine, "weights_fine": w_f,
    }


REFERENCE_KEYS = ("rgb_coarse", "disp_coarse", "acc_coarse", "depth_coarse", "raw_coarse", "rgb_fine",
                  "disp_fine", "acc_fine", "depth_fine", "z_std", "raw_fine")


def render_rays(flat_rays: torch.Tensor, sd_coarse, sd_fine, cfg: RenderConfig,
                keys: Sequence[str] = REFERENCE_KEYS, **kw) -> Dict[str, torch.Tensor]:
    """Chunked render of [n,11] rays.  utils/batch_utils.py:7-25 + inference handler:187-201."""
    outs: Dict[str, List[torch.Tensor]] = {}
    for i in range(0, flat_rays.shape[0], cfg.chunk):
        part = volumetric_rendering(flat_rays[i:i + cfg.chunk], sd_coarse, sd_fine, cfg, **kw)
        for k in keys:
            outs.setdefault(k, []).append(part[k])
    return {k: torch.cat(v, 0) for k, v in outs.items()}


def to8b(x: np.ndarray) -> np.ndarray:
    """model_utils.py:9."""
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


def render_image(c2w: torch.Tensor, sd_coarse, sd_fine, cfg: RenderConfig, height: int, width: int,
                 fx: float, fy: float, cx: float, cy: float, near: float, far: float) -> np.ndarray:
    """render_coordinates minus the COORD->pose step.  inference handler:166-185."""
    with torch.no_grad():
        rays = create_rays(c2w.shape[0], c2w, height, width, fx, fy, cx, cy, near, far, True)
        out = render_rays(rays[0], sd_coarse, sd_fine, cfg, keys=("rgb_fine",))
        return to8b(out["rgb_fine"].numpy().reshape((height, width, 3)))


# --------------------------------------------------------------------------- #
# training math  (reference: nerf/training/...handler.py:277-315)
# --------------------------------------------------------------------------- #


def training_loss_and_grads(rays: torch.Tensor, gt_rgb: torch.Tensor, sd_coarse, sd_fine, cfg: RenderConfig,
                            t_rand, u, noise_coarse, noise_fine):
    """loss = mse(rgb_c, gt) + mse(rgb_f, gt) and d(loss)/d(params) via autograd.
    training handler:288-308 (GT is float64 there, replica_dataset.py:114, so the loss is fp64)."""
    pc = {k: v.clone().requires_grad_(True) for k, v in sd_coarse.items()}
    pf = {k: v.clone().requires_grad_(True) for k, v in sd_fine.items()}
    out = volumetric_rendering(rays, pc, pf, cfg, train_mode=True, t_rand=t_rand, u=u,
                               noise_coarse=noise_coarse, noise_fine=noise_fine)
    gt = gt_rgb.double()
    loss_c = torch.mean((out["rgb_coarse"] - gt) ** 2)
    loss_f = torch.mean((out["rgb_fine"] - gt) ** 2)
    (loss_c + loss_f).backward()
    return (loss_c.detach(), loss_f.detach(), {k: v.grad for k, v in pc.items()},
            {k: v.grad for k, v in pf.items()}, out)


def adam_step(param: torch.Tensor, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int,
              lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
    """torch.optim.Adam defaults (training handler:234), single tensor, in place; ``step`` is 1-based."""
    m.mul_(beta1).add_(grad, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    param.addcdiv_(m, denom, value=-(lr / bc1))


def lr_at(step: int, lr0: float = 5e-4, decay_rate: float = 0.1, decay_steps: int = 50000) -> float:
    """training handler:312-313."""
    return lr0 * (decay_rate ** (step / decay_steps))


# --------------------------------------------------------------------------- #
# synthetic workloads  
# --------------------------------------------------------------------------- #


def intrinsics(height: int, width: int, hfov_deg: float = 90.0) -> Tuple[float, float, float, float]:
    """fx, fy, cx, cy as the handlers derive them.  inference handler:67-74."""
    fx = width / 2.0 / math.tan(math.radians(hfov_deg / 2.0))
    return fx, fx, (width - 1.0) / 2.0, (height - 1.0) / 2.0


def _rot(axis: str, th: float) -> np.ndarray:
    c, s = np.cos(th), np.sin(th)
    m = {"yaw": [[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]],
         "pitch": [[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]],
         "roll": [[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]}[axis]
    return np.array(m, dtype=np.float32)


def rodrigues(rvec: Sequence[float]) -> np.ndarray:
    """Axis-angle -> 3x3 rotation in float64 (what cv2.Rodrigues returns for a float64 vector,
    utils/camera_poses.py:62-63), so the synthetic poses need no OpenCV on the GPU box."""
    r = np.asarray(rvec, dtype=np.float64)
    th = np.linalg.norm(r)
    if th < np.finfo(np.float64).eps:
        return np.eye(3)
    k = r / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * K


def camera_pose(x, y, z, yaw, pitch, roll, view_yaw, view_pitch) -> np.ndarray:
    """COORD pair -> 4x4 c2w float32.  utils/camera_poses.py:30-49 and :52-75."""
    d2r = np.pi / 180.0
    R = _rot("roll", roll * d2r) @ _rot("pitch", pitch * d2r) @ _rot("yaw", yaw * d2r)
    T = np.array([[1, 0, 0, x], [0, 1, 0, y], [0, 0, 1, z], [0, 0, 0, 1]], dtype=np.float32)
    ext = (R @ T).reshape(4, 4)
    hor = rodrigues([0, 0, view_yaw * d2r])
    ver = rodrigues([view_pitch * d2r, 0, 0])
    ext[:3, :3] = hor @ ver @ ext[:3, :3]
    return ext.astype(np.float32)


def synthetic_poses(n: int = 36, seed: int = 0) -> torch.Tensor:
    """The 36-pose GUI sweep of BASELINE config 5 at a random spot of the office_tokyo room
    (application/workspace.py:77-100; GUI step 30 deg, application/app.py:389-413)."""
    rng = np.random.RandomState(seed)
    x, z = rng.uniform(-2, 2), rng.uniform(-3, 1.5)
    poses = []
    for v in (-30, 0, 30):
        for h in range(0, 360, 30):
            poses.append(camera_pose(x, -0.5, z, 0.0, -90.0, 0.0, -float(h), float(v)))
    return torch.tensor(np.asarray(poses[:n], dtype=np.float32).reshape(-1, 4, 4))
