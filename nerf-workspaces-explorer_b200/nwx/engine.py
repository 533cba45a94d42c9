"""Tensor-level front end of libnwx: torch owns device memory and streams, every op is a
hand-written sm_100a kernel behind the C ABI (include/nwx.h).  No torch compute on the path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Mapping, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import RENDER_OUT_FIELDS, RenderOpts, RenderOut, check

COARSE, FINE = 0, 1

# NeRFModel.state_dict() order (reference nerf/models/nerf_model.py:32-41)
STATE_KEYS: Tuple[str, ...] = tuple(
    [f"_pts_linears.{i}.{p}" for i in range(8) for p in ("weight", "bias")]
    + [f"_views_linears.0.{p}" for p in ("weight", "bias")]
    + [f"_{n}_linear.{p}" for n in ("feature", "alpha", "rgb") for p in ("weight", "bias")])
STATE_SHAPES = dict(
    [(f"_pts_linears.{i}.weight", (256, 63 if i == 0 else (319 if i == 5 else 256))) for i in range(8)]
    + [(f"_pts_linears.{i}.bias", (256,)) for i in range(8)]
    + [("_views_linears.0.weight", (128, 283)), ("_views_linears.0.bias", (128,)),
       ("_feature_linear.weight", (256, 256)), ("_feature_linear.bias", (256,)),
       ("_alpha_linear.weight", (1, 256)), ("_alpha_linear.bias", (1,)),
       ("_rgb_linear.weight", (3, 128)), ("_rgb_linear.bias", (3,))])

REFERENCE_KEYS = ("rgb_coarse", "disp_coarse", "acc_coarse", "depth_coarse", "raw_coarse", "rgb_fine",
                  "disp_fine", "acc_fine", "depth_fine", "z_std", "raw_fine")
"""The 11 keys of the handlers' output dict (inference handler:256-268)."""

_OUT_SHAPES = {
    "rgb_coarse": lambda n, sc, ni: (n, 3), "disp_coarse": lambda n, sc, ni: (n,),
    "acc_coarse": lambda n, sc, ni: (n,), "depth_coarse": lambda n, sc, ni: (n,),
    "raw_coarse": lambda n, sc, ni: (n, sc, 4), "rgb_fine": lambda n, sc, ni: (n, 3),
    "disp_fine": lambda n, sc, ni: (n,), "acc_fine": lambda n, sc, ni: (n,),
    "depth_fine": lambda n, sc, ni: (n,), "raw_fine": lambda n, sc, ni: (n, sc + ni, 4),
    "z_std": lambda n, sc, ni: (n,), "z_vals_coarse": lambda n, sc, ni: (n, sc),
    "weights_coarse": lambda n, sc, ni: (n, sc), "z_samples": lambda n, sc, ni: (n, ni),
    "z_vals_fine": lambda n, sc, ni: (n, sc + ni), "weights_fine": lambda n, sc, ni: (n, sc + ni),
    "inds": lambda n, sc, ni: (n, ni), "rgb8_fine": lambda n, sc, ni: (n, 3),
}
_OUT_DTYPES = {"inds": torch.int64, "rgb8_fine": torch.uint8}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.NwxError(f"{what}: expected a CUDA tensor (the engine has no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_LINSPACE: Dict[Tuple[int, int], torch.Tensor] = {}


def linspace01(n: int, device: torch.device) -> torch.Tensor:
    """torch.linspace(0, 1, n) computed by torch-CPU (bit-identical to the reference's vector;
    inference handler:216, rays.py:95) and cached on the device."""
    key = (n, device.index or 0)
    if key not in _LINSPACE:
        _LINSPACE[key] = torch.linspace(0., 1., steps=n).to(device)
    return _LINSPACE[key]


def normalize_state_dict(sd: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Accept both key styles of the reference: module attrs are `_pts_linears...`, shipped
    checkpoints carry `pts_linears...` (transform_state_dict, inference handler:150-164)."""
    out = {}
    for k, v in sd.items():
        nk = k if k.startswith("_") else "_" + k
        out[nk] = v
    missing = [k for k in STATE_KEYS if k not in out]
    if missing:
        raise _lib.NwxError(f"state_dict is missing {missing[:3]}... (expected NeRFModel(8,256,63,27,skips=(4,),"
                            " use_view_dirs=True) keys)")
    for k in STATE_KEYS:
        if tuple(out[k].shape) != STATE_SHAPES[k]:
            raise _lib.NwxError(f"{k}: shape {tuple(out[k].shape)} != {STATE_SHAPES[k]}; the fused kernel "
                                "implements the one architecture the reference instantiates")
    return out


class RngOptions:
    """In-kernel random draws (counter-based Philox keyed by seed/offset) for the training-mode
    render: stratified jitter, random importance uniforms, sigma noise N(0,1)*noise_std."""

    def __init__(self, seed: int = 0, offset: int = 0, jitter: bool = False, random_u: bool = False,
                 noise_std: float = 0.0):
        self.seed, self.offset, self.jitter, self.random_u, self.noise_std = seed, offset, jitter, random_u, noise_std

    def fields(self):
        return (self.seed & 0xFFFFFFFFFFFFFFFF, self.offset & 0xFFFFFFFFFFFFFFFF, float(self.noise_std),
                int(self.jitter), int(self.random_u))


def rng_fill(kind: str, seed: int, offset: int, stream: int, n: int, scale: float = 1.0,
             device: Optional[torch.device] = None) -> torch.Tensor:
    """The library's draws as a tensor: kind 'uniform' | 'normal'; stream 0 jitter, 1 u, 2/3 noise."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty((n,), device=dev)
    check(_lib.lib().nwx_rng_fill({"uniform": 0, "normal": 1}[kind], seed, offset, stream, float(scale), n,
                                  out.data_ptr(), _stream()), "nwx_rng_fill")
    return out


class Engine:
    """One nwx_ctx: bf16-packed coarse+fine weights and scratch on one GPU."""

    def __init__(self, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise _lib.NwxError("no CUDA device: the engine is sm_100a-only and has no fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._lib = _lib.lib()
        handle = C.c_void_p()
        check(self._lib.nwx_ctx_create(self.device.index or 0, C.byref(handle)), "nwx_ctx_create")
        self._ctx = handle
        self._keep = {}
        self.weights_version = 0          # bumped by load_weights: captured graphs bake the biases in
        self.profiling = False

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.nwx_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights -------------------------------------------------------------------------
    def load_weights(self, which: int, state_dict: Mapping[str, torch.Tensor]) -> None:
        sd = normalize_state_dict(state_dict)
        tensors = [_f32(sd[k].detach().to(self.device), k) for k in STATE_KEYS]
        arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        with torch.cuda.device(self.device):
            check(self._lib.nwx_load_weights(self._ctx, which, arr, _stream()), "nwx_load_weights")
        self._keep[which] = tensors
        self.weights_version += 1

    def set_mlp_variant(self, variant: int) -> None:
        check(self._lib.nwx_set_mlp_variant(self._ctx, variant), "nwx_set_mlp_variant")

    def debug_tap(self, layer: int, out: Optional[torch.Tensor]) -> None:
        check(self._lib.nwx_debug_tap(self._ctx, layer, _ptr(out)), "nwx_debug_tap")

    def debug_diag(self, pinned: Optional[torch.Tensor]) -> None:
        check(self._lib.nwx_debug_diag(self._ctx, _ptr(pinned)), "nwx_debug_diag")

    def last_diag(self) -> Tuple[int, int, int, int]:
        """(0xDEADxxxx | waiter, block, barrier, parity) of the last barrier wait that timed out, zeros if none.
        Lives in mapped host memory, so it is readable after a trap has poisoned the CUDA context."""
        buf = (C.c_uint32 * 4)()
        check(self._lib.nwx_ctx_last_diag(self._ctx, buf), "nwx_ctx_last_diag")
        return tuple(int(v) for v in buf)

    STAGES = ("coarse_z", "dirbias_coarse", "mlp_coarse", "composite_coarse", "sample_pdf", "dirbias_fine",
              "mlp_fine", "composite_fine")

    def set_profiling(self, on: bool) -> None:
        check(self._lib.nwx_ctx_set_profiling(self._ctx, int(on)), "nwx_ctx_set_profiling")
        self.profiling = bool(on)

    def scratch_state(self) -> Tuple[int, int]:
        """(bytes, generation) of the context's scratch; the generation changes when it is re-allocated."""
        b, g = C.c_int64(), C.c_int64()
        check(self._lib.nwx_ctx_scratch_state(self._ctx, C.byref(b), C.byref(g)), "nwx_ctx_scratch_state")
        return int(b.value), int(g.value)

    def stage_ms(self) -> Dict[str, float]:
        """Device time of each stage of the last render_rays call (waits for it)."""
        buf = (C.c_float * len(self.STAGES))()
        check(self._lib.nwx_ctx_stage_ms(self._ctx, buf), "nwx_ctx_stage_ms")
        return dict(zip(self.STAGES, [float(v) for v in buf]))

    def reserve(self, max_rays: int, n_samples: int = 64, n_importance: int = 128) -> None:
        check(self._lib.nwx_ctx_reserve(self._ctx, max_rays, n_samples, n_importance), "nwx_ctx_reserve")

    # ---- K1 ------------------------------------------------------------------------------
    def raygen(self, c2w: torch.Tensor, H: int, W: int, fx: float, fy: float, cx: float, cy: float,
               near: float, far: float, use_view_dirs: bool = True, ray0: int = 0,
               nrays: Optional[int] = None) -> torch.Tensor:
        c2w = _f32(c2w.to(self.device), "c2w").reshape(-1, 16)
        B = c2w.shape[0]
        nrays = B * H * W - ray0 if nrays is None else nrays
        out = torch.empty((nrays, 11 if use_view_dirs else 8), device=self.device, dtype=torch.float32)
        check(self._lib.nwx_raygen(c2w.data_ptr(), B, H, W, fx, fy, cx, cy, near, far, int(use_view_dirs),
                                   ray0, nrays, out.data_ptr(), _stream()), "nwx_raygen")
        return out

    def coarse_z(self, rays: torch.Tensor, n_samples: int, t_rand: Optional[torch.Tensor] = None) -> torch.Tensor:
        rays = _f32(rays, "rays")
        t_rand = None if t_rand is None else _f32(t_rand, "t_rand")
        z = torch.empty((rays.shape[0], n_samples), device=rays.device, dtype=torch.float32)
        check(self._lib.nwx_coarse_z(rays.data_ptr(), rays.shape[1], rays.shape[0], n_samples,
                                     linspace01(n_samples, rays.device).data_ptr(), _ptr(t_rand), z.data_ptr(),
                                     _stream()), "nwx_coarse_z")
        return z

    # ---- K3 ------------------------------------------------------------------------------
    def mlp_forward(self, which: int, rays: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        rays, z = _f32(rays, "rays"), _f32(z, "z")
        N, S = z.shape
        raw = torch.empty((N, S, 4), device=z.device, dtype=torch.float32)
        check(self._lib.nwx_mlp_forward(self._ctx, which, rays.data_ptr(), rays.shape[1], z.data_ptr(), N, S,
                                        raw.data_ptr(), _stream()), "nwx_mlp_forward")
        return raw

    def mlp_forward_points(self, which: int, pts: torch.Tensor, dirs: torch.Tensor,
                           pts_per_dir: int = 1) -> torch.Tensor:
        pts, dirs = _f32(pts, "pts").reshape(-1, 3), _f32(dirs, "dirs").reshape(-1, 3)
        P = pts.shape[0]
        if dirs.shape[0] * pts_per_dir < P:
            raise _lib.NwxError("mlp_forward_points: not enough view directions")
        raw = torch.empty((P, 4), device=pts.device, dtype=torch.float32)
        check(self._lib.nwx_mlp_forward_points(self._ctx, which, pts.data_ptr(), dirs.data_ptr(), P, pts_per_dir,
                                               raw.data_ptr(), _stream()), "nwx_mlp_forward_points")
        return raw

    def mlp_forward_embedded(self, which: int, x: torch.Tensor) -> torch.Tensor:
        x = _f32(x, "x")
        if x.dim() != 2 or x.shape[1] != 90:
            raise _lib.NwxError(f"embedded input must be [P,90], got {tuple(x.shape)}")
        raw = torch.empty((x.shape[0], 4), device=x.device, dtype=torch.float32)
        check(self._lib.nwx_mlp_forward_embedded(self._ctx, which, x.data_ptr(), x.shape[0], raw.data_ptr(),
                                                 _stream()), "nwx_mlp_forward_embedded")
        return raw

    # ---- training batch sampling ------------------------------------------------------------
    def sample_training_batch(self, rays_bank: torch.Tensor, rgb_bank: torch.Tensor, n: int, seed: int, offset: int,
                              want_indices: bool = False):
        """_sample_training_data (training handler:341-370) without leaving the device: one random image of
        rays_bank [num_img, num_ray, 11], n random pixels of it with replacement, and the gather of the rays and
        of the ground-truth pixels rgb_bank [num_img, num_ray, 3] -> (rays [n,11], gt [n,3][, idx int64 [1+n]])."""
        if not (rays_bank.is_cuda and rgb_bank.is_cuda and rays_bank.is_contiguous() and rgb_bank.is_contiguous()
                and rays_bank.dtype == torch.float32 and rgb_bank.dtype == torch.float32):
            raise _lib.NwxError("sample_training_batch: banks must be contiguous fp32 CUDA tensors")
        num_img, num_ray, ray_dim = rays_bank.shape
        if tuple(rgb_bank.shape) != (num_img, num_ray, 3):
            raise _lib.NwxError(f"rgb bank must be {(num_img, num_ray, 3)}, got {tuple(rgb_bank.shape)}")
        dev = rays_bank.device
        rays = torch.empty((n, ray_dim), device=dev)
        gt = torch.empty((n, 3), device=dev)
        idx = torch.empty((1 + n,), device=dev, dtype=torch.int64) if want_indices else None
        check(self._lib.nwx_sample_training_batch(rays_bank.data_ptr(), rgb_bank.data_ptr(), num_img, num_ray, ray_dim, n,
                                                  seed & 0xFFFFFFFFFFFFFFFF, offset & 0xFFFFFFFFFFFFFFFF, rays.data_ptr(),
                                                  gt.data_ptr(), _ptr(idx), _stream()), "nwx_sample_training_batch")
        return (rays, gt, idx) if want_indices else (rays, gt)

    # ---- whole chunk ---------------------------------------------------------------------
    def render_rays(self, rays: torch.Tensor, n_samples: int = 64, n_importance: int = 128,
                    white_bkgd: bool = False, want: Iterable[str] = REFERENCE_KEYS,
                    t_rand: Optional[torch.Tensor] = None, u: Optional[torch.Tensor] = None,
                    noise_coarse: Optional[torch.Tensor] = None, noise_fine: Optional[torch.Tensor] = None,
                    out: Optional[Dict[str, torch.Tensor]] = None, rng: Optional["RngOptions"] = None
                    ) -> Dict[str, torch.Tensor]:
        """The body of _volumetric_rendering for all rays at once (inference handler:203-277;
        with t_rand/u/noise_*: training handler:534-618).  `want` selects output tensors; only
        those are written to HBM.  Adds "flags" (int32: bit0 NaN, bit1 Inf).  `rng` switches on the
        in-kernel counter-based draws for whichever of t_rand / u / noise_* is not given."""
        rays = _f32(rays, "rays")
        N, dev = rays.shape[0], rays.device
        want = set(want)
        res = dict(out) if out else {}
        if not ({"rgb_fine", "rgb8_fine"} & (want | set(res))):
            want.add("rgb_fine")                  # the library needs one of the two final images
        unknown = [k for k in set(want) | set(res) if k not in _OUT_SHAPES]
        if unknown:
            raise _lib.NwxError(f"render_rays: unknown output(s) {sorted(unknown)}; known: {sorted(_OUT_SHAPES)}")
        for k, t in res.items():                  # caller-provided outputs are written in place: validate them
            shape, dtype = _OUT_SHAPES[k](N, n_samples, n_importance), _OUT_DTYPES.get(k, torch.float32)
            if not (t.is_cuda and t.device == dev and t.is_contiguous() and t.dtype == dtype and tuple(t.shape) == shape):
                raise _lib.NwxError(f"render_rays: out[{k!r}] must be a contiguous {dtype} CUDA tensor of shape {shape} "
                                    f"on {dev}, got {t.dtype} {tuple(t.shape)} on {t.device}")
        for k in want:
            if k not in res:
                res[k] = torch.empty(_OUT_SHAPES[k](N, n_samples, n_importance), device=dev,
                                     dtype=_OUT_DTYPES.get(k, torch.float32))
        res["flags"] = torch.zeros(1, device=dev, dtype=torch.int32)
        keep = [None if t is None else _f32(t, "rand") for t in (t_rand, u, noise_coarse, noise_fine)]
        rng = rng or RngOptions()
        opts = RenderOpts(n_samples, n_importance, int(white_bkgd), rays.shape[1],
                          linspace01(n_samples, dev).data_ptr(), linspace01(n_importance, dev).data_ptr(),
                          _ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2]), _ptr(keep[3]), *rng.fields())
        outs = RenderOut(*[_ptr(res.get(name)) for name in RENDER_OUT_FIELDS])
        check(self._lib.nwx_render_rays(self._ctx, rays.data_ptr(), N, C.byref(opts), C.byref(outs), _stream()),
              "nwx_render_rays")
        return res


# ---- context-free kernels (K2, K4, embed) ---------------------------------------------------


def composite(raw: torch.Tensor, z_vals: torch.Tensor, rays_d: torch.Tensor, noise: Optional[torch.Tensor] = None,
              white_bkgd: bool = False, want_weights: bool = True, flags: Optional[torch.Tensor] = None):
    """raw2outputs (model_utils.py:33-100) -> (rgb, disp, acc, weights, depth)."""
    raw, z_vals, rays_d = _f32(raw, "raw"), _f32(z_vals, "z_vals"), _f32(rays_d, "rays_d")
    N, S = z_vals.shape
    dev = raw.device
    rgb = torch.empty((N, 3), device=dev)
    disp, acc, depth = (torch.empty((N,), device=dev) for _ in range(3))
    weights = torch.empty((N, S), device=dev) if want_weights else None
    noise = None if noise is None else _f32(noise, "noise")
    check(_lib.lib().nwx_composite_fwd(raw.data_ptr(), z_vals.data_ptr(), rays_d.data_ptr(), rays_d.shape[-1],
                                       _ptr(noise), N, S, int(white_bkgd), rgb.data_ptr(), disp.data_ptr(),
                                       acc.data_ptr(), depth.data_ptr(), _ptr(weights), _ptr(flags), _stream()),
          "nwx_composite_fwd")
    return rgb, disp, acc, weights, depth


def composite_backward(raw, z_vals, rays_d, d_rgb, noise=None, white_bkgd=False) -> torch.Tensor:
    raw, z_vals, rays_d, d_rgb = (_f32(t, "arg") for t in (raw, z_vals, rays_d, d_rgb))
    N, S = z_vals.shape
    d_raw = torch.empty((N, S, 4), device=raw.device)
    noise = None if noise is None else _f32(noise, "noise")
    check(_lib.lib().nwx_composite_bwd(raw.data_ptr(), z_vals.data_ptr(), rays_d.data_ptr(), rays_d.shape[-1],
                                       _ptr(noise), None, d_rgb.data_ptr(), N, S, int(white_bkgd),
                                       d_raw.data_ptr(), _stream()), "nwx_composite_bwd")
    return d_raw


def sample_pdf_merge(z_c: torch.Tensor, w_c: torch.Tensor, n_importance: int, u: Optional[torch.Tensor] = None,
                     want_inds: bool = True):
    """sample_pdf(mid(z_c), w_c[:,1:-1]) + sort-merge -> (z_samples, z_fine, inds, z_std)."""
    z_c, w_c = _f32(z_c, "z_c"), _f32(w_c, "w_c")
    N, Sc = z_c.shape
    dev = z_c.device
    u = None if u is None else _f32(u, "u")
    z_s = torch.empty((N, n_importance), device=dev)
    z_f = torch.empty((N, Sc + n_importance), device=dev)
    inds = torch.empty((N, n_importance), device=dev, dtype=torch.int64) if want_inds else None
    z_std = torch.empty((N,), device=dev)
    check(_lib.lib().nwx_sample_pdf(z_c.data_ptr(), w_c.data_ptr(), Sc, _ptr(u),
                                    linspace01(n_importance, dev).data_ptr(), n_importance, N, z_s.data_ptr(),
                                    z_f.data_ptr(), _ptr(inds), z_std.data_ptr(), _stream()), "nwx_sample_pdf")
    return z_s, z_f, inds, z_std


def sample_pdf_bins(bins: torch.Tensor, weights: torch.Tensor, n_samples: int, u: Optional[torch.Tensor] = None,
                    want_inds: bool = False, want_cdf: bool = False):
    """The literal sample_pdf(bins, weights, N_samples, det) of rays.py:74."""
    bins, weights = _f32(bins, "bins"), _f32(weights, "weights")
    N, M = bins.shape
    if weights.shape != (N, M - 1):
        raise _lib.NwxError(f"weights must be [N, M-1] = {(N, M - 1)}, got {tuple(weights.shape)}")
    dev = bins.device
    u = None if u is None else _f32(u, "u")
    smp = torch.empty((N, n_samples), device=dev)
    inds = torch.empty((N, n_samples), device=dev, dtype=torch.int64) if want_inds else None
    cdf = torch.empty((N, M), device=dev) if want_cdf else None
    check(_lib.lib().nwx_sample_pdf_bins(bins.data_ptr(), weights.data_ptr(), M, _ptr(u),
                                         linspace01(n_samples, dev).data_ptr(), n_samples, N, smp.data_ptr(),
                                         _ptr(inds), _ptr(cdf), _stream()), "nwx_sample_pdf_bins")
    return smp, inds, cdf


def embed(x: torch.Tensor, num_freqs: int, scalar_factor: float) -> torch.Tensor:
    x = _f32(x, "x")
    lead = x.shape[:-1]
    flat = x.reshape(-1, 3)
    out = torch.empty((flat.shape[0], 3 + 6 * num_freqs), device=x.device)
    check(_lib.lib().nwx_embed(flat.data_ptr(), flat.shape[0], num_freqs, float(scalar_factor), out.data_ptr(),
                               _stream()), "nwx_embed")
    return out.reshape(*lead, 3 + 6 * num_freqs)


def to8b(x: torch.Tensor) -> torch.Tensor:
    x = _f32(x, "x")
    out = torch.empty(x.shape, device=x.device, dtype=torch.uint8)
    check(_lib.lib().nwx_to8b(x.data_ptr(), x.numel(), out.data_ptr(), _stream()), "nwx_to8b")
    return out


_GRAPH_LAUNCHES = 0        # kernels launched by CUDA-graph replays (the library only sees the capture)


def count_graph_launches(n: int) -> None:
    global _GRAPH_LAUNCHES
    _GRAPH_LAUNCHES += int(n)


def launch_count() -> int:
    """Kernels of libnwx launched so far by this process: direct launches (counted by the library) plus the
    kernel nodes of every CUDA-graph replay (counted here)."""
    return int(_lib.lib().nwx_launch_count()) + _GRAPH_LAUNCHES
