"""Function-level drop-in: run the reference's OWN handler code on the CUDA engine.

The reference has no plugin layer; its hot path is reached through five imports at the top of
nerf/inference/nerf_replica_inference_handler.py (:11-17) and nerf/training/...handler.py (:12-19):

    from nerf.models.embedding import Embedding
    from nerf.models.model_utils import raw2outputs, run_network, to8b_np   (+ img2mse, mse2psnr)
    from nerf.models.nerf_model import NeRFModel
    from nerf.rays.rays import create_rays, sample_pdf
    from utils.batch_utils import batchify_rays

`patch_reference()` installs nwx's implementations under exactly those module names (sys.modules), so an
UNMODIFIED checkout of the reference -- application/workspace.py, the handlers' _volumetric_rendering --
imports the hand-written kernels instead of the torch-eager ops.  Nothing of the reference is copied or
edited; undo with `unpatch_reference()`.

`reference_volumetric_rendering` restates the handler's call sequence (inference handler:203-277) on top of
those entry points -- including the lambda around the fine network (:248) and the torch.sort(cat()) of :243 --
so that the function-level path can be tested and timed on machines where the reference itself is absent.
"""
from __future__ import annotations

import importlib
import sys
import types
from typing import Dict, Optional

import torch

_PATCHED: Dict[str, Optional[types.ModuleType]] = {}


def _module(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__nwx_patch__ = True
    return mod


def _ensure_package(name: str) -> None:
    """The parent package of a patched module: the reference's own package when it is importable
    (PYTHONPATH holds the reference root), otherwise an empty stand-in."""
    if name in sys.modules:
        return
    try:
        importlib.import_module(name)
    except Exception:  # noqa: BLE001
        pkg = types.ModuleType(name)
        pkg.__path__ = []          # a package, so that `import a.b` resolves through sys.modules
        pkg.__nwx_patch__ = True
        sys.modules[name] = pkg
        _PATCHED.setdefault(name, None)


def patch_reference() -> None:
    """Route the reference's hot-path imports to nwx.  Call before importing the reference's handlers."""
    from . import batch_utils, models, rays
    table = {
        "nerf.rays.rays": _module("nerf.rays.rays", create_rays=rays.create_rays, sample_pdf=rays.sample_pdf),
        "nerf.models.embedding": _module("nerf.models.embedding", Embedding=models.Embedding),
        "nerf.models.nerf_model": _module("nerf.models.nerf_model", NeRFModel=models.NeRFModel),
        "nerf.models.model_utils": _module("nerf.models.model_utils", run_network=models.run_network,
                                           raw2outputs=models.raw2outputs, img2mse=models.img2mse,
                                           mse2psnr=models.mse2psnr, to8b_np=models.to8b_np, to8b=models.to8b),
        "utils.batch_utils": _module("utils.batch_utils", batchify_rays=batch_utils.batchify_rays,
                                     batchify=batch_utils.batchify),
    }
    for name, mod in table.items():
        parts = name.split(".")
        for i in range(1, len(parts)):
            _ensure_package(".".join(parts[:i]))
        if name not in _PATCHED:
            _PATCHED[name] = sys.modules.get(name)
        sys.modules[name] = mod
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], mod)


def unpatch_reference() -> None:
    for name, old in list(_PATCHED.items()):
        if old is None:
            sys.modules.pop(name, None)
        else:
            sys.modules[name] = old
        del _PATCHED[name]


class ReferenceStyleRenderer:
    """The attributes _volumetric_rendering reads from its handler (inference handler:93-119), built from
    nwx's entry points exactly as `initialize_models` builds them."""

    def __init__(self, sd_coarse, sd_fine, device: torch.device, n_samples: int = 64, n_importance: int = 128,
                 net_chunk: int = 1024 * 32, white_bkgd: bool = False):
        from .engine import normalize_state_dict
        from .models import Embedding, NeRFModel
        self._n_samples, self._n_importance, self._net_chunk = n_samples, n_importance, net_chunk
        self._white_bkgd, self._endpoint_feat, self._perturb = white_bkgd, False, 0.0
        self._embed_fcn = Embedding(10, 10).embed
        self._embed_dirs_fcn = Embedding(4, 1).embed
        mk = lambda: NeRFModel(D=8, W=256, input_ch=63, output_ch=5, input_ch_views=27, use_view_dirs=True).to(device)
        self._nerf_net_coarse, self._nerf_net_fine = mk(), mk()
        self._nerf_net_coarse.load_state_dict(normalize_state_dict(sd_coarse))
        self._nerf_net_fine.load_state_dict(normalize_state_dict(sd_fine))
        self._nerf_net_coarse.eval(); self._nerf_net_fine.eval()

    @torch.no_grad()
    def _volumetric_rendering(self, ray_batch: torch.Tensor) -> Dict[str, torch.Tensor]:
        return reference_volumetric_rendering(self, ray_batch)


def reference_volumetric_rendering(self, ray_batch: torch.Tensor) -> Dict[str, torch.Tensor]:
    """The reference's call sequence (inference handler:203-277) on nwx's function-level entry points:
    run_network (coarse: the module itself; fine: the handler's lambda), raw2outputs, sample_pdf, and
    torch for the glue the handler does itself (point construction, cat + sort, std)."""
    from .models import raw2outputs, run_network
    from .rays import sample_pdf
    N_rays = ray_batch.shape[0]
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 8 else None
    bounds = torch.reshape(ray_batch[..., 6:8], [-1, 1, 2])
    near, far = bounds[..., 0], bounds[..., 1]
    t_vals = torch.linspace(0., 1., steps=self._n_samples).to(ray_batch.device)            # :216
    z_vals = near * (1. - t_vals) + far * t_vals                                           # :218
    z_vals = z_vals.expand([N_rays, self._n_samples])
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]                # :223
    raw_coarse = run_network(pts, viewdirs, self._nerf_net_coarse, self._embed_fcn, self._embed_dirs_fcn,
                             netchunk=self._net_chunk)                                     # :226
    rgb_c, disp_c, acc_c, weights_c, depth_c, _ = raw2outputs(raw_coarse, z_vals, rays_d, 0, self._white_bkgd,
                                                              endpoint_feat=False)         # :229
    z_mid = .5 * (z_vals[..., 1:] + z_vals[..., :-1])                                       # :236
    z_samples = sample_pdf(z_mid, weights_c[..., 1:-1], self._n_importance,
                           det=(self._perturb == 0.) or True).detach()                     # :237-239
    z_vals, _ = torch.sort(torch.cat([z_vals, z_samples], -1), -1)                          # :243
    pts_f = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]              # :246
    raw_fine = run_network(pts_f, viewdirs, lambda x: self._nerf_net_fine(x, self._endpoint_feat),
                           self._embed_fcn, self._embed_dirs_fcn, netchunk=self._net_chunk)   # :248
    rgb_f, disp_f, acc_f, _, depth_f, _ = raw2outputs(raw_fine, z_vals, rays_d, 0, self._white_bkgd,
                                                      endpoint_feat=self._endpoint_feat)   # :251
    return {"rgb_coarse": rgb_c, "disp_coarse": disp_c, "acc_coarse": acc_c, "depth_coarse": depth_c,
            "raw_coarse": raw_coarse, "rgb_fine": rgb_f, "disp_fine": disp_f, "acc_fine": acc_f,
            "depth_fine": depth_f, "z_std": torch.std(z_samples, dim=-1, unbiased=False), "raw_fine": raw_fine}
