"""Reference entry points of nerf/rays/rays.py on the CUDA engine (same names, arguments, results)."""
from typing import Optional

import torch

from . import engine as _engine


def create_rays(num_images: int, Ts_c2w: torch.Tensor, height: int, width: int, fx: float, fy: float, cx: float,
                cy: float, near: float, far: float, use_view_dirs: bool = True,
                device: Optional[torch.device] = None) -> torch.Tensor:
    """rays.py:6-32 -> [B, H*W, 11|8] (o, d, near, far, d/|d|).  The reference builds this on
    the CPU and the handlers then call .cuda(); here it is produced on the GPU directly."""
    if Ts_c2w.shape[0] != num_images:
        raise ValueError(f"num_images={num_images} but {Ts_c2w.shape[0]} poses given")
    eng = _engine_for(device)
    rays = eng.raygen(Ts_c2w, height, width, fx, fy, cx, cy, near, far, use_view_dirs)
    return rays.view(num_images, height * width, 11 if use_view_dirs else 8)


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, N_samples: int, det: bool = False,
               u: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rays.py:74-121.  det=False draws u ~ U[0,1) on the device like rays.py:98 unless `u` is given."""
    if not det and u is None:
        u = torch.rand(list(bins.shape[:-1]) + [N_samples], device=bins.device)
    samples, _, _ = _engine.sample_pdf_bins(bins, weights, N_samples, None if det else u)
    return samples


_ENGINES = {}


def _engine_for(device: Optional[torch.device] = None) -> "_engine.Engine":
    """A lazily created per-device Engine for the context-free reference entry points."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = dev.index or 0
    if key not in _ENGINES:
        _ENGINES[key] = _engine.Engine(dev)
    return _ENGINES[key]
