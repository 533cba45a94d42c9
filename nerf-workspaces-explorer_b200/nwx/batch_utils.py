"""Chunk loops of the reference (utils/batch_utils.py:7-39), kept for callers that use them.
The engine itself renders a whole chunk in one launch sequence and never needs them."""
from typing import Callable, Dict, Optional

import torch


def batchify_rays(render_fn: Callable, rays_flat: torch.Tensor, chunk: int = 1024 * 32) -> Dict[str, torch.Tensor]:
    parts: Dict[str, list] = {}
    for start in range(0, rays_flat.shape[0], chunk):
        for key, val in render_fn(rays_flat[start:start + chunk]).items():
            parts.setdefault(key, []).append(val)
    return {key: torch.cat(vals, 0) for key, vals in parts.items()}


def batchify(fn: Callable, chunk: Optional[int]) -> Callable:
    if chunk is None:
        return fn
    return lambda inputs: torch.cat([fn(inputs[s:s + chunk]) for s in range(0, inputs.shape[0], chunk)], 0)
