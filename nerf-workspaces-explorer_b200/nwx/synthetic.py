"""Synthetic Replica-shaped workloads: the dataset and the trained
checkpoints are not available, so benchmarks use the GUI's 36-pose sweep at a random spot of the
office_tokyo room and random-init weights of the reference architecture."""
import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .camera_poses import get_camera_poses_from_list_of_coordinates
from .data_descriptors import COORD
from .models import NeRFModel


def intrinsics(height: int, width: int, hfov_deg: float = 90.0) -> Tuple[float, float, float, float]:
    """fx, fy, cx, cy as the handlers derive them (inference handler:67-74)."""
    fx = width / 2.0 / math.tan(math.radians(hfov_deg / 2.0))
    return fx, fx, (width - 1.0) / 2.0, (height - 1.0) / 2.0


def sweep_poses(n: int = 36, seed: int = 0) -> torch.Tensor:
    """yaw in {0,30,...,330} x pitch in {-30,0,30} around one spot (application/app.py:389-413;
    room bounds application/workspace.py:77-89) -> [n,4,4]."""
    rng = np.random.RandomState(seed)
    x, z = rng.uniform(-2, 2), rng.uniform(-3, 1.5)
    init = COORD(x=x, y=-0.5, z=z, yaw=0.0, pitch=-90.0, roll=0.0)
    views = [COORD(yaw=-float(h), pitch=float(v)) for v in (-30, 0, 30) for h in range(0, 360, 30)]
    return get_camera_poses_from_list_of_coordinates(init, views[:n])


def random_state_dicts(seed: int = 0, alpha_bias: Optional[float] = 0.1) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """Coarse and fine weights exactly as the reference handler would create them under
    torch.manual_seed(seed) (inference handler:106-119), with _alpha_linear.bias pinned so the
    density sign at the far sample is not a coin flip."""
    with torch.random.fork_rng(devices=[]):       # CPU generator only
        torch.manual_seed(seed)
        nets = [NeRFModel(8, 256, 63, 27, 5, use_view_dirs=True) for _ in range(2)]
    out = []
    for net in nets:
        if alpha_bias is not None:
            with torch.no_grad():
                net._alpha_linear.bias.fill_(alpha_bias)
        out.append({k: v.detach().clone() for k, v in net.state_dict().items()})
    return out[0], out[1]
