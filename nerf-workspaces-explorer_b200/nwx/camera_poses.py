"""COORD -> 4x4 camera-to-world poses: the boundary input producer of the hot path
(reference utils/camera_poses.py:30-75).  Host code, 16 floats per view; kept on the CPU like
the reference, with cv2.Rodrigues replaced by its closed form so no OpenCV is needed."""
from typing import List, Sequence

import numpy as np
import torch

from .data_descriptors import COORD

_D2R = np.pi / 180.0


def _axis_rotation(axis: int, theta: float) -> np.ndarray:
    c, s = np.cos(theta), np.sin(theta)
    m = np.eye(4, dtype=np.float32)
    i, j = [(1, 2), (0, 2), (0, 1)][axis]
    sign = -1.0 if axis == 1 else 1.0           # yaw (about y) has the transposed sign pattern
    m[i, i], m[j, j] = c, c
    m[i, j], m[j, i] = -sign * s, sign * s
    return m


def _rodrigues(rvec: Sequence[float]) -> np.ndarray:
    r = np.asarray(rvec, dtype=np.float64)
    theta = np.linalg.norm(r)
    if theta < np.finfo(np.float64).eps:
        return np.eye(3)
    k = r / theta
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(theta) * np.eye(3) + (1 - np.cos(theta)) * np.outer(k, k) + np.sin(theta) * K


def camera_to_world(c: COORD) -> np.ndarray:
    """R_roll @ R_pitch @ R_yaw @ T (camera_poses.py:30-49), float32."""
    R = _axis_rotation(2, c.roll * _D2R) @ _axis_rotation(0, c.pitch * _D2R) @ _axis_rotation(1, c.yaw * _D2R)
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = (c.x, c.y, c.z)
    return R @ T


def get_camera_poses_from_list_of_coordinates(init_coordinates: COORD, coordinates: List[COORD]) -> torch.Tensor:
    """One pose per view: the init pose with a local yaw (about z) then pitch (about x) applied to
    its rotation block (camera_poses.py:52-75).  Returns [B,4,4] float32 on the CPU."""
    poses = []
    for view in coordinates:
        ext = camera_to_world(init_coordinates).reshape(4, 4)
        hor = _rodrigues([0.0, 0.0, view.yaw * _D2R])
        ver = _rodrigues([view.pitch * _D2R, 0.0, 0.0])
        ext[:3, :3] = hor @ ver @ ext[:3, :3]
        poses.append(ext)
    return torch.tensor(np.asarray(poses, dtype=np.float32).reshape(-1, 4, 4))
