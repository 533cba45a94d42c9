"""The caller directly above the hot path: floor-plan click -> camera pose -> image
(reference application/workspace.py:13-196).  One table-driven class instead of four copies;
the reference's class names are kept as thin subclasses so `application/app.py:12-15` works as is."""
from __future__ import annotations

import math
import os
from typing import Dict, Mapping, Optional, Sequence, Tuple

import numpy as np

from .data_descriptors import COORD, HW
from .inference import NeRFReplicaInferenceHandler

# name: (floor_plan_scale, x' range (max, min), z' range (max, min), axis driving x' , angle_diff deg)
# x' is interpolated along rel_y for every room except New York (rel_x); workspace.py:77-196
_ROOMS: Dict[str, Tuple[HW, Tuple[float, float], Tuple[float, float], str, float]] = {
    "Office Tokyo": (HW(600, 600), (2.0, -2.0), (1.5, -3.0), "y", -10.0),
    "Office New York": (HW(600, 800), (1.8, -1.2), (2.0, -1.6), "x", 45.0),
    "Office Geneve": (HW(600, 1000), (1.7, -2.5), (4.2, -2.8), "y", 35.0),
    "Office Belgrade": (HW(600, 750), (4.7, -0.7), (3.5, -2.3), "y", -10.0),
}
FIXED_Y, INIT_PITCH = -0.5, -90.0


class Workspace:

    def __init__(self, name: str, ckpt_path: Optional[str] = None, config: Optional[Mapping] = None,
                 project_path: Optional[str] = None) -> None:
        if name not in _ROOMS:
            raise KeyError(f"unknown workspace {name!r}; known: {sorted(_ROOMS)}")
        self._name = name
        self._floor_plan_scale, self._x_range, self._z_range, self._x_axis, self._angle_diff = _ROOMS[name]
        self._office_name = name.replace(" ", "_").lower()
        root = project_path or os.getcwd()
        self._folder_path = os.path.normpath(os.path.join(root, "application", "workspaces", self._office_name))
        self._model_path = ckpt_path or os.path.normpath(
            os.path.join(root, "nerf", "final_models", self._office_name, "model.ckpt"))
        self._nerf_inference = NeRFReplicaInferenceHandler(self._office_name, self._model_path, config=config)

    def __repr__(self) -> str:
        return self._name

    name = property(lambda self: self._name)
    folder_path = property(lambda self: self._folder_path)
    floor_plan_scale = property(lambda self: self._floor_plan_scale)
    inference = property(lambda self: self._nerf_inference)

    def initialize_models(self) -> None:
        self._nerf_inference.initialize_models()

    def _transform_relative_coordinates(self, rel_x: float, rel_y: float, hor_angle: int, ver_angle: int
                                        ) -> Tuple[COORD, COORD]:
        """Relative click position on the floor plan -> (camera position COORD, local view COORD)."""
        rel_for_x, rel_for_z = (rel_y, rel_x) if self._x_axis == "y" else (rel_x, rel_y)
        x_prim = (self._x_range[1] - self._x_range[0]) * rel_for_x + self._x_range[0]
        z_prim = (self._z_range[1] - self._z_range[0]) * rel_for_z + self._z_range[0]
        c = np.cos(self._angle_diff / 180.0 * np.pi)
        return (COORD(x=x_prim / c, y=FIXED_Y, z=z_prim / c, yaw=0.0, pitch=INIT_PITCH, roll=0.0),
                COORD(x=0.0, y=0.0, z=0.0, yaw=-float(hor_angle), pitch=float(ver_angle), roll=0.0))

    def render_image(self, rel_x: float, rel_y: float, horizontal_angle: int, vertical_angle: int) -> np.ndarray:
        """workspace.py:54-68 -> uint8 [H,W,3]."""
        init, view = self._transform_relative_coordinates(rel_x, rel_y, horizontal_angle, vertical_angle)
        return self._nerf_inference.render_coordinates(init, view)

    def render_sweep(self, rel_x: float, rel_y: float, horizontal_angles: Sequence[int] = tuple(range(0, 360, 30)),
                     vertical_angles: Sequence[int] = (-30, 0, 30)) -> np.ndarray:
        """Every camera-button state of one clicked spot (application/app.py:389-413, 30 degree steps) in
        one batched launch sequence -> uint8 [len(v)*len(h), H, W, 3]."""
        init = self._transform_relative_coordinates(rel_x, rel_y, 0, 0)[0]
        views = [COORD(yaw=-float(h), pitch=float(v)) for v in vertical_angles for h in horizontal_angles]
        return self._nerf_inference.render_coordinates_batch(init, views)


def _named(room: str):
    class _W(Workspace):
        def __init__(self, ckpt_path: Optional[str] = None, config: Optional[Mapping] = None,
                     project_path: Optional[str] = None) -> None:
            super().__init__(room, ckpt_path, config, project_path)
    _W.__name__ = _W.__qualname__ = room.replace(" ", "") + "Workspace"
    return _W


OfficeTokyoWorkspace = _named("Office Tokyo")
OfficeNewYorkWorkspace = _named("Office New York")
OfficeGeneveWorkspace = _named("Office Geneve")
OfficeBelgradeWorkspace = _named("Office Belgrade")
