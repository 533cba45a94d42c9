"""Plain value types of the reference boundary (utils/data_descriptors.py:3-23)."""
from typing import NamedTuple


class HW(NamedTuple):
    h: int = 0
    w: int = 0


class XYZ(NamedTuple):
    x: float = 0.0
    y: float = 0.0
    z: float = 0.0


class COORD(NamedTuple):
    """Camera position (x, y, z) and Euler angles in degrees."""
    x: float = 0.0
    y: float = 0.0
    z: float = 0.0
    yaw: float = 0.0
    pitch: float = 0.0
    roll: float = 0.0
