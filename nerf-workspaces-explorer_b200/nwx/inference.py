"""NeRFReplicaInferenceHandler on the CUDA engine: same constructor, methods and results as
reference nerf/inference/nerf_replica_inference_handler.py, so application/workspace.py
(`Workspace.render_image`, :54-68) can use it unchanged.

What changed underneath: a frame is ONE launch sequence of hand-written kernels
(raygen -> coarse z -> fused PE+MLP -> composite -> sample_pdf+merge -> fused PE+MLP -> composite
-> uint8) instead of 38 Python ray chunks x 64 network chunks with 22 host syncs each."""
from __future__ import annotations

import copy
import math
from typing import Any, Dict, List, Mapping, Optional, Sequence

import numpy as np
import torch

from . import engine as _engine
from .camera_poses import get_camera_poses_from_list_of_coordinates
from .config import default_config, number
from .data_descriptors import COORD
from .models import Embedding, NeRFModel


class _FrameGraph:
    """The launch sequence of one fixed ray range (raygen + the 8-launch render, uint8 out) captured ONCE into a
    CUDA graph and replayed: one graph launch per frame instead of a Python -> ctypes -> 9 x cudaLaunch walk, which
    is what the host adds to a frame once the GPU is drained by the read-back (0.2-0.3 ms, all of it exposed)."""

    def __init__(self, handler: "NeRFReplicaInferenceHandler", B: int, ray0: int, count: int):
        eng, dev = handler.engine, handler._device
        self.pose = torch.zeros((B, 4, 4), device=dev, dtype=torch.float32)      # static input
        self.rgb8 = torch.empty((count, 3), device=dev, dtype=torch.uint8)       # static output
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):                   # sizes the scratch, sets the kernels' attributes
            handler._render_rays_u8_eager(self.pose, ray0, count, self.rgb8)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        before = _engine.launch_count()
        with torch.cuda.graph(self.graph):
            handler._render_rays_u8_eager(self.pose, ray0, count, self.rgb8)
        self.kernels = _engine.launch_count() - before
        self.stamp = (eng.weights_version, eng.scratch_state()[1])

    def run(self, c2w: torch.Tensor) -> torch.Tensor:
        self.pose.copy_(c2w, non_blocking=True)          # host (pinned: asynchronous) or device poses
        self.graph.replay()
        _engine.count_graph_launches(self.kernels)
        return self.rgb8


class NeRFReplicaInferenceHandler:

    def __init__(self, office_name: str, ckpt_path: Optional[str], config: Optional[Mapping] = None,
                 device: Optional[torch.device] = None) -> None:
        self._office_name = office_name
        self._ckpt_path = ckpt_path
        cfg = default_config() if config is None else config
        self._config = cfg
        exp, mdl, rnd = cfg["experiment"], cfg["model"], cfg["rendering"]
        self._endpoint_feat = bool(exp.get("endpoint_feat", False))
        self._net_chunk = number(mdl["net_chunk"])
        self._chunk = number(cfg["inference"]["chunk"])       # kept for API parity; see _render_rays
        self._n_rays = number(rnd["n_rays"])
        self._n_samples = int(rnd["n_samples"])
        self._n_importance = int(rnd["n_importance"])
        self._num_freqs_3d = int(rnd["num_freqs_3d"])
        self._num_freqs_2d = int(rnd["num_freqs_2d"])
        self._use_view_dirs = bool(rnd["use_view_dirs"])
        self._raw_noise_std = float(rnd["raw_noise_std"])
        self._white_bkgd = bool(rnd["white_background"])
        self._perturb = float(rnd["perturb"])
        self._img_h, self._img_w = int(exp["image_height"]), int(exp["image_width"])
        self._n_pix = self._img_h * self._img_w
        self._hfov = 90
        self._fx = self._img_w / 2.0 / math.tan(math.radians(self._hfov / 2.0))     # handler:71
        self._fy = self._fx
        self._cx = (self._img_w - 1.0) / 2.0
        self._cy = (self._img_h - 1.0) / 2.0
        self._depth_close_bound, self._depth_far_bound = rnd["depth_range"]
        if (self._num_freqs_3d, self._num_freqs_2d, self._use_view_dirs, self._endpoint_feat) != (10, 4, True, False) \
                or number(mdl["net_depth"]) != 8 or number(mdl["net_width"]) != 256:
            raise _engine._lib.NwxError("only the shipped configuration (8x256, 10/4 frequencies, view dirs, "
                                        "no endpoint features) is implemented by the fused kernels")
        self._device_arg = device              # resolved lazily so construction needs no GPU
        self._engine: Optional[_engine.Engine] = None
        self._nerf_net_coarse: Optional[NeRFModel] = None
        self._nerf_net_fine: Optional[NeRFModel] = None
        self._embed_fcn = self._embed_dirs_fcn = None
        self.max_rays_per_launch = 1 << 20       # scratch is ~6.4 KB per ray
        self._host_frames: Optional[torch.Tensor] = None     # persistent pinned read-back buffer (uint8)
        self.use_cuda_graphs = True              # replay a captured launch sequence for repeated frame shapes
        self._graphs: Dict[Any, _FrameGraph] = {}

    @property
    def _device(self) -> torch.device:
        if self._device_arg is None:
            self._device_arg = torch.device("cuda", torch.cuda.current_device())
        return torch.device(self._device_arg)

    # ---- model set-up (handler:88-164) -----------------------------------------------------
    def _build_models(self) -> None:
        self._embed_fcn = Embedding(self._num_freqs_3d, 10).embed
        self._embed_dirs_fcn = Embedding(self._num_freqs_2d, 1).embed
        mk = lambda: NeRFModel(D=8, W=256, input_ch=63, output_ch=5, input_ch_views=27,
                               use_view_dirs=True).to(self._device)
        self._nerf_net_coarse, self._nerf_net_fine = mk(), mk()
        self._nerf_net_coarse.eval(); self._nerf_net_fine.eval()

    def initialize_models(self) -> None:
        """Create both networks and load the checkpoint {network_coarse_state_dict,
        network_fine_state_dict} (training handler:404-407).  Missing file -> RuntimeError, as in
        the reference (handler:147-148)."""
        try:
            ckpt = torch.load(self._ckpt_path, map_location="cpu")
        except (FileNotFoundError, TypeError, AttributeError) as exc:
            raise RuntimeError(f"Checkpoint path: {self._ckpt_path} for model cannot be found!") from exc
        self._build_models()
        self.load_state_dicts(self.transform_state_dict(ckpt["network_coarse_state_dict"]),
                              self.transform_state_dict(ckpt["network_fine_state_dict"]))

    def load_state_dicts(self, coarse: Mapping[str, torch.Tensor], fine: Mapping[str, torch.Tensor]) -> None:
        """Install weights from state dicts (either key style) -- what initialize_models does after
        torch.load; also the entry for synthetic / in-memory weights."""
        if self._nerf_net_coarse is None:
            self._build_models()
        self._nerf_net_coarse.load_state_dict(_engine.normalize_state_dict(coarse))
        self._nerf_net_fine.load_state_dict(_engine.normalize_state_dict(fine))
        self._engine = _engine.Engine(self._device)
        self._engine.load_weights(_engine.COARSE, self._nerf_net_coarse.state_dict())
        self._engine.load_weights(_engine.FINE, self._nerf_net_fine.state_dict())
        self._graphs.clear()                     # captured sequences carry the old biases in their kernel parameters
        # size the scratch for one frame now, so that no render call allocates (or frees) device memory
        self._engine.reserve(max(1, min(self.max_rays_per_launch, self._n_pix)), self._n_samples, self._n_importance)

    @staticmethod
    def transform_state_dict(state_dict: Dict[str, Any]) -> Dict[str, Any]:
        """handler:150-164: shipped checkpoints lack the leading underscore of the module attrs."""
        out = {}
        for key, val in state_dict.items():
            named = key.endswith("weight") or key.endswith("bias")
            out[f"_{key}" if named and not key.startswith("_") else key] = val
        return copy.deepcopy(out)

    @property
    def engine(self) -> _engine.Engine:
        if self._engine is None:
            raise RuntimeError("models are not initialised: call initialize_models() or load_state_dicts()")
        return self._engine

    # ---- rendering -------------------------------------------------------------------------
    def render_coordinates(self, init_coordinates: COORD, coordinates: COORD) -> np.ndarray:
        """handler:166-185 -> uint8 [H,W,3]."""
        pose = get_camera_poses_from_list_of_coordinates(init_coordinates, [coordinates])
        return self.render_poses(pose)[0]

    def render_coordinates_batch(self, init_coordinates: COORD, coordinates: Sequence[COORD]) -> np.ndarray:
        """Many views of one spot in one launch sequence (the GUI's camera sweep) -> uint8 [B,H,W,3]."""
        return self.render_poses(get_camera_poses_from_list_of_coordinates(init_coordinates, list(coordinates)))

    def _pinned_frames(self, nbytes: int) -> torch.Tensor:
        """Persistent page-locked host buffer for the uint8 read-back (grown on demand, never per frame)."""
        if self._host_frames is None or self._host_frames.numel() < nbytes:
            self._host_frames = torch.empty((nbytes,), dtype=torch.uint8, pin_memory=True)
        return self._host_frames[:nbytes]

    def render_rays_u8(self, c2w_dev: torch.Tensor, ray0: int, count: int, out: Optional[torch.Tensor] = None
                       ) -> torch.Tensor:
        """uint8 pixels [count,3] of the rays [ray0, ray0+count) of the B*H*W rays of the poses `c2w_dev` ([B,4,4],
        on the device or in host memory): raygen + the 8-launch render, the compositing kernel writes the bytes.  This is the unit the
        multi-GPU path shards (nwx/dist.py).  Repeated shapes replay a captured CUDA graph (use_cuda_graphs); the
        returned tensor is then the graph's static output buffer, overwritten by the next call of the same shape."""
        eng = self.engine
        if (self.use_cuda_graphs and out is None and not eng.profiling and 0 < count <= self.max_rays_per_launch
                and not torch.cuda.is_current_stream_capturing()):
            key = (int(c2w_dev.shape[0]), int(ray0), int(count), self._img_h, self._img_w, self._fx, self._fy, self._cx,
                   self._cy, self._depth_close_bound, self._depth_far_bound, self._n_samples, self._n_importance,
                   self._white_bkgd, self.max_rays_per_launch)       # everything the captured launches bake in
            g = self._graphs.get(key)
            if g is not None and g.stamp != (eng.weights_version, eng.scratch_state()[1]):
                g = None                              # weights reloaded or scratch re-allocated since the capture
            if g is None:
                if len(self._graphs) >= 64:           # e.g. a sweep over many batch sizes: start over
                    self._graphs.clear()
                g = self._graphs[key] = _FrameGraph(self, int(c2w_dev.shape[0]), int(ray0), int(count))
            return g.run(c2w_dev)
        rgb8 = torch.empty((count, 3), device=self._device, dtype=torch.uint8) if out is None else out
        c2w_dev = c2w_dev.to(self._device, dtype=torch.float32, non_blocking=True)
        return self._render_rays_u8_eager(c2w_dev, ray0, count, rgb8)

    def _render_rays_u8_eager(self, c2w_dev: torch.Tensor, ray0: int, count: int, rgb8: torch.Tensor) -> torch.Tensor:
        eng = self.engine
        for s in range(0, count, self.max_rays_per_launch):
            n = min(self.max_rays_per_launch, count - s)
            rays = eng.raygen(c2w_dev, self._img_h, self._img_w, self._fx, self._fy, self._cx, self._cy,
                              self._depth_close_bound, self._depth_far_bound, True, ray0=ray0 + s, nrays=n)
            eng.render_rays(rays, self._n_samples, self._n_importance, self._white_bkgd, want=("rgb8_fine",),
                            out={"rgb8_fine": rgb8[s:s + n]})
        return rgb8

    def frames_to_host(self, rgb8: torch.Tensor, B: int) -> np.ndarray:
        """Device uint8 [B*H*W,3] -> a fresh host array [B,H,W,3]: one cudaMemcpyAsync into the persistent
        pinned buffer on the current stream, one stream synchronisation, one host memcpy."""
        host = self._pinned_frames(rgb8.numel())
        host.copy_(rgb8.reshape(-1), non_blocking=True)
        torch.cuda.current_stream(self._device).synchronize()
        return host.numpy().reshape(B, self._img_h, self._img_w, 3).copy()

    @torch.no_grad()
    def render_poses(self, c2w: torch.Tensor) -> np.ndarray:
        """[B,4,4] camera-to-world poses -> uint8 [B,H,W,3]; H2D 64 B per view, D2H 3 B per pixel."""
        B = c2w.shape[0]
        return self.frames_to_host(self.render_rays_u8(c2w, 0, B * self._n_pix), B)

    def _render_rays(self, flat_rays: torch.Tensor) -> Dict[str, torch.Tensor]:
        """handler:187-201 -> the 11-key dict for [n,11] rays.  The reference chunks by
        self._chunk to bound memory; the fused path only chunks at max_rays_per_launch."""
        shape = flat_rays.shape
        if flat_rays.shape[0] <= self.max_rays_per_launch:
            out = self._volumetric_rendering(flat_rays)
        else:
            parts: Dict[str, List[torch.Tensor]] = {}
            for s in range(0, flat_rays.shape[0], self.max_rays_per_launch):
                for k, v in self._volumetric_rendering(flat_rays[s:s + self.max_rays_per_launch]).items():
                    parts.setdefault(k, []).append(v)
            out = {k: torch.cat(v, 0) for k, v in parts.items()}
        return {k: torch.reshape(v, list(shape[:-1]) + list(v.shape[1:])) for k, v in out.items()}

    @torch.no_grad()
    def _volumetric_rendering(self, ray_batch: torch.Tensor) -> Dict[str, torch.Tensor]:
        """handler:203-277: inference never jitters, never adds noise and always samples the fine
        depths deterministically (`det=... or True`, :238)."""
        out = self.engine.render_rays(ray_batch.to(self._device), self._n_samples, self._n_importance,
                                      self._white_bkgd, want=_engine.REFERENCE_KEYS)
        flags = out.pop("flags")
        self.last_flags = flags      # device int32: bit0 NaN, bit1 Inf (replaces the 22 syncs of :273-275)
        return {k: out[k] for k in _engine.REFERENCE_KEYS}

    def check_numerics(self) -> None:
        """Reads the device flag word (one sync) and prints the reference's warning (handler:273-275)."""
        flags = int(getattr(self, "last_flags", torch.zeros(1)).item())
        if flags:
            print(f"[Numerical Error] outputs contain {'NaN ' if flags & 1 else ''}{'inf' if flags & 2 else ''}.")
