"""Training on the CUDA engine: the compute of NeRFReplicaTrainingHandler.step (reference
nerf/training/nerf_replica_training_handler.py:265-339) -- ray/pixel sampling, render in training
mode (stratified jitter, sigma noise, random importance samples), MSE(coarse)+MSE(fine), backward,
Adam, exponential learning-rate decay -- with every kernel hand-written (libnwx) and the
data-parallel gradient all-reduce on torch.distributed (NCCL).  Dataset loading, TensorBoard, eval
renders and checkpoint writing stay with the caller (out of scope)."""
from __future__ import annotations

import os
import ctypes as C
from typing import Dict, Mapping, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from . import engine as _engine
from ._lib import RenderOpts, TrainIO, check
from .config import default_config, number
from .engine import COARSE, FINE, STATE_KEYS, STATE_SHAPES, _f32, _ptr, _stream, linspace01

PARAMS_PER_NET = 595844


def param_offsets() -> Tuple[int, ...]:
    buf = (C.c_int * 24)()
    check(_lib.lib().nwx_param_offsets(buf), "nwx_param_offsets")
    return tuple(int(v) for v in buf)


def rank_of(group: Optional[dist.ProcessGroup] = None) -> int:
    """This process' rank in the data-parallel group (0 without torch.distributed)."""
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def world_of(group: Optional[dist.ProcessGroup] = None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank_seed(seed: int, group: Optional[dist.ProcessGroup] = None) -> int:
    """Seed of the per-rank random streams (pixel sampling, jitter, importance uniforms, sigma noise): every
    rank must draw a DIFFERENT batch, otherwise the all-reduce averages N copies of one gradient and data
    parallelism buys nothing.  Parameter initialisation keeps the unshifted seed (identical on every rank)."""
    return int(seed) + rank_of(group)


def allreduce_sum_(flat: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> float:
    """Sum `flat` over the ranks in place (one collective for both networks' gradients) and return
    the factor that turns the sum into the data-parallel mean (1 / world_size): per-rank batches
    are equal, so the mean of per-rank mean-losses' gradients is the global-batch gradient."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat, group=group)
    return 1.0 / world


class Trainer:
    """Flat fp32 master parameters / gradients / Adam moments for the coarse and fine networks
    (one [2, 595 844] tensor each, so the data-parallel all-reduce is a single NCCL call)."""

    def __init__(self, eng: _engine.Engine, sd_coarse: Mapping[str, torch.Tensor], sd_fine: Mapping[str, torch.Tensor],
                 lr: float = 5e-4, lr_decay_rate: float = 0.1, lr_decay_steps: int = 50000,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, n_samples: int = 64,
                 n_importance: int = 128, white_bkgd: bool = False, perturb: float = 1.0, raw_noise_std: float = 1.0,
                 group: Optional[dist.ProcessGroup] = None, seed: int = 0, overlap_allreduce: bool = True,
                 data_parallel: bool = True):
        """`group`: the data-parallel process group (None = the default group when torch.distributed is
        initialised).  data_parallel=False keeps this trainer local even under torch.distributed (no collectives,
        unshifted seed) -- e.g. a single-rank reference run inside a multi-rank job."""
        self.engine, self.device, self.group = eng, eng.device, group
        self.data_parallel = bool(data_parallel)
        self.lr0, self.lr, self.lr_decay_rate, self.lr_decay_steps = lr, lr, lr_decay_rate, lr_decay_steps
        self.betas, self.eps = betas, eps
        self.n_samples, self.n_importance, self.white_bkgd = n_samples, n_importance, white_bkgd
        self.perturb, self.raw_noise_std = perturb, raw_noise_std
        self.offsets = param_offsets()
        self.params = torch.empty((2, PARAMS_PER_NET), device=self.device)
        for w, sd in ((COARSE, sd_coarse), (FINE, sd_fine)):
            sd = _engine.normalize_state_dict(sd)
            self.params[w].copy_(torch.cat([sd[k].detach().reshape(-1).float() for k in STATE_KEYS]).to(self.device))
        self.grads = torch.zeros_like(self.params)
        self.m = torch.zeros_like(self.params)
        self.v = torch.zeros_like(self.params)
        self.loss = torch.zeros(2, device=self.device, dtype=torch.float64)
        self.opt_steps = 0
        # in-kernel RNG key (rank-folded: each rank draws its own jitter / u / noise / pixels) and per-call offset
        self.base_seed, self.seed, self.draws = int(seed), (rank_seed(seed, group) if self.data_parallel else int(seed)), 0
        self.world = world_of(group) if self.data_parallel else 1
        if self.world > 1:
            # identical parameters on every rank regardless of what each rank was constructed with
            dist.broadcast(self.params, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        # data-parallel overlap: the coarse network's gradients are complete before the fine network's
        # backward starts, so their all-reduce runs on a side stream underneath it
        self._overlap = (bool(overlap_allreduce) and self.world > 1 and self.device.type == "cuda"
                         and os.environ.get("NWX_TRAIN_OVERLAP", "1") != "0")      # env: A/B switch for measurements
        self._comm_stream = torch.cuda.Stream(self.device) if self._overlap else None
        self._ev_coarse = torch.cuda.Event() if self._overlap else None
        self._coarse_reduced = False
        self.pack()

    # ---- parameters ------------------------------------------------------------------------
    def state_dict(self, which: int) -> Dict[str, torch.Tensor]:
        """Views of the flat master parameters under the reference's state_dict keys."""
        flat = self.params[which]
        return {k: flat[o:o + int(torch.Size(STATE_SHAPES[k]).numel())].view(STATE_SHAPES[k])
                for k, o in zip(STATE_KEYS, self.offsets)}

    def grad_dict(self, which: int) -> Dict[str, torch.Tensor]:
        flat = self.grads[which]
        return {k: flat[o:o + int(torch.Size(STATE_SHAPES[k]).numel())].view(STATE_SHAPES[k])
                for k, o in zip(STATE_KEYS, self.offsets)}

    def pack(self) -> None:
        for w in (COARSE, FINE):
            check(self.engine._lib.nwx_train_pack(self.engine._ctx, w, self.params[w].data_ptr(), _stream()),
                  "nwx_train_pack")

    PACKED = {"wimg": (0, 1245184), "wimg_t": (1, 983040), "consts": (2, None), "wdir_t": (3, 27 * 128 * 4),
              "bview": (4, 512), "bview_fold": (5, 512)}

    def packed_bytes(self, which: int, what: str) -> torch.Tensor:
        """Test hook: one packed device buffer of a network as a uint8 tensor (see nwx_debug_copy_packed)."""
        code, size = self.PACKED[what]
        if size is None:
            size = 4 * (9 * 256 + 256 + 3 * 128 + 4)            # sizeof(MlpConsts)
        out = torch.empty((size,), device=self.device, dtype=torch.uint8)
        check(self.engine._lib.nwx_debug_copy_packed(self.engine._ctx, which, code, out.data_ptr(), size, _stream()),
              "nwx_debug_copy_packed")
        return out

    def sync_inference_weights(self) -> None:
        """Refresh the engine's inference-side copy (host constants) from the master parameters."""
        self.pack()                                              # training-side images / device constants
        self.engine.load_weights(COARSE, self.state_dict(COARSE))   # host-side constants (clears the stale mark)
        self.engine.load_weights(FINE, self.state_dict(FINE))

    # ---- checkpoints (reference format, training handler:394-409) ---------------------------------
    def checkpoint(self, global_step: int) -> Dict[str, object]:
        """{global_step, network_coarse_state_dict, network_fine_state_dict, optimizer_state_dict}; the
        optimizer entry has torch.optim.Adam's layout for the 48 parameters (coarse then fine, the
        order of `learnable_params`, training handler:227-234), so the reference can resume from it."""
        sds = [{k: v.detach().clone().cpu() for k, v in self.state_dict(w).items()} for w in (COARSE, FINE)]
        state, idx = {}, 0
        for w in (COARSE, FINE):
            for k, o in zip(STATE_KEYS, self.offsets):
                n = int(torch.Size(STATE_SHAPES[k]).numel())
                state[idx] = {"step": torch.tensor(float(self.opt_steps)),
                              "exp_avg": self.m[w, o:o + n].view(STATE_SHAPES[k]).clone().cpu(),
                              "exp_avg_sq": self.v[w, o:o + n].view(STATE_SHAPES[k]).clone().cpu()}
                idx += 1
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(idx))}
        # nwx_rng: not a reference key (its loaders ignore it); lets a resumed run continue the random streams
        # instead of replaying the draws of steps 0..k
        return {"global_step": global_step, "network_coarse_state_dict": sds[0], "network_fine_state_dict": sds[1],
                "optimizer_state_dict": {"state": state, "param_groups": [group]},
                "nwx_rng": {"seed": self.base_seed, "draws": self.draws}}

    def save_checkpoint(self, path: str, global_step: int) -> None:
        torch.save(self.checkpoint(global_step), path)

    def load_checkpoint(self, ckpt: Mapping[str, object]) -> int:
        """Resume from a checkpoint dict in the reference format (either key style)."""
        for w, name in ((COARSE, "network_coarse_state_dict"), (FINE, "network_fine_state_dict")):
            sd = _engine.normalize_state_dict(ckpt[name])
            self.params[w].copy_(torch.cat([sd[k].detach().reshape(-1).float() for k in STATE_KEYS]).to(self.device))
        opt = ckpt.get("optimizer_state_dict")
        if opt and opt.get("state"):
            idx = 0
            for w in (COARSE, FINE):
                for k, o in zip(STATE_KEYS, self.offsets):
                    n = int(torch.Size(STATE_SHAPES[k]).numel())
                    st = opt["state"][idx]
                    self.m[w, o:o + n].copy_(st["exp_avg"].reshape(-1).to(self.device))
                    self.v[w, o:o + n].copy_(st["exp_avg_sq"].reshape(-1).to(self.device))
                    self.opt_steps = int(st["step"])
                    idx += 1
            self.lr = float(opt["param_groups"][0]["lr"])
        step = int(ckpt.get("global_step", 0))
        rng = ckpt.get("nwx_rng")
        # continue the random streams: our own checkpoints carry the draw counter; a reference checkpoint does
        # not, there the optimiser step count (one render per step) is the best available offset
        self.draws = int(rng["draws"]) if rng else max(self.draws, self.opt_steps)
        self.pack()
        return step

    # ---- one step ----------------------------------------------------------------------------
    def forward_backward(self, rays: torch.Tensor, gt_rgb: torch.Tensor, t_rand: Optional[torch.Tensor] = None,
                         u: Optional[torch.Tensor] = None, noise_coarse: Optional[torch.Tensor] = None,
                         noise_fine: Optional[torch.Tensor] = None, want_rgb: bool = False):
        """Render `rays` in training mode, loss = mse(rgb_c, gt) + mse(rgb_f, gt), gradients into
        self.grads.  The three random draws of the reference (t_rand training handler:560, noise
        model_utils.py:65, u rays.py:98) are generated inside the kernels (counter-based, keyed by the
        trainer's seed and a per-call offset) unless tensors are injected."""
        rays, gt = _f32(rays, "rays"), _f32(gt_rgb, "gt_rgb")
        N, dev = rays.shape[0], rays.device
        Sc, Ni = self.n_samples, self.n_importance
        keep = [None if t is None else _f32(t, "rand") for t in (t_rand, u, noise_coarse, noise_fine)]
        rng = _engine.RngOptions(self.seed, self.draws, jitter=self.perturb > 0., random_u=self.perturb > 0.,
                                 noise_std=self.raw_noise_std if self.raw_noise_std > 0. else 0.0)
        self.draws += 1
        opts = RenderOpts(Sc, Ni, int(self.white_bkgd), rays.shape[1], linspace01(Sc, dev).data_ptr(),
                          linspace01(Ni, dev).data_ptr(), _ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2]), _ptr(keep[3]),
                          *rng.fields())
        rgb_c = torch.empty((N, 3), device=dev) if want_rgb else None
        rgb_f = torch.empty((N, 3), device=dev) if want_rgb else None
        if self._overlap:      # the previous step's side-stream all-reduce must be done with grads[COARSE]
            torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)
        ev = None
        if self._overlap:
            self._ev_coarse.record()       # torch creates the handle lazily; the library re-records it mid-call
            ev = self._ev_coarse.cuda_event
        io = TrainIO(rays.data_ptr(), gt.data_ptr(), self.grads[COARSE].data_ptr(), self.grads[FINE].data_ptr(),
                     self.loss.data_ptr(), _ptr(rgb_c), _ptr(rgb_f), ev)
        check(self.engine._lib.nwx_train_fwd_bwd(self.engine._ctx, C.byref(io), N, C.byref(opts), _stream()),
              "nwx_train_fwd_bwd")
        self._coarse_reduced = False
        if self._overlap:
            self._comm_stream.wait_event(self._ev_coarse)
            with torch.cuda.stream(self._comm_stream):
                dist.all_reduce(self.grads[COARSE], group=self.group)      # underneath the fine network's backward
            self._coarse_reduced = True
        return (self.loss, rgb_c, rgb_f) if want_rgb else self.loss

    def optimizer_step(self, global_step: int) -> None:
        """All-reduce (mean) the gradients across ranks, Adam, re-pack, then the reference's
        learning-rate schedule lr = lr0 * rate^(step/decay_steps) (training handler:312-315),
        which -- as in the reference -- takes effect from the NEXT step."""
        if not self.data_parallel:
            scale = 1.0
        elif self._coarse_reduced:                                    # coarse half already in flight on the side stream
            scale = allreduce_sum_(self.grads[FINE], self.group)     # 2.38 MB
            torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)
            self._coarse_reduced = False
        else:
            scale = allreduce_sum_(self.grads, self.group)           # 4.77 MB, one NCCL call
        self.apply_optimizer(global_step, scale)

    def apply_optimizer(self, global_step: int, grad_scale: float = 1.0) -> None:
        """Adam on self.grads * grad_scale, re-pack, learning-rate schedule (the local part of optimizer_step)."""
        self.opt_steps += 1
        # two launches: Adam over both networks fused with the re-pack of every kernel image, then the folded views layer
        check(self.engine._lib.nwx_adam_pack_step(self.engine._ctx, self.params.data_ptr(), self.grads.data_ptr(),
                                                  self.m.data_ptr(), self.v.data_ptr(), self.lr, self.betas[0],
                                                  self.betas[1], self.eps, self.opt_steps, grad_scale, _stream()),
              "nwx_adam_pack_step")
        self.lr = self.lr0 * (self.lr_decay_rate ** (global_step / self.lr_decay_steps))

    def step(self, rays: torch.Tensor, gt_rgb: torch.Tensor, global_step: int, **rand) -> torch.Tensor:
        """One optimisation step; returns a COPY of the two MSE terms (self.loss is overwritten every step)."""
        self.forward_backward(rays, gt_rgb, **rand)
        self.optimizer_step(global_step)
        return self.loss.clone()


class NeRFReplicaTrainingHandler:
    """Drop-in for the compute of the reference training handler.  The reference builds its ray
    bank and pixel bank from the Replica dataset (`initialize_rays` :243-263, `prepare_data`
    :118-194); here they are passed in (any source, e.g. nwx.create_rays on the dataset's poses)."""

    def __init__(self, office_name: str, config: Optional[Mapping], rays_train: torch.Tensor, train_rgbs: torch.Tensor,
                 sd_coarse: Optional[Mapping] = None, sd_fine: Optional[Mapping] = None,
                 device: Optional[torch.device] = None, seed: int = 0, group: Optional[dist.ProcessGroup] = None):
        from .synthetic import random_state_dicts
        cfg = default_config() if config is None else config
        self._office_name, self._config = office_name, cfg
        rnd, trn = cfg["rendering"], cfg["training"]
        self._n_rays = number(rnd["n_rays"])
        self._img_h, self._img_w = int(cfg["experiment"]["image_height"]), int(cfg["experiment"]["image_width"])
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.rays_train = rays_train.to(dev).float().contiguous()   # [num_images, H*W, 11]
        self._train_rgbs = train_rgbs.to(dev).float().reshape(rays_train.shape[0], -1, 3).contiguous()
        if sd_coarse is None:
            sd_coarse, sd_fine = random_state_dicts(seed, alpha_bias=None)
        self._engine = _engine.Engine(dev)
        self.trainer = Trainer(self._engine, sd_coarse, sd_fine, lr=float(trn["learning_rate"]),
                               lr_decay_rate=float(trn["learning_rate_decay_rate"]),
                               lr_decay_steps=int(trn["learning_rate_decay_steps"]),
                               n_samples=int(rnd["n_samples"]), n_importance=int(rnd["n_importance"]),
                               white_bkgd=bool(rnd["white_background"]), perturb=float(rnd["perturb"]),
                               raw_noise_std=float(rnd["raw_noise_std"]), seed=seed, group=group)
        self._chunk = number(cfg.get("model", {}).get("chunk", 1024 * 32))   # rays per render call (yaml model.chunk)
        self._train_mode, self._weights_version = True, -1

    def _sample_training_data(self, want_indices: bool = False):
        """One random image, n_rays random pixels with replacement (training handler:341-370) -- drawn and
        gathered by ONE kernel on the device (no host round trip, no synchronisation); the draws are keyed by
        the trainer's rank-folded seed and its draw counter, so every rank and every step gets its own batch."""
        tr = self.trainer
        return self._engine.sample_training_batch(self.rays_train, self._train_rgbs, self._n_rays, tr.seed, tr.draws,
                                                  want_indices=want_indices)

    # ---- forward-only renders (eval renders of the reference's loop, training handler:479-532) ----
    def set_train_mode(self, on: bool) -> None:
        """training handler:110-116 toggles `_train_mode`: jitter / noise only while training."""
        self._train_mode = bool(on)

    @torch.no_grad()
    def _volumetric_rendering(self, ray_batch: torch.Tensor, t_rand: Optional[torch.Tensor] = None,
                              u: Optional[torch.Tensor] = None, noise_coarse: Optional[torch.Tensor] = None,
                              noise_fine: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """training handler:534-618 -> the reference's 11-key dict with the CURRENT weights.  In train mode the
        coarse depths are jittered (:547-565), u is random (:578) and sigma noise is added (raw_noise_std);
        the draws come from the in-kernel generator unless tensors are injected (tests inject the
        reference's own draws).  Forward only: gradients are produced by `step`."""
        tr = self.trainer
        if self._weights_version != tr.opt_steps:                 # host-side constants follow the optimiser
            tr.sync_inference_weights()
            self._weights_version = tr.opt_steps
        train = getattr(self, "_train_mode", True)
        rng = None
        if train:
            rng = _engine.RngOptions(tr.seed, tr.draws, jitter=tr.perturb > 0., random_u=tr.perturb > 0.,
                                     noise_std=tr.raw_noise_std if tr.raw_noise_std > 0. else 0.0)
            tr.draws += 1
        out = self._engine.render_rays(ray_batch.to(self._engine.device), tr.n_samples, tr.n_importance, tr.white_bkgd,
                                       want=_engine.REFERENCE_KEYS, t_rand=t_rand, u=u, noise_coarse=noise_coarse,
                                       noise_fine=noise_fine, rng=rng)
        self.last_flags = out.pop("flags")
        return {k: out[k] for k in _engine.REFERENCE_KEYS}

    def _render_rays(self, flat_rays: torch.Tensor) -> Dict[str, torch.Tensor]:
        """training handler:510-532: chunked render of [n,11] rays -> dict of [n, ...] tensors."""
        from .batch_utils import batchify_rays
        shape = flat_rays.shape
        out = batchify_rays(self._volumetric_rendering, flat_rays.to(self._engine.device), self._chunk)
        return {k: torch.reshape(v, list(shape[:-1]) + list(v.shape[1:])) for k, v in out.items()}

    def save_checkpoint(self, path: str, global_step: int) -> None:
        """training handler:394-409 (same dict layout, loadable by the reference's inference handler)."""
        self.trainer.save_checkpoint(path, global_step)

    def step(self, global_step: int) -> Dict[str, torch.Tensor]:
        rays, gt = self._sample_training_data()
        mse = self.trainer.step(rays, gt, global_step)
        psnr = -10.0 * torch.log10(mse)                            # mse2psnr, model_utils.py:8
        return {"rgb_loss_coarse": mse[0], "rgb_loss_fine": mse[1], "total_loss": mse.sum(),
                "psnr_coarse": psnr[0], "psnr_fine": psnr[1]}
