"""ctypes binding of libnwx.so -- the C ABI declared in include/nwx.h.

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
The library is built in-tree (``make -C nerf-workspaces-explorer_b200`` or
``__graft_entry__.build()``) so that it travels with the repository snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libnwx.so")

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float


class RenderOpts(C.Structure):
    _fields_ = [("n_samples", _i), ("n_importance", _i), ("white_bkgd", _i), ("ray_dim", _i),
                ("t_vals", _vp), ("u_lin", _vp), ("t_rand", _vp), ("u", _vp),
                ("noise_coarse", _vp), ("noise_fine", _vp),
                ("rng_seed", C.c_uint64), ("rng_offset", C.c_uint64), ("raw_noise_std", _f),
                ("rng_jitter", _i), ("rng_u", _i)]


RENDER_OUT_FIELDS = ("rgb_coarse", "disp_coarse", "acc_coarse", "depth_coarse", "raw_coarse",
                     "rgb_fine", "disp_fine", "acc_fine", "depth_fine", "raw_fine", "z_std",
                     "z_vals_coarse", "weights_coarse", "z_samples", "z_vals_fine", "weights_fine",
                     "inds", "flags", "rgb8_fine")


class RenderOut(C.Structure):
    _fields_ = [(name, _vp) for name in RENDER_OUT_FIELDS]


class TrainIO(C.Structure):
    _fields_ = [(name, _vp) for name in ("rays", "gt_rgb", "grad_coarse", "grad_fine", "loss", "rgb_coarse",
                                         "rgb_fine", "ev_coarse_done")]


# name -> (restype, argtypes); mirrors include/nwx.h one to one
PROTOTYPES = {
    "nwx_version": (_i, []),
    "nwx_error_string": (C.c_char_p, [_i]),
    "nwx_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "nwx_ctx_destroy": (_i, [_vp]),
    "nwx_load_weights": (_i, [_vp, _i, C.POINTER(_vp), _vp]),
    "nwx_raygen": (_i, [_vp, _i, _i, _i, _f, _f, _f, _f, _f, _f, _i, _i64, _i64, _vp, _vp]),
    "nwx_coarse_z": (_i, [_vp, _i, _i64, _i, _vp, _vp, _vp, _vp]),
    "nwx_mlp_forward": (_i, [_vp, _i, _vp, _i, _vp, _i64, _i, _vp, _vp]),
    "nwx_mlp_forward_points": (_i, [_vp, _i, _vp, _vp, _i64, _i, _vp, _vp]),
    "nwx_mlp_forward_embedded": (_i, [_vp, _i, _vp, _i64, _vp, _vp]),
    "nwx_embed": (_i, [_vp, _i64, _i, _f, _vp, _vp]),
    "nwx_composite_fwd": (_i, [_vp, _vp, _vp, _i, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nwx_composite_bwd": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp]),
    "nwx_sample_pdf": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i64, _vp, _vp, _vp, _vp, _vp]),
    "nwx_sample_pdf_bins": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i64, _vp, _vp, _vp, _vp]),
    "nwx_ctx_reserve": (_i, [_vp, _i64, _i, _i]),
    "nwx_ctx_scratch_state": (_i, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "nwx_param_offsets": (_i, [C.POINTER(_i)]),
    "nwx_train_pack": (_i, [_vp, _i, _vp, _vp]),
    "nwx_train_fwd_bwd": (_i, [_vp, C.POINTER(TrainIO), _i64, C.POINTER(RenderOpts), _vp]),
    "nwx_adam_step": (_i, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _i, _f, _vp]),
    "nwx_adam_pack_step": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _i, _f, _vp]),
    "nwx_debug_copy_packed": (_i, [_vp, _i, _i, _vp, _i64, _vp]),
    "nwx_ctx_set_profiling": (_i, [_vp, _i]),
    "nwx_ctx_stage_ms": (_i, [_vp, C.POINTER(_f)]),
    "nwx_render_rays": (_i, [_vp, _vp, _i64, C.POINTER(RenderOpts), C.POINTER(RenderOut), _vp]),
    "nwx_rng_fill": (_i, [_i, C.c_uint64, C.c_uint64, C.c_uint32, _f, _i64, _vp, _vp]),
    "nwx_to8b": (_i, [_vp, _i64, _vp, _vp]),
    "nwx_launch_count": (_i64, []),
    "nwx_set_mlp_variant": (_i, [_vp, _i]),
    "nwx_debug_tap": (_i, [_vp, _i, _vp]),
    "nwx_debug_diag": (_i, [_vp, _vp]),
    "nwx_debug_experiment": (_i, [_vp, _i]),
    "nwx_ctx_last_diag": (_i, [_vp, C.POINTER(C.c_uint32)]),
    "nwx_sample_training_batch": (_i, [_vp, _vp, _i, _i64, _i, _i64, C.c_uint64, C.c_uint64, _vp, _vp, _vp, _vp]),
}


class NwxError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile libnwx.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    proc = subprocess.run(["make", "-C", PKG_ROOT, "-j8"], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout[-4000:])
        print(proc.stderr[-4000:])
    if proc.returncode != 0:
        raise NwxError("building libnwx.so failed (see output above)")
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NwxError(f"{LIB_PATH} is missing: build it with `make -C {PKG_ROOT}` "
                           "(there is no CPU or PyTorch fallback path)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = lib().nwx_error_string(code).decode()
        raise NwxError(f"{what or 'libnwx call'} failed: [{code}] {msg}")
