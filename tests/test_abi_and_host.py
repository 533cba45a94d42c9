"""CPU: the C-ABI library builds/loads and exports every symbol include/nwx.h declares (no compute
without a GPU), and the host-side logic around it (pose producer, config, state-dict plumbing,
drop-in module structure, failure behaviour)."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import nerf_oracle as orc


@pytest.fixture(scope="module")
def nwx_mod():
    import nwx
    if not os.path.exists(nwx._lib.LIB_PATH):
        nwx.build()
    return nwx


def test_library_exports_every_declared_symbol(nwx_mod):
    header = open(os.path.join(ROOT, "include", "nwx.h")).read()
    declared = sorted(set(re.findall(r"\b(nwx_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 20
    lib = nwx_mod.lib()
    missing = [name for name in declared if not hasattr(lib, name)]
    assert not missing, missing
    from nwx._lib import PROTOTYPES
    assert set(declared) == set(PROTOTYPES)                    # the binding covers the header one to one
    assert lib.nwx_version() == 100
    assert lib.nwx_error_string(0) == b"ok" and b"invalid" in lib.nwx_error_string(1)
    assert b"fallback" in lib.nwx_error_string(3)


def test_library_is_blackwell_native(nwx_mod):
    """SASS evidence (B200_PROFILING.md): tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, bulk TMA ->
    UBLKCP; and no legacy HMMA tensor path."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", nwx_mod._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR"):
        assert mnemonic in sass, mnemonic
    assert "UTCHMMA.2CTA" in sass                              # cta_group::2 variant present
    assert not re.search(r"\bHMMA\b", sass)
    # the MLP epilogue: biases through uniform registers (LDCU) into packed fp32x2 adds (FADD2), packed
    # bf16 conversion with the ReLU folded in, 128-bit shared-memory stores; TMA tile stores in training
    for mnemonic in ("LDCU.64", "FADD2", "FFMA2", "F2FP.RELU.BF16.F32.PACK_AB", "STS.128", "UBLKCP.G.S"):
        assert mnemonic in sass, mnemonic


def test_no_gpu_means_error_not_fallback(nwx_mod):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nwx_mod.NwxError):
        nwx_mod.Engine()
    with pytest.raises(nwx_mod.NwxError):
        nwx_mod.raw2outputs(torch.zeros(2, 64, 4), torch.zeros(2, 64), torch.zeros(2, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "nerf-workspaces-explorer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "/root/reference" not in text, f


def test_pose_producer_matches_reference(nwx_mod):
    g = load_golden("poses36")
    rng = np.random.RandomState(0)
    x, z = rng.uniform(-2, 2), rng.uniform(-3, 1.5)
    init = nwx_mod.COORD(x=x, y=-0.5, z=z, yaw=0.0, pitch=-90.0, roll=0.0)
    views = [nwx_mod.COORD(yaw=-float(h), pitch=float(v)) for v in (-30, 0, 30) for h in range(0, 360, 30)]
    poses = nwx_mod.get_camera_poses_from_list_of_coordinates(init, views)
    assert poses.shape == (36, 4, 4) and poses.dtype == torch.float32
    assert torch.allclose(poses, g["poses"], atol=1e-6, rtol=0)
    assert nwx_mod.COORD() == (0.0,) * 6 and nwx_mod.HW(3, 4).w == 4


def test_config_and_state_dict_plumbing(nwx_mod):
    from nwx.config import default_config, number
    from nwx.engine import STATE_KEYS, normalize_state_dict
    cfg = default_config()
    assert number(cfg["model"]["net_chunk"]) == 32768 and number(cfg["inference"]["chunk"]) == 8192
    assert number(cfg["rendering"]["n_rays"]) == 1024 and number(7) == 7
    sd = orc.init_state_dict(0)
    assert tuple(STATE_KEYS) == tuple(orc.STATE_KEYS)
    shipped = {k[1:]: v for k, v in sd.items()}                 # checkpoint style: no leading underscore
    assert set(normalize_state_dict(shipped)) == set(sd)
    H = nwx_mod.NeRFReplicaInferenceHandler
    assert set(H.transform_state_dict(shipped)) == set(sd)
    with pytest.raises(nwx_mod.NwxError):
        normalize_state_dict({k: v for k, v in sd.items() if "alpha" not in k})
    bad = dict(sd); bad["_pts_linears.0.weight"] = torch.zeros(128, 63)
    with pytest.raises(nwx_mod.NwxError):
        normalize_state_dict(bad)


def test_drop_in_modules_on_cpu(nwx_mod):
    torch.manual_seed(0)
    m = nwx_mod.NeRFModel(8, 256, 63, 27, 5, use_view_dirs=True)
    sd = orc.init_state_dict(0, alpha_bias=None)
    assert list(m.state_dict()) == list(orc.STATE_KEYS)
    assert all(torch.equal(m.state_dict()[k], sd[k]) for k in sd)     # same construction order / RNG stream
    assert sum(p.numel() for p in m.parameters()) == 595844
    assert nwx_mod.Embedding(10, 10).output_dim == 63 and nwx_mod.Embedding(4, 1).output_dim == 27
    h = nwx_mod.NeRFReplicaInferenceHandler("office_tokyo", "/nonexistent/model.ckpt")
    assert (h._img_h, h._img_w, h._cx, h._cy) == (240, 320, 159.5, 119.5) and abs(h._fx - 160.0) < 1e-9
    with pytest.raises(RuntimeError, match="cannot be found"):     # reference behaviour, handler:147-148
        h.initialize_models()
    with pytest.raises(RuntimeError):
        h.engine
    out = nwx_mod.batchify(lambda x: x * 2, 3)(torch.arange(10.))
    assert torch.equal(out, torch.arange(10.) * 2)
    parts = nwx_mod.batchify_rays(lambda r: {"a": r[:, 0]}, torch.arange(20.).reshape(10, 2), chunk=4)
    assert torch.equal(parts["a"], torch.arange(0., 20., 2))


def test_workspace_transforms_match_reference(nwx_mod):
    """application/workspace.py:91-196: floor-plan click -> (camera COORD, view COORD), all 4 rooms."""
    import numpy as np
    from conftest import GOLDEN
    import nwx.workspace as ws
    g = np.load(os.path.join(GOLDEN, "workspace.npz"))
    for cls in ("OfficeTokyoWorkspace", "OfficeNewYorkWorkspace", "OfficeGeneveWorkspace", "OfficeBelgradeWorkspace"):
        w = getattr(ws, cls)()
        for case, ref in zip(g["cases"], g[cls]):
            init, view = w._transform_relative_coordinates(float(case[0]), float(case[1]), int(case[2]), int(case[3]))
            assert list(init) + list(view) == list(ref), (cls, case)
    assert repr(ws.OfficeGeneveWorkspace()) == "Office Geneve" and ws.OfficeGeneveWorkspace().floor_plan_scale == (600, 1000)
    with pytest.raises(KeyError):
        ws.Workspace("Office Atlantis")
    with pytest.raises(RuntimeError, match="cannot be found"):
        ws.OfficeTokyoWorkspace().initialize_models()              # no checkpoint shipped: reference behaviour


def test_run_network_recognises_the_handlers_lambda(nwx_mod):
    """inference handler:248 passes `lambda x: self._nerf_net_fine(x, self._endpoint_feat)`: the probe must
    resolve it to the model (fused path) and must NOT resolve callables that are not pure pass-throughs."""
    from nwx.models import _resolve_model
    net = nwx_mod.NeRFModel(8, 256, 63, 27, 5, use_view_dirs=True)
    other = nwx_mod.NeRFModel(8, 256, 63, 27, 5, use_view_dirs=True)
    like = torch.zeros(1)

    class Handler:
        _nerf_net_fine, _endpoint_feat = net, False
    h = Handler()
    assert _resolve_model(net, like) is net
    assert _resolve_model(lambda x: h._nerf_net_fine(x, h._endpoint_feat), like) is net
    assert _resolve_model(lambda x: net(x), like) is net
    assert _resolve_model(lambda x: 2.0 * net(x), like) is None              # post-processes the output
    assert _resolve_model(lambda x: net(x * 1.0), like) is None              # pre-processes the input
    assert _resolve_model(lambda x: net(x) + other(x), like) is None         # two models
    assert _resolve_model(lambda x: net(x, True), like) is None              # show_endpoint
    assert _resolve_model(lambda x: x[:, :4], like) is None                  # no model at all
    assert _resolve_model(lambda x: 1 / 0, like) is None                     # raises on the probe
    small = nwx_mod.NeRFModel(4, 128, 63, 27, 5, use_view_dirs=True)         # not the fused architecture
    assert _resolve_model(small, like) is None and _resolve_model(lambda x: small(x), like) is None
    from nwx.models import _Probe
    assert _Probe.active is None                                            # always disarmed afterwards


def test_patch_reference_routes_the_reference_imports(nwx_mod):
    """nwx.patch_reference(): the five hot-path modules of the reference resolve to nwx, and are restored by
    unpatch_reference().  With the reference checkout present (build container), its UNMODIFIED inference
    handler module then binds nwx's functions."""
    import importlib
    import sys
    names = ("nerf.rays.rays", "nerf.models.embedding", "nerf.models.nerf_model", "nerf.models.model_utils",
             "utils.batch_utils")
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("nerf", "utils")}
    ref_root = os.path.join(os.sep, "root", "reference")
    have_ref = os.path.isdir(os.path.join(ref_root, "nerf"))
    if have_ref:
        sys.path.insert(0, ref_root)
    try:
        nwx_mod.patch_reference()
        from nerf.models.model_utils import raw2outputs, run_network, to8b_np   # noqa: F401
        from nerf.models.nerf_model import NeRFModel
        from nerf.rays.rays import create_rays, sample_pdf
        from utils.batch_utils import batchify_rays
        assert run_network is nwx_mod.run_network and raw2outputs is nwx_mod.raw2outputs
        assert NeRFModel is nwx_mod.NeRFModel and create_rays is nwx_mod.create_rays and sample_pdf is nwx_mod.sample_pdf
        assert batchify_rays is nwx_mod.batchify_rays
        assert all(getattr(sys.modules[n], "__nwx_patch__", False) for n in names)
        if have_ref:
            try:
                mod = importlib.import_module("nerf.inference.nerf_replica_inference_handler")
            except ImportError as exc:          # a dependency of the reference (cv2, yaml) missing here
                pytest.skip(f"reference handler not importable: {exc}")
            assert mod.run_network is nwx_mod.run_network and mod.NeRFModel is nwx_mod.NeRFModel
            assert mod.create_rays is nwx_mod.create_rays and mod.sample_pdf is nwx_mod.sample_pdf
            assert mod.batchify_rays is nwx_mod.batchify_rays and mod.Embedding is nwx_mod.Embedding
    finally:
        nwx_mod.unpatch_reference()
        for k in [k for k in sys.modules if k.split(".")[0] in ("nerf", "utils") and k not in saved]:
            del sys.modules[k]
        sys.modules.update(saved)
        if have_ref and ref_root in sys.path:
            sys.path.remove(ref_root)
    assert not any(getattr(sys.modules.get(n), "__nwx_patch__", False) for n in names)


def test_forward_only_guard_fails_loudly_on_backward(nwx_mod):
    """The fused forward keeps no autograd graph: a backward() through it must raise, not silently skip."""
    from nwx.models import _ForwardOnly
    w = torch.nn.Parameter(torch.ones(3))
    out = _ForwardOnly.apply(torch.zeros(2, 4), w)
    assert out.requires_grad
    with pytest.raises(nwx_mod.NwxError, match="Trainer"):
        out.sum().backward()
