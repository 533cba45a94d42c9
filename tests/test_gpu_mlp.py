"""GPU parity of the fused PE + MLP tcgen05 kernel (K3).

Two checks per case: against the oracle's fp32 reference MLP (the reference's arithmetic; bf16
operands cost ~2e-4 on raw outputs of magnitude 0.2), and against the oracle's bf16 numerics
model of the kernel, which pins the implementation (layout, swizzle, descriptors, heads) an order
of magnitude tighter."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT, load_golden
from oracle import nerf_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

TOL_EMUL_MAX, TOL_EMUL_MEAN = 6e-4, 3e-5    # accumulation-order noise flipping a few bf16 roundings
TOL_FP32_MAX = 1.5e-3                        # bf16 operands vs the fp32 reference (|raw| ~ 0.2)


def _worker(*args, timeout=420):
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "gpu_worker.py"), *map(str, args)],
                          capture_output=True, text=True, timeout=timeout)
    lines = [l for l in proc.stdout.splitlines() if l.startswith("RESULT ")]
    assert lines, f"worker died (rc={proc.returncode}): {proc.stderr[-1500:]}"
    return json.loads(lines[-1][7:])


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_mlp_variants_isolated(variant):
    """Every kernel variant (production: CTA pair, resident weights, folded feature layer / CTA pair
    streaming / single CTA / CTA pair resident with the reference's layer structure), in its own
    process so a protocol bug is a failed test, not a hung GPU.  1000 points = ragged last tile."""
    r = _worker("mlp", variant, 1000)
    assert r["ok"], r
    assert r["finite"] and r["max_vs_emul"] <= TOL_EMUL_MAX and r["mean_vs_emul"] <= TOL_EMUL_MEAN, r
    assert r["max_vs_fp32"] <= TOL_FP32_MAX, r


@pytest.fixture(scope="module")
def eng():
    import nwx
    from nwx import engine as E
    e = nwx.Engine(torch.device(DEV))
    e.load_weights(E.COARSE, orc.init_state_dict(0))
    e.load_weights(E.FINE, orc.init_state_dict(7, trained_like=True))
    return e


def test_nerfmodel_forward_reference_signature():
    """NeRFModel(...).cuda()(x[P,90]) == the reference's module on its golden input."""
    import nwx
    g = load_golden("mlp")
    torch.manual_seed(0)
    model = nwx.NeRFModel(8, 256, 63, 27, 5, use_view_dirs=True)
    with torch.no_grad():
        model._alpha_linear.bias.fill_(0.1)
    sd = orc.init_state_dict(0)
    assert list(model.state_dict().keys()) == list(orc.STATE_KEYS)
    assert all(torch.equal(model.state_dict()[k], sd[k]) for k in sd)            # same init stream as the reference
    model = model.cuda()
    x = torch.cat([g["pe_xyz"], g["pe_dir"]], -1).to(DEV)
    raw = model(x).cpu()
    assert raw.shape == (160, 4)
    assert float((raw - g["raw"]).abs().max()) <= TOL_FP32_MAX
    emu = orc.mlp_forward_bf16_emul(sd, g["pe_xyz"], g["pe_dir"])
    assert float((raw - emu).abs().max()) <= TOL_EMUL_MAX
    with torch.no_grad():
        model._rgb_linear.bias.add_(1.0)                                          # weights changed -> re-packed
    assert float((model(x).cpu()[:, :3] - raw[:, :3] - 1.0).abs().max()) <= 1e-5
    with pytest.raises(nwx.NwxError):
        model(x, show_endpoint=True)


def test_run_network_fused_and_literal_paths():
    import nwx
    torch.manual_seed(0)
    model = nwx.NeRFModel(8, 256, 63, 27, 5, use_view_dirs=True).cuda()
    e3, e2 = nwx.Embedding(10, 10), nwx.Embedding(4, 1)
    g = torch.Generator().manual_seed(1)
    pts = ((torch.rand(37, 64, 3, generator=g) - 0.5) * 10).to(DEV)
    dirs = torch.nn.functional.normalize(torch.randn(37, 3, generator=g), dim=-1).to(DEV)
    fused = nwx.run_network(pts, dirs, model, e3.embed, e2.embed, 1024 * 32)
    literal = nwx.run_network(pts, dirs, lambda x: model(x), e3.embed, e2.embed, 1000)
    assert fused.shape == literal.shape == (37, 64, 4)
    assert float((fused - literal).abs().max()) <= TOL_EMUL_MAX
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    ref = orc.run_network(pts.cpu(), dirs.cpu(), lambda x: orc.mlp_forward(sd, x),
                          lambda x: orc.positional_encoding(x, 10, 10), lambda x: orc.positional_encoding(x, 4, 1))
    assert float((fused.cpu() - ref).abs().max()) <= TOL_FP32_MAX


@pytest.mark.parametrize("n_rays,S", [(1, 64), (2, 64), (5, 192), (257, 64), (100, 192)])
def test_mlp_rays_mode_tiles_and_tails(eng, n_rays, S):
    """rays + z interface used by the render path; sizes straddle the 128-point tile and the
    4-tile CTA-pair iteration."""
    from nwx import engine as E
    g = torch.Generator().manual_seed(n_rays * 1000 + S)
    poses = orc.synthetic_poses(1, 3)
    fx, fy, cx, cy = orc.intrinsics(24, 32)
    rays = orc.create_rays(1, poses, 24, 32, fx, fy, cx, cy, 0.1, 10.0)[0][torch.randperm(768, generator=g)[:n_rays]]
    z = torch.sort(torch.rand(n_rays, S, generator=g) * 9.9 + 0.1, -1)[0]
    pts = rays[:, None, 0:3] + rays[:, None, 3:6] * z[..., None]
    for which, sd, scale in ((E.COARSE, orc.init_state_dict(0), 1.0), (E.FINE, orc.init_state_dict(7, trained_like=True), 10.0)):
        pe3 = orc.positional_encoding(pts.reshape(-1, 3), 10, 10)
        pe2 = orc.positional_encoding(rays[:, None, 8:11].expand(n_rays, S, 3).reshape(-1, 3), 4, 1)
        emu = orc.mlp_forward_bf16_emul(sd, pe3, pe2).reshape(n_rays, S, 4)
        ref = orc.mlp_forward(sd, torch.cat([pe3, pe2], -1)).reshape(n_rays, S, 4)
        raw = eng.mlp_forward(which, rays.to(DEV), z.to(DEV)).cpu()
        assert torch.isfinite(raw).all()
        assert float((raw - emu).abs().max()) <= TOL_EMUL_MAX * scale * 3, (which, float((raw - emu).abs().max()))
        assert float((raw - ref).abs().max()) <= TOL_FP32_MAX * scale * 4


def test_mlp_large_batch_matches_small_batches(eng):
    """Size-independent property at full-frame scale: results do not depend on how points are
    tiled over CTAs (2.4 M points vs the same points in ragged slices), bit for bit."""
    from nwx import engine as E
    g = torch.Generator().manual_seed(2)
    n = 12800
    rays = torch.randn(n, 11, generator=g)
    rays[:, 8:11] = torch.nn.functional.normalize(rays[:, 3:6], dim=-1)
    rays, z = rays.to(DEV), torch.sort(torch.rand(n, 192, generator=g) * 9.9 + 0.1, -1)[0].to(DEV)
    whole = eng.mlp_forward(E.COARSE, rays, z)
    cuts = [0, 1, 130, 4097, 9000, n]
    parts = torch.cat([eng.mlp_forward(E.COARSE, rays[a:b], z[a:b]) for a, b in zip(cuts[:-1], cuts[1:])], 0)
    assert torch.equal(whole, parts)
