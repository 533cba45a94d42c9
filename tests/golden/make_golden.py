"""Freeze golden vectors from the *reference itself* and pin the oracle against it.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

It imports the reference's own modules from /root/reference (read-only), runs them on CPU on
seeded inputs, runs oracle/nerf_oracle.py on the same inputs, asserts BIT-EQUALITY between the
two, and writes the reference outputs to tests/golden/*.npz.  The handlers hard-code ``.cuda()``
(inference handler:111,119,174,216; model_utils.py:54,67,75), so for the end-to-end cases
``Tensor.cuda`` / ``Module.cuda`` are patched to the identity for the duration of this script --
the reference code itself runs unmodified.  ``nerf.training`` imports imageio/imgviz (absent);
empty stand-in modules are registered so the module imports, nothing from them is called.
"""
import os
import sys
import types
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

for missing in ("imageio", "imgviz"):
    sys.modules.setdefault(missing, types.ModuleType(missing))
sys.modules["imgviz"].depth2rgb = None             # imported by name, never called here
tb = types.ModuleType("torch.utils.tensorboard")
tb.SummaryWriter = object
sys.modules.setdefault("torch.utils.tensorboard", tb)

torch.Tensor.cuda = lambda self, *a, **k: self          # see module docstring
torch.nn.Module.cuda = lambda self, *a, **k: self

from nerf.rays import rays as ref_rays                                    # noqa: E402
from nerf.models.embedding import Embedding as RefEmbedding               # noqa: E402
from nerf.models.nerf_model import NeRFModel as RefNeRFModel               # noqa: E402
from nerf.models import model_utils as ref_mu                             # noqa: E402
from nerf.inference.nerf_replica_inference_handler import NeRFReplicaInferenceHandler  # noqa: E402
from utils.camera_poses import get_camera_poses_from_list_of_coordinates  # noqa: E402
from utils.data_descriptors import COORD                                   # noqa: E402

torch.autograd.set_detect_anomaly(False)  # the reference switches it on at import (nerf_model.py:7)

from oracle import nerf_oracle as orc                                      # noqa: E402


def same(a: torch.Tensor, b: torch.Tensor, what: str):
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    bits = {4: torch.int32, 8: torch.int64}[a.element_size()]      # bit compare: NaN == NaN (empty rays)
    assert torch.equal(a.contiguous().view(bits), b.contiguous().view(bits)), \
        f"oracle != reference (bitwise) for {what}: max|d|={float((a - b).abs().max())}"


def save(name: str, **arrays):
    out = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"  wrote {name}.npz ({os.path.getsize(os.path.join(HERE, name + '.npz')) / 1024:.0f} KiB)")


def ref_state_dict(seed: int, alpha_bias=0.1, trained_like=False):
    torch.manual_seed(seed)
    m = RefNeRFModel(8, 256, 63, 27, 5, use_view_dirs=True)
    with torch.no_grad():
        if trained_like:
            m._alpha_linear.weight *= 30.0
            m._rgb_linear.weight *= 10.0
            m._alpha_linear.bias.fill_(1.0)
        elif alpha_bias is not None:
            m._alpha_linear.bias.fill_(alpha_bias)
    return m, {k: v.detach().clone() for k, v in m.state_dict().items()}


def main():
    torch.set_num_threads(1)  # one thread: ATen reductions are then order-stable
    print("torch", torch.__version__, "cpu capability", torch.backends.cpu.get_cpu_capability())

    # ---- weights: oracle init == reference constructor under the same seed ----------------
    wsum = {}
    for seed, kw in ((0, {}), (1, {}), (7, {"trained_like": True})):
        _, sd_ref = ref_state_dict(seed, **kw)
        sd = orc.init_state_dict(seed, **kw)
        assert tuple(sd_ref.keys()) == orc.STATE_KEYS
        for k in sd_ref:
            same(sd[k], sd_ref[k], f"init seed {seed} {k}")
        flat = torch.cat([sd[k].reshape(-1) for k in orc.STATE_KEYS])
        wsum[str(seed)] = np.array([zlib.crc32(flat.numpy().tobytes())], dtype=np.uint32)
    save("weights_crc", **{f"seed{k}": v for k, v in wsum.items()})

    # ---- poses (boundary input producer; tolerance, cv2 vs closed form) -------------------
    rng = np.random.RandomState(0)
    x, z = rng.uniform(-2, 2), rng.uniform(-3, 1.5)
    init = COORD(x=x, y=-0.5, z=z, yaw=0.0, pitch=-90.0, roll=0.0)
    views = [COORD(yaw=-float(h), pitch=float(v)) for v in (-30, 0, 30) for h in range(0, 360, 30)]
    poses_ref = get_camera_poses_from_list_of_coordinates(init, views)
    poses = orc.synthetic_poses(36, seed=0)
    assert torch.allclose(poses, poses_ref, atol=1e-6, rtol=0), float((poses - poses_ref).abs().max())
    save("poses36", poses=poses_ref)

    # ---- create_rays ----------------------------------------------------------------------
    H, W = 12, 16
    fx, fy, cx, cy = orc.intrinsics(H, W)
    c2w = poses_ref[[3, 17]]
    r_ref = ref_rays.create_rays(2, c2w, H, W, fx, fy, cx, cy, 0.1, 10.0, True)
    same(orc.create_rays(2, c2w, H, W, fx, fy, cx, cy, 0.1, 10.0, True), r_ref, "create_rays")
    r8_ref = ref_rays.create_rays(2, c2w, H, W, fx, fy, cx, cy, 0.1, 10.0, False)
    same(orc.create_rays(2, c2w, H, W, fx, fy, cx, cy, 0.1, 10.0, False), r8_ref, "create_rays(no dirs)")
    save("rays", c2w=c2w, H=H, W=W, fx=fx, fy=fy, cx=cx, cy=cy, near=0.1, far=10.0, rays=r_ref)

    # ---- sample_pdf (samples + the searchsorted indices the reference computed) -----------
    g = torch.Generator().manual_seed(11)
    N = 192
    z = orc.coarse_z(torch.cat([torch.zeros(N, 6), torch.full((N, 1), 0.1), torch.full((N, 1), 10.0)], 1), 64)
    bins = .5 * (z[..., 1:] + z[..., :-1])
    wts = torch.rand(N, 62, generator=g) ** 4
    wts[:8] = 0.0                                   # all-zero weights -> uniform pdf
    wts[8:16, :] = 0.0
    wts[8:16, 30] = 1.0                             # one-hot pdf (denominator guards, rays.py:114)
    u_rand = torch.rand(N, 128, generator=g)
    captured = []
    real_ss = torch.searchsorted

    def spy(*a, **k):
        out = real_ss(*a, **k)
        captured.append(out.clone())
        return out

    torch.searchsorted = spy
    det_ref = ref_rays.sample_pdf(bins, wts, 128, det=True)
    real_rand = torch.rand
    torch.rand = lambda *a, **k: u_rand.clone()     # inject u into rays.py:98
    rnd_ref = ref_rays.sample_pdf(bins, wts, 128, det=False)
    torch.rand = real_rand
    torch.searchsorted = real_ss
    det_o, det_i = orc.sample_pdf(bins, wts, 128, det=True, return_inds=True)
    rnd_o, rnd_i = orc.sample_pdf(bins, wts, 128, det=False, u=u_rand, return_inds=True)
    same(det_o, det_ref, "sample_pdf det"); same(det_i, captured[0], "sample_pdf det inds")
    same(rnd_o, rnd_ref, "sample_pdf rand"); same(rnd_i, captured[1], "sample_pdf rand inds")
    save("sample_pdf", bins=bins, weights=wts, u=u_rand, det_samples=det_ref, det_inds=captured[0],
         rand_samples=rnd_ref, rand_inds=captured[1], cdf=orc.pdf_to_cdf(wts))

    # ---- embedding + MLP ------------------------------------------------------------------
    pts = (torch.rand(160, 3, generator=g) - 0.5) * 24.0
    dirs = torch.nn.functional.normalize(torch.randn(160, 3, generator=g), dim=-1)
    e3, e2 = RefEmbedding(10, 10), RefEmbedding(4, 1)
    pe3_ref, pe2_ref = e3.embed(pts), e2.embed(dirs)
    same(orc.positional_encoding(pts, 10, 10), pe3_ref, "embed xyz")
    same(orc.positional_encoding(dirs, 4, 1), pe2_ref, "embed dir")
    model, sd0 = ref_state_dict(0)
    with torch.no_grad():
        raw_ref = model(torch.cat([pe3_ref, pe2_ref], -1))
    same(orc.mlp_forward(sd0, torch.cat([pe3_ref, pe2_ref], -1)), raw_ref, "NeRFModel.forward")
    save("mlp", pts=pts, dirs=dirs, pe_xyz=pe3_ref, pe_dir=pe2_ref, raw=raw_ref, seed=0)

    # ---- raw2outputs ----------------------------------------------------------------------
    Nr = 96
    raw = torch.randn(Nr, 64, 4, generator=g) * 2.0
    raw[:4, :, 3] = -1.0                            # empty rays: acc = 0, disp hits the 1e-10 guard
    raw[4:8, :, 3] = 50.0                           # opaque at the first sample
    zv = torch.sort(torch.rand(Nr, 64, generator=g) * 9.9 + 0.1, -1)[0]
    rd = torch.randn(Nr, 3, generator=g)
    noise = torch.randn(Nr, 64, generator=g)
    real_randn = torch.randn
    packs = {}
    for tag, std, wb in (("plain", 0.0, False), ("white", 0.0, True), ("noise", 1.0, False)):
        torch.randn = lambda *a, **k: noise.clone()  # inject the draw of model_utils.py:65
        ref = ref_mu.raw2outputs(raw, zv, rd, std, wb, False, cuda_enabled=False)
        torch.randn = real_randn
        mine = orc.raw2outputs(raw, zv, rd, std, wb, noise=noise * std if std > 0 else None)
        for name, a, b in zip(("rgb", "disp", "acc", "weights", "depth"), mine, ref[:5]):
            same(a, b, f"raw2outputs[{tag}].{name}")
            packs[f"{tag}_{name}"] = b
    save("raw2outputs", raw=raw, z_vals=zv, rays_d=rd, noise=noise, **packs)

    # ---- end-to-end: the reference handler's own _volumetric_rendering, inference ----------
    handler = NeRFReplicaInferenceHandler("office_tokyo", "/nonexistent/model.ckpt")
    torch.manual_seed(0)
    try:
        handler.initialize_models()
    except RuntimeError:
        pass                                         # no checkpoint: random-init models stay in place
    with torch.no_grad():
        handler._nerf_net_coarse._alpha_linear.bias.fill_(0.1)
        handler._nerf_net_fine._alpha_linear.bias.fill_(0.1)
    sd_c = {k: v.detach().clone() for k, v in handler._nerf_net_coarse.state_dict().items()}
    sd_f = {k: v.detach().clone() for k, v in handler._nerf_net_fine.state_dict().items()}
    # the handler builds coarse then fine from one RNG stream: same as two consecutive oracle inits
    gen = torch.Generator().manual_seed(0)
    o_c = orc.init_state_dict(0, generator=gen)
    o_f = orc.init_state_dict(0, generator=gen)
    for k in sd_c:
        same(o_c[k], sd_c[k], f"handler coarse {k}"); same(o_f[k], sd_f[k], f"handler fine {k}")

    Hs, Ws = 6, 8
    fx, fy, cx, cy = orc.intrinsics(Hs, Ws)
    rays = ref_rays.create_rays(1, poses_ref[5:6], Hs, Ws, fx, fy, cx, cy, 0.1, 10.0, True)[0]
    cfg = orc.RenderConfig()
    with torch.no_grad():
        ref_out = handler._volumetric_rendering(rays)
        mine = orc.volumetric_rendering(rays, sd_c, sd_f, cfg, train_mode=False)
    assert tuple(ref_out.keys()) == orc.REFERENCE_KEYS
    for k in ref_out:
        same(mine[k], ref_out[k], f"inference _volumetric_rendering[{k}]")
    extras = {k: mine[k] for k in ("z_vals_coarse", "weights_coarse", "z_samples", "inds", "z_vals_fine")}
    save("render_infer", rays=rays, **{k: v for k, v in ref_out.items()}, **extras)

    # ---- end-to-end, training variant (jitter, sigma noise, random u) + loss/grads ---------
    from nerf.training.nerf_replica_training_handler import NeRFReplicaTrainingHandler
    th = object.__new__(NeRFReplicaTrainingHandler)
    for k, v in dict(_n_samples=64, _n_importance=128, _perturb=1.0, _train_mode=True, _raw_noise_std=1.0,
                     _white_bkgd=False, _endpoint_feat=False, _net_chunk=1024 * 32,
                     _nerf_net_coarse=handler._nerf_net_coarse, _nerf_net_fine=handler._nerf_net_fine,
                     _embed_fcn=handler._embed_fcn, _embed_dirs_fcn=handler._embed_dirs_fcn).items():
        setattr(th, k, v)
    n = rays.shape[0]
    torch.manual_seed(123)                           # draw order: t_rand, noise_c, u, noise_f
    t_rand = torch.rand(n, 64); noise_c = torch.randn(n, 64) * 1.0
    u = torch.rand(n, 128); noise_f = torch.randn(n, 192) * 1.0
    gt = torch.rand(n, 3, generator=g).double()
    torch.manual_seed(123)
    for p in list(handler._nerf_net_coarse.parameters()) + list(handler._nerf_net_fine.parameters()):
        p.grad = None
    ref_tr = th._volumetric_rendering(rays)
    loss_c = ref_mu.img2mse(ref_tr["rgb_coarse"], gt)        # training handler:291
    loss_f = ref_mu.img2mse(ref_tr["rgb_fine"], gt)          # :298
    (loss_c + loss_f).backward()                            # :305-308
    lc, lf, gc, gf, mine_tr = orc.training_loss_and_grads(rays, gt, sd_c, sd_f, cfg, t_rand, u, noise_c, noise_f)
    for k in ref_tr:
        same(mine_tr[k].detach(), ref_tr[k].detach(), f"training _volumetric_rendering[{k}]")
    same(lc, loss_c.detach(), "loss_coarse"); same(lf, loss_f.detach(), "loss_fine")
    grads = {}
    for tag, net, og in (("c", handler._nerf_net_coarse, gc), ("f", handler._nerf_net_fine, gf)):
        for k, p in net.named_parameters():
            same(og[k], p.grad, f"grad {tag} {k}")
            grads[f"g{tag}.{k}"] = p.grad
    small = {k: v.detach() for k, v in ref_tr.items() if not k.startswith("raw_")}
    save("render_train", rays=rays, gt=gt, t_rand=t_rand, u=u, noise_c=noise_c, noise_f=noise_f,
         loss_c=loss_c.detach(), loss_f=loss_f.detach(), **small,
         **{k: mine_tr[k].detach() for k in ("z_vals_coarse", "z_samples", "inds", "z_vals_fine")})
    # grads are 2 x 595 844 floats: keep per-tensor norms + a strided subsample, enough to catch any slip
    gsmall = {}
    for k, v in grads.items():
        gsmall[k + ".norm"] = v.double().norm()
        gsmall[k + ".sub"] = v.reshape(-1)[::97].clone()
    save("train_grads", **gsmall)
    # ---- the caller above the path: Workspace coordinate transforms (application/workspace.py) ----
    from application import workspace as ref_ws
    sys.path.insert(0, os.path.join(ROOT, "nerf-workspaces-explorer_b200"))
    import nwx.workspace as my_ws
    cases = [(0.0, 0.0, 0, 0), (1.0, 1.0, 30, -30), (0.25, 0.7, 330, 30), (0.5, 0.5, 90, 0)]
    packs = {}
    for cls in ("OfficeTokyoWorkspace", "OfficeNewYorkWorkspace", "OfficeGeneveWorkspace", "OfficeBelgradeWorkspace"):
        r, m = getattr(ref_ws, cls)(), getattr(my_ws, cls)()
        rows = []
        for c in cases:
            a, b = r._transform_relative_coordinates(*c), m._transform_relative_coordinates(*c)
            assert tuple(a[0]) == tuple(b[0]) and tuple(a[1]) == tuple(b[1]), (cls, c)
            rows.append(list(a[0]) + list(a[1]))
        assert r.name == m.name and tuple(r.floor_plan_scale) == tuple(m.floor_plan_scale)
        packs[cls] = np.array(rows)
    save("workspace", cases=np.array(cases, dtype=np.float64), **packs)
    print("golden vectors written; oracle == reference bit-exactly on every case")


if __name__ == "__main__":
    main()
