"""CPU: the oracle against the golden vectors frozen from the reference's own outputs
(tests/golden/make_golden.py).  Bit-exact: the oracle restates the reference op for op."""
import zlib

import numpy as np
import torch

from conftest import bits_equal, load_golden
from oracle import nerf_oracle as orc


def test_weight_init_matches_reference_constructor():
    g = load_golden("weights_crc")
    for seed, kw in ((0, {}), (1, {}), (7, {"trained_like": True})):
        sd = orc.init_state_dict(seed, **kw)
        flat = torch.cat([sd[k].reshape(-1) for k in orc.STATE_KEYS])
        assert flat.numel() == 595844
        assert zlib.crc32(flat.numpy().tobytes()) == int(g[f"seed{seed}"][0])


def test_poses():
    g = load_golden("poses36")
    assert torch.allclose(orc.synthetic_poses(36, 0), g["poses"], atol=1e-6, rtol=0)


def test_create_rays():
    g = load_golden("rays")
    rays = orc.create_rays(2, g["c2w"], g["H"], g["W"], g["fx"], g["fy"], g["cx"], g["cy"], g["near"], g["far"])
    assert bits_equal(rays, g["rays"])
    # ray index = row * W + col: the pixel at (row, col) looks along R @ ((col-cx)/fx, (row-cy)/fy, 1)
    row, col, W = 5, 11, g["W"]
    d_cam = torch.tensor([(col - g["cx"]) / g["fx"], (row - g["cy"]) / g["fy"], 1.0])
    assert torch.allclose(rays[0, row * W + col, 3:6], g["c2w"][0, :3, :3] @ d_cam, atol=1e-6)


def test_sample_pdf_and_indices():
    g = load_golden("sample_pdf")
    s, i = orc.sample_pdf(g["bins"], g["weights"], 128, det=True, return_inds=True)
    assert bits_equal(s, g["det_samples"]) and torch.equal(i, g["det_inds"])
    s, i = orc.sample_pdf(g["bins"], g["weights"], 128, det=False, u=g["u"], return_inds=True)
    assert bits_equal(s, g["rand_samples"]) and torch.equal(i, g["rand_inds"])
    assert bits_equal(orc.pdf_to_cdf(g["weights"]), g["cdf"])
    assert int(i.min()) >= 1 and int(i.max()) <= 63


def test_embedding_and_mlp():
    g = load_golden("mlp")
    pe3 = orc.positional_encoding(g["pts"], 10, 10)
    pe2 = orc.positional_encoding(g["dirs"], 4, 1)
    assert bits_equal(pe3, g["pe_xyz"]) and bits_equal(pe2, g["pe_dir"])
    raw = orc.mlp_forward(orc.init_state_dict(g["seed"]), torch.cat([pe3, pe2], -1))
    assert bits_equal(raw, g["raw"])
    emul = orc.mlp_forward_bf16_emul(orc.init_state_dict(g["seed"]), pe3, pe2)
    assert float((emul - raw).abs().max()) < 2e-2        # bf16 operands vs fp32: small, not zero
    assert float((emul - raw).abs().max()) > 0
    # the folded views layer (production inference kernel) is the same function: no further from the
    # reference's fp32 output than the layer-by-layer bf16 model is
    unfolded = orc.mlp_forward_bf16_emul(orc.init_state_dict(g["seed"]), pe3, pe2, fold_feature=False)
    assert float((emul - raw).abs().max()) <= 1.25 * float((unfolded - raw).abs().max())
    assert float((emul - unfolded).abs().max()) < 1e-3
    sd = orc.init_state_dict(7, trained_like=True)
    ref = orc.mlp_forward(sd, torch.cat([pe3, pe2], -1))
    e_fold = float((orc.mlp_forward_bf16_emul(sd, pe3, pe2) - ref).abs().max())
    e_plain = float((orc.mlp_forward_bf16_emul(sd, pe3, pe2, fold_feature=False) - ref).abs().max())
    assert e_fold <= 1.25 * e_plain, (e_fold, e_plain)


def test_raw2outputs():
    g = load_golden("raw2outputs")
    for tag, std, wb in (("plain", 0.0, False), ("white", 0.0, True), ("noise", 1.0, False)):
        out = orc.raw2outputs(g["raw"], g["z_vals"], g["rays_d"], std, wb,
                              noise=g["noise"] * std if std > 0 else None)
        for name, val in zip(("rgb", "disp", "acc", "weights", "depth"), out):
            assert bits_equal(val, g[f"{tag}_{name}"]), (tag, name)


def _nets():
    gen = torch.Generator().manual_seed(0)     # the handler builds coarse then fine from one stream
    return orc.init_state_dict(0, generator=gen), orc.init_state_dict(0, generator=gen)


def test_volumetric_rendering_inference():
    g = load_golden("render_infer")
    sd_c, sd_f = _nets()
    with torch.no_grad():
        out = orc.volumetric_rendering(g["rays"], sd_c, sd_f, orc.RenderConfig(), train_mode=False)
    for k in orc.REFERENCE_KEYS + ("z_vals_coarse", "weights_coarse", "z_samples", "inds", "z_vals_fine"):
        assert bits_equal(out[k], g[k]), k


def test_volumetric_rendering_training_and_grads():
    g = load_golden("render_train")
    gg = load_golden("train_grads")
    sd_c, sd_f = _nets()
    lc, lf, gc, gf, out = orc.training_loss_and_grads(g["rays"], g["gt"], sd_c, sd_f, orc.RenderConfig(),
                                                      g["t_rand"], g["u"], g["noise_c"], g["noise_f"])
    for k in ("rgb_coarse", "rgb_fine", "depth_fine", "acc_fine", "z_std", "z_vals_coarse", "z_samples", "inds",
              "z_vals_fine"):
        assert bits_equal(out[k].detach(), g[k]), k
    assert float(lc) == g["loss_c"] and float(lf) == g["loss_f"]        # fp64 losses, exact
    for tag, grads in (("c", gc), ("f", gf)):
        for k, v in grads.items():
            # weight-gradient GEMMs reduce over rays: their summation order depends on the BLAS
            # thread count, so these are compared to 1e-5 of the tensor's scale, not bitwise
            sub, ref = v.reshape(-1)[::97], gg[f"g{tag}.{k}.sub"]
            assert float((sub - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-12, k
            assert abs(float(v.double().norm()) - gg[f"g{tag}.{k}.norm"]) <= 1e-5 * gg[f"g{tag}.{k}.norm"] + 1e-12


def test_adam_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(1000); g1 = torch.randn(1000); g2 = torch.randn(1000)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=5e-4)
    mine, m, v = p.clone(), torch.zeros(1000), torch.zeros(1000)
    for step, g in enumerate((g1, g2), 1):
        ref.grad = g.clone(); opt.step()
        orc.adam_step(mine, g, m, v, step, 5e-4)
    assert torch.allclose(mine, ref.detach(), atol=1e-7, rtol=1e-6)
    assert abs(orc.lr_at(50000) - 5e-5) < 1e-12
