"""GPU parity of the streaming kernels (K1 raygen/coarse-z, K2 sample_pdf+merge, K4 composite)
against the golden vectors of the reference and against the oracle on seeded inputs.
All calls go through the C ABI (nwx.engine -> libnwx.so)."""
import pytest
import torch

from conftest import bits_equal, load_golden
from oracle import nerf_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def eng():
    import nwx
    return nwx.Engine(torch.device(DEV))


def test_native_library_is_loaded():
    import nwx
    assert nwx.lib().nwx_version() == 100
    assert any("libnwx.so" in line for line in open("/proc/self/maps"))


# ---------------------------------------------------------------- K1 ----
def test_raygen_bit_exact_vs_reference_golden():
    import nwx
    g = load_golden("rays")
    rays = nwx.create_rays(2, g["c2w"], g["H"], g["W"], g["fx"], g["fy"], g["cx"], g["cy"], g["near"], g["far"], True)
    assert rays.is_cuda and bits_equal(rays.cpu(), g["rays"])
    rays8 = nwx.create_rays(2, g["c2w"], g["H"], g["W"], g["fx"], g["fy"], g["cx"], g["cy"], g["near"], g["far"], False)
    assert bits_equal(rays8.cpu(), g["rays"][..., :8].contiguous())


def test_raygen_full_frame_and_ranges(eng):
    poses = orc.synthetic_poses(3, 0)
    H, W = 480, 640
    fx, fy, cx, cy = orc.intrinsics(H, W)
    ref = orc.create_rays(3, poses, H, W, fx, fy, cx, cy, 0.1, 10.0, True).reshape(-1, 11)
    got = eng.raygen(poses, H, W, fx, fy, cx, cy, 0.1, 10.0, True)
    assert bits_equal(got.cpu(), ref)                       # 921 600 rays, bit for bit
    part = eng.raygen(poses, H, W, fx, fy, cx, cy, 0.1, 10.0, True, ray0=300001, nrays=12345)   # ragged shard
    assert bits_equal(part.cpu(), ref[300001:300001 + 12345].contiguous())
    assert eng.raygen(poses, H, W, fx, fy, cx, cy, 0.1, 10.0, True, ray0=5, nrays=0).shape == (0, 11)


def test_coarse_z_bit_exact(eng):
    g = torch.Generator().manual_seed(3)
    rays = torch.randn(1000, 11, generator=g)
    rays[:, 6], rays[:, 7] = 0.1, 10.0
    rays[::7, 6], rays[::7, 7] = 0.35, 4.2
    t_rand = torch.rand(1000, 64, generator=g)
    for tr in (None, t_rand):
        ref = orc.coarse_z(rays, 64, tr)
        got = eng.coarse_z(rays.to(DEV), 64, None if tr is None else tr.to(DEV))
        assert bits_equal(got.cpu(), ref.contiguous())
    for S in (2, 6, 8, 33, 100):               # S % 4 == 0 takes the 128-bit-store kernel, the others the scalar one
        assert bits_equal(eng.coarse_z(rays.to(DEV), S, None).cpu(), orc.coarse_z(rays, S, None).contiguous()), S
        assert bits_equal(eng.coarse_z(rays[:1].to(DEV), S, None).cpu(), orc.coarse_z(rays[:1], S, None).contiguous()), S


def test_embed_and_to8b():
    import nwx
    g = load_golden("mlp")
    pe3 = nwx.Embedding(10, 10).embed(g["pts"].to(DEV)).cpu()
    pe2 = nwx.Embedding(4, 1).embed(g["dirs"].to(DEV)).cpu()
    assert pe3.shape == (160, 63) and pe2.shape == (160, 27)
    assert float((pe3 - g["pe_xyz"]).abs().max()) <= 2e-6 and float((pe2 - g["pe_dir"]).abs().max()) <= 5e-7
    x = torch.linspace(-0.5, 1.5, 10001)
    x[5] = float("nan")
    assert torch.equal(nwx.to8b(x.to(DEV)).cpu(), torch.from_numpy(orc.to8b(x.numpy())))


# ---------------------------------------------------------------- K2 ----
def test_sample_pdf_bit_exact_vs_reference_golden():
    import nwx
    from nwx import engine as E
    g = load_golden("sample_pdf")
    bins, w, u = g["bins"].to(DEV), g["weights"].to(DEV), g["u"].to(DEV)
    s, i, cdf = E.sample_pdf_bins(bins, w, 128, None, want_inds=True, want_cdf=True)
    assert bits_equal(cdf.cpu(), g["cdf"])                   # torch.sum / torch.cumsum order reproduced
    assert torch.equal(i.cpu(), g["det_inds"])               # searchsorted indices: exact
    assert bits_equal(s.cpu(), g["det_samples"])
    s, i, _ = E.sample_pdf_bins(bins, w, 128, u, want_inds=True)
    assert torch.equal(i.cpu(), g["rand_inds"]) and bits_equal(s.cpu(), g["rand_samples"])
    # the reference signature
    assert bits_equal(nwx.sample_pdf(bins, w, 128, det=True).cpu(), g["det_samples"])
    assert bits_equal(nwx.sample_pdf(bins, w, 128, det=False, u=u).cpu(), g["rand_samples"])
    r = nwx.sample_pdf(bins, w, 128, det=False)
    assert r.shape == (192, 128) and float(r.min()) >= float(bins.min()) and float(r.max()) <= float(bins.max())


@pytest.mark.parametrize("n_rays", [1, 33, 20000])
def test_sample_pdf_merge_vs_oracle(n_rays):
    from nwx import engine as E
    g = torch.Generator().manual_seed(n_rays)
    rays = torch.zeros(n_rays, 8); rays[:, 6], rays[:, 7] = 0.1, 10.0
    z_c = orc.coarse_z(rays, 64, torch.rand(n_rays, 64, generator=g)).contiguous()
    w_c = torch.rand(n_rays, 64, generator=g) ** 6
    w_c[: n_rays // 10] = 0.0
    u = torch.rand(n_rays, 128, generator=g)
    z_mid = .5 * (z_c[..., 1:] + z_c[..., :-1])
    for uu in (None, u):
        ref_s, ref_i = orc.sample_pdf(z_mid, w_c[..., 1:-1], 128, det=uu is None, u=uu, return_inds=True)
        ref_f = torch.sort(torch.cat([z_c, ref_s], -1), -1)[0]
        z_s, z_f, inds, z_std = E.sample_pdf_merge(z_c.to(DEV), w_c.to(DEV), 128, None if uu is None else uu.to(DEV))
        assert torch.equal(inds.cpu(), ref_i)
        assert bits_equal(z_s.cpu(), ref_s) and bits_equal(z_f.cpu(), ref_f)   # merge == torch.sort values
        ref_std = torch.std(ref_s, dim=-1, unbiased=False)
        assert torch.allclose(z_std.cpu(), ref_std, rtol=2e-5, atol=1e-6)
        assert bool((z_f[:, 1:] >= z_f[:, :-1]).all())


# ---------------------------------------------------------------- K4 ----
def test_composite_vs_reference_golden():
    import nwx
    g = load_golden("raw2outputs")
    raw, z, d, noise = (g[k].to(DEV) for k in ("raw", "z_vals", "rays_d", "noise"))
    from nwx import engine as E
    for tag, nz, wb in (("plain", None, False), ("white", None, True), ("noise", noise, False)):
        flags = torch.zeros(1, dtype=torch.int32, device=DEV)
        out = E.composite(raw, z, d, nz, wb, flags=flags)
        for name, val in zip(("rgb", "disp", "acc", "weights", "depth"), out):
            ref = g[f"{tag}_{name}"]
            val = val.cpu()
            assert torch.equal(torch.isnan(val), torch.isnan(ref)), (tag, name)     # empty rays: disp = NaN
            ok = ~torch.isnan(ref)
            scale = 1.0 if name != "disp" else ref[ok].abs().clamp(min=1.0)
            err = ((val[ok] - ref[ok]).abs() / scale).max()
            assert float(err) <= 2e-6, (tag, name, float(err))
        has_nan = any(bool(torch.isnan(g[f"{tag}_{n}"]).any()) for n in ("rgb", "disp", "acc", "depth"))
        assert bool(int(flags.item()) & 1) == has_nan                            # device NaN flag (disp of empty rays)
    # weights: bit-exact fraction is reported by the probe; here the reference signature
    rgb, disp, acc, weights, depth, feat = nwx.raw2outputs(raw, z, d, 0, False)
    assert float((weights.cpu() - g["plain_weights"]).abs().max()) <= 1e-7 and feat.item() == 0


@pytest.mark.parametrize("S", [64, 192, 100])
def test_composite_shapes_vs_oracle(S):
    from nwx import engine as E
    g = torch.Generator().manual_seed(S)
    N = 3000
    raw = torch.randn(N, S, 4, generator=g) * 3
    z = torch.sort(torch.rand(N, S, generator=g) * 9.9 + 0.1, -1)[0]
    d = torch.randn(N, 3, generator=g)
    ref = orc.raw2outputs(raw, z, d)
    out = E.composite(raw.to(DEV), z.to(DEV), d.to(DEV))
    for name, a, b in zip(("rgb", "disp", "acc", "weights", "depth"), out, ref):
        if name == "disp":
            continue
        assert float((a.cpu() - b).abs().max()) <= 5e-6, name
    assert float((out[3].cpu().sum(-1) - out[2].cpu()).abs().max()) <= 1e-5     # acc == sum of weights


def test_composite_backward_vs_autograd():
    from nwx import engine as E
    g = torch.Generator().manual_seed(9)
    for S, wb in ((64, False), (192, False), (64, True)):
        N = 500
        raw = (torch.randn(N, S, 4, generator=g) * 2).requires_grad_(True)
        z = torch.sort(torch.rand(N, S, generator=g) * 9.9 + 0.1, -1)[0]
        d = torch.randn(N, 3, generator=g)
        noise = torch.randn(N, S, generator=g)
        d_rgb = torch.randn(N, 3, generator=g)
        rgb = orc.raw2outputs(raw, z, d, 1.0, wb, noise=noise)[0]
        rgb.backward(d_rgb)
        got = E.composite_backward(raw.detach().to(DEV), z.to(DEV), d.to(DEV), d_rgb.to(DEV), noise.to(DEV), wb).cpu()
        ref = raw.grad
        assert float((got - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max())), (S, wb)


@pytest.mark.parametrize("M", [12, 20, 47, 63, 100, 127])
def test_sample_pdf_bins_general_lengths_bit_exact(M):
    """The literal sample_pdf for other bin counts than the reference's 63: indices and samples stay
    bit-exact (the CDF's summation order is reproduced for every row length)."""
    from nwx import engine as E
    g = torch.Generator().manual_seed(M)
    N, n_imp = 3000, 96
    bins = torch.sort(torch.rand(N, M, generator=g) * 9.9 + 0.1, -1)[0]
    w = torch.rand(N, M - 1, generator=g) ** 5
    u = torch.rand(N, n_imp, generator=g)
    for uu in (None, u):
        ref_s, ref_i = orc.sample_pdf(bins, w, n_imp, det=uu is None, u=uu, return_inds=True)
        s, i, cdf = E.sample_pdf_bins(bins.to(DEV), w.to(DEV), n_imp, None if uu is None else uu.to(DEV),
                                      want_inds=True, want_cdf=True)
        assert bits_equal(cdf.cpu(), orc.pdf_to_cdf(w))
        assert torch.equal(i.cpu(), ref_i) and bits_equal(s.cpu(), ref_s)
