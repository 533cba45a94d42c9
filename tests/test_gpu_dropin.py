"""GPU: the function-level drop-in (the reference handler's own call sequence on nwx's entry points),
the host path around the render (uint8 written by the compositing kernel, pinned read-back) and the
training step's host behaviour (device-side batch sampling, zero synchronisations per step, per-rank
random streams, RNG state in checkpoints)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import nerf_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DEPTH_RANGE = 9.9


def _nets():
    gen = torch.Generator().manual_seed(0)
    return orc.init_state_dict(0, generator=gen), orc.init_state_dict(0, generator=gen)


def test_reference_call_sequence_on_nwx_entry_points_is_fused_and_matches_golden():
    """inference handler:203-277 restated call by call (nwx.reference_patch.reference_volumetric_rendering):
    run_network(coarse module) -> raw2outputs -> sample_pdf -> torch.sort(cat) -> run_network(the handler's
    LAMBDA around the fine module, :248) -> raw2outputs.  Must (a) equal the unmodified reference handler's
    golden output within the north-star tolerances and (b) stay on the fused kernels: 7 libnwx launches
    (2 x (dirbias + MLP), 2 x composite, sample_pdf), no embedding materialised, no chunk loop."""
    import nwx
    from nwx.reference_patch import ReferenceStyleRenderer
    g = load_golden("render_infer")
    r = ReferenceStyleRenderer(*_nets(), device=torch.device(DEV))
    rays = g["rays"].to(DEV)
    r._volumetric_rendering(rays)                                   # first call packs the weights
    torch.cuda.synchronize()
    before = nwx.engine.launch_count()
    out = r._volumetric_rendering(rays)
    torch.cuda.synchronize()
    launches = nwx.engine.launch_count() - before
    assert launches == 7, launches
    assert tuple(out.keys()) == orc.REFERENCE_KEYS
    errs = {}
    for k in ("rgb_coarse", "rgb_fine", "acc_coarse", "acc_fine"):
        errs[k] = float((out[k].cpu() - g[k]).abs().max())
        assert errs[k] <= 1e-3, (k, errs[k])
    for k in ("depth_coarse", "depth_fine"):
        errs[k] = float((out[k].cpu() - g[k]).abs().max())
        assert errs[k] <= 1e-3 * DEPTH_RANGE, (k, errs[k])
    assert float((out["raw_fine"].cpu() - g["raw_fine"]).abs().max()) <= 1.5e-3
    assert float((out["z_std"].cpu() - g["z_std"]).abs().max()) <= 2e-3
    print("function-level path vs reference handler golden:", {k: f"{v:.2e}" for k, v in errs.items()})
    # the same rays through the engine's own fused sequence: identical MLP inputs -> identical images
    h = nwx.NeRFReplicaInferenceHandler("office_tokyo", None)
    h.load_state_dicts(*_nets())
    eng_out = h._volumetric_rendering(rays)
    assert torch.equal(out["rgb_coarse"], eng_out["rgb_coarse"])
    assert float((out["rgb_fine"] - eng_out["rgb_fine"]).abs().max()) <= 1e-6


def test_literal_path_still_serves_arbitrary_callables():
    """A callable that is not a pass-through around a NeRFModel takes the literal path (embed kernels + fn per
    chunk) and gives the same numbers as the fused one, to the embedding's rounding."""
    import nwx
    sd_c, _ = _nets()
    net = nwx.NeRFModel(8, 256, 63, 27, 5, use_view_dirs=True).to(DEV)
    net.load_state_dict(sd_c)
    e3, e2 = nwx.Embedding(10, 10).embed, nwx.Embedding(4, 1).embed
    g = load_golden("render_infer")
    rays = g["rays"].to(DEV)
    n = rays.shape[0]
    z = torch.linspace(0.1, 10.0, 16, device=DEV).expand(n, 16)
    pts = rays[:, None, 0:3] + rays[:, None, 3:6] * z[..., None]
    with torch.no_grad():
        fused = nwx.run_network(pts, rays[:, 8:11], lambda x: net(x, False), e3, e2, 1024)
        literal = nwx.run_network(pts, rays[:, 8:11], lambda x: net(x) * 1.0, e3, e2, 1024)
    assert fused.shape == literal.shape == (n, 16, 4)
    assert float((fused - literal).abs().max()) <= 5e-3      # MUFU vs sinf features, re-rounded to bf16


def test_backward_through_the_forward_only_entry_points_raises():
    import nwx
    sd_c, _ = _nets()
    net = nwx.NeRFModel(8, 256, 63, 27, 5, use_view_dirs=True).to(DEV)
    net.load_state_dict(sd_c)
    x = torch.randn(32, 90, device=DEV)
    out = net(x)                                                     # grad mode on, parameters require grad
    assert out.requires_grad
    with pytest.raises(nwx.NwxError, match="Trainer"):
        out.sum().backward()
    with torch.no_grad():
        assert not net(x).requires_grad


def test_uint8_from_compositing_kernel_and_pinned_readback():
    """f4: the uint8 image is written by composite_fwd itself (no to8b launch, 8 launches per chunk + raygen)
    and equals to8b_np(rgb_fine); render_poses reads back through a persistent pinned buffer and returns a
    fresh array every call."""
    import nwx
    h = nwx.NeRFReplicaInferenceHandler("office_tokyo", None)
    H, W = 24, 32
    h._img_h, h._img_w, h._n_pix = H, W, H * W
    fx, fy, cx, cy = orc.intrinsics(H, W)
    h._fx = h._fy = fx
    h._cx, h._cy = cx, cy
    h.load_state_dicts(*_nets())
    pose = orc.synthetic_poses(36, 0)[3:5]
    h.render_poses(pose)
    torch.cuda.synchronize()
    before = nwx.engine.launch_count()
    a = h.render_poses(pose)
    assert nwx.engine.launch_count() - before == 9                  # raygen + 8
    buf = h._host_frames
    b = h.render_poses(pose)
    assert h._host_frames is buf and buf.is_pinned()                 # persistent, page-locked
    assert a is not b and np.array_equal(a, b) and a.shape == (2, H, W, 3) and a.dtype == np.uint8
    rays = nwx.create_rays(2, pose, H, W, fx, fy, cx, cy, 0.1, 10.0).view(-1, 11)
    both = h.engine.render_rays(rays, want=("rgb_fine", "rgb8_fine"))
    assert np.array_equal(both["rgb8_fine"].cpu().numpy().reshape(2, H, W, 3), a)
    assert np.array_equal(orc.to8b(both["rgb_fine"].cpu().numpy()).reshape(2, H, W, 3), a)
    only8 = h.engine.render_rays(rays, want=("rgb8_fine",))
    assert "rgb_fine" not in only8 and torch.equal(only8["rgb8_fine"], both["rgb8_fine"])


def _bank(n_img=3, H=24, W=32):
    import nwx
    fx, fy, cx, cy = orc.intrinsics(H, W)
    rays = nwx.create_rays(n_img, orc.synthetic_poses(n_img, 1), H, W, fx, fy, cx, cy, 0.1, 10.0)
    rgbs = torch.rand(n_img, H, W, 3, generator=torch.Generator().manual_seed(8))
    return rays, rgbs


def test_device_side_batch_sampling():
    """training handler:341-370: one image, n pixels with replacement -- drawn and gathered in one kernel."""
    import nwx
    rays, rgbs = _bank()
    h = nwx.NeRFReplicaTrainingHandler("office_tokyo", None, rays, rgbs)
    r, gt, idx = h._sample_training_data(want_indices=True)
    img, pix = int(idx[0]), idx[1:]
    assert 0 <= img < 3 and int(pix.min()) >= 0 and int(pix.max()) < 24 * 32 and r.shape == (1024, 11)
    assert torch.equal(r, h.rays_train[img, pix]) and torch.equal(gt, h._train_rgbs[img, pix])
    assert len(torch.unique(pix)) > 500                              # spread over the image, with replacement
    r2, _, idx2 = h._sample_training_data(want_indices=True)
    assert torch.equal(idx, idx2)                                    # same (seed, draw counter) -> same batch
    h.trainer.draws += 1
    assert not torch.equal(idx, h._sample_training_data(want_indices=True)[2])
    imgs = set()
    for d in range(40):
        imgs.add(int(h._engine.sample_training_batch(h.rays_train, h._train_rgbs, 4, 0, d, want_indices=True)[2][0]))
    assert imgs == {0, 1, 2}                                         # every image gets drawn
    other = nwx.NeRFReplicaTrainingHandler("office_tokyo", None, rays, rgbs, seed=1)    # what rank 1 would use
    assert not torch.equal(idx, other._sample_training_data(want_indices=True)[2])


def test_training_step_does_not_synchronise_the_host():
    """Zero host synchronisations per step: no .item(), no int(tensor), no pageable copies.  torch's sync
    debug mode turns any synchronising torch call into an error; libnwx itself never synchronises after its
    scratch is sized (first step)."""
    import nwx
    rays, rgbs = _bank()
    h = nwx.NeRFReplicaTrainingHandler("office_tokyo", None, rays, rgbs)
    h.step(0)                                                        # sizes the scratch (cudaMalloc)
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        for i in range(1, 4):
            out = h.step(i)
    finally:
        torch.cuda.set_sync_debug_mode("default")
    assert bool(torch.isfinite(out["total_loss"])) and out["total_loss"].is_cuda
    losses = [h.step(i)["total_loss"] for i in range(4, 7)]         # returned losses are copies, not one buffer
    assert len({float(l) for l in losses}) == 3


def test_checkpoint_carries_the_rng_state():
    import nwx
    eng = nwx.Engine(torch.device(DEV))
    tr = nwx.Trainer(eng, *_nets(), seed=5)
    g = load_golden("render_train")
    rays, gt = g["rays"].to(DEV), g["gt"].float().to(DEV)
    for s in range(3):
        tr.step(rays, gt, s)
    ck = tr.checkpoint(3)
    assert ck["nwx_rng"] == {"seed": 5, "draws": 3}
    tr2 = nwx.Trainer(nwx.Engine(torch.device(DEV)), *_nets(), seed=5)
    assert tr2.load_checkpoint(ck) == 3 and tr2.draws == 3
    a, b = tr.step(rays, gt, 3), tr2.step(rays, gt, 3)              # same draws -> same step, bit for bit
    assert torch.equal(a, b) and torch.equal(tr.params, tr2.params)
    del ck["nwx_rng"]                                                # a reference checkpoint: best-effort offset
    tr3 = nwx.Trainer(nwx.Engine(torch.device(DEV)), *_nets(), seed=5)
    tr3.load_checkpoint(ck)
    assert tr3.draws == 3


def test_two_training_contexts_on_two_streams_do_not_race():
    """the training kernels read biases from process-global constant banks; two contexts on two
    streams are chained by an event instead of racing.  Interleaved steps must equal isolated steps."""
    import nwx
    g = load_golden("render_train")
    rays, gt = g["rays"].to(DEV), g["gt"].float().to(DEV)
    sd_a = _nets()
    gen = torch.Generator().manual_seed(9)
    sd_b = (orc.init_state_dict(9, generator=gen), orc.init_state_dict(9, generator=gen))

    def solo(sd):
        tr = nwx.Trainer(nwx.Engine(torch.device(DEV)), *sd, seed=3)
        return [tr.step(rays, gt, s).cpu() for s in range(3)], tr.params.clone()
    la, pa = solo(sd_a)
    lb, pb = solo(sd_b)
    ta = nwx.Trainer(nwx.Engine(torch.device(DEV)), *sd_a, seed=3)
    tb = nwx.Trainer(nwx.Engine(torch.device(DEV)), *sd_b, seed=3)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    got_a, got_b = [], []
    for s in range(3):
        with torch.cuda.stream(s1):
            got_a.append(ta.step(rays, gt, s))
        with torch.cuda.stream(s2):
            got_b.append(tb.step(rays, gt, s))
    torch.cuda.synchronize()
    assert all(torch.equal(x.cpu(), y) for x, y in zip(got_a, la)) and torch.equal(ta.params, pa)
    assert all(torch.equal(x.cpu(), y) for x, y in zip(got_b, lb)) and torch.equal(tb.params, pb)


def test_cuda_graph_replay_equals_direct_launches():
    """render_poses replays a captured launch sequence per frame shape: same bytes as the direct launches for
    every pose, re-captured when the weights change, counted in launch_count()."""
    import nwx
    H, W = 24, 32
    fx, fy, cx, cy = orc.intrinsics(H, W)

    def make(graphs):
        h = nwx.NeRFReplicaInferenceHandler("office_tokyo", None)
        h._img_h, h._img_w, h._n_pix, h._fx, h._fy, h._cx, h._cy = H, W, H * W, fx, fy, cx, cy
        h.use_cuda_graphs = graphs
        h.load_state_dicts(*_nets())
        return h
    hg, he = make(True), make(False)
    poses = orc.synthetic_poses(36, 0)
    for i in (0, 5, 9):
        a, b = hg.render_poses(poses[i:i + 1]), he.render_poses(poses[i:i + 1])
        assert np.array_equal(a, b), i
    assert len(hg._graphs) == 1 and len(he._graphs) == 0
    before = nwx.engine.launch_count()
    hg.render_poses(poses[3:4])
    assert nwx.engine.launch_count() - before == 9                  # replayed kernels are counted
    two = hg.render_poses(poses[2:4])                                # another shape: another graph
    assert len(hg._graphs) == 2 and np.array_equal(two, he.render_poses(poses[2:4]))
    gen = torch.Generator().manual_seed(4)
    other = (orc.init_state_dict(4, generator=gen), orc.init_state_dict(4, generator=gen))
    hg.load_state_dicts(*other); he.load_state_dicts(*other)         # new biases: the old graphs must not survive
    assert len(hg._graphs) == 0
    c, d = hg.render_poses(poses[5:6]), he.render_poses(poses[5:6])
    assert np.array_equal(c, d) and not np.array_equal(c, b)
    # a larger eager render re-allocates the scratch underneath the captured graph: it must notice and re-capture
    big = hg.engine.raygen(poses[:4], 48, 64, *orc.intrinsics(48, 64), 0.1, 10.0)
    hg.engine.render_rays(big, want=("rgb_fine",))
    assert np.array_equal(hg.render_poses(poses[5:6]), d)
