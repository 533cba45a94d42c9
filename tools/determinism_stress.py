#!/usr/bin/env python
"""Determinism stress: every kernel of the frame is a pure function of its inputs, so repeated launches on the same
inputs must give the same bits.  Runs each HBM-side kernel and the whole 640x480 render `--iters` times and counts
launches whose output differs from the first one (a race shows up as a non-zero count).  One JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nerf-workspaces-explorer_b200"))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=100)
    args = ap.parse_args()
    import nwx
    from nwx import engine as E
    from nwx import synthetic
    dev = torch.device("cuda", 0)
    eng = nwx.Engine(dev)
    sd_c, sd_f = synthetic.random_state_dicts(0)
    eng.load_weights(E.COARSE, sd_c); eng.load_weights(E.FINE, sd_f)
    H, W = 480, 640
    fx, fy, cx, cy = synthetic.intrinsics(H, W)
    rays = eng.raygen(synthetic.sweep_poses(36, 0)[:1], H, W, fx, fy, cx, cy, 0.1, 10.0)
    out = eng.render_rays(rays, want=("raw_coarse", "raw_fine", "z_vals_coarse", "z_vals_fine", "weights_coarse", "rgb_fine"))
    raw_c, raw_f, z_c, z_f, w_c = (out[k].clone() for k in ("raw_coarse", "raw_fine", "z_vals_coarse", "z_vals_fine", "weights_coarse"))
    rays_d = rays[:, 3:6].contiguous()
    cases = {
        "composite_coarse": lambda: torch.cat([t.reshape(-1) for t in E.composite(raw_c, z_c, rays_d, want_weights=True) if t is not None]),
        "composite_fine": lambda: torch.cat([t.reshape(-1) for t in E.composite(raw_f, z_f, rays_d, want_weights=False) if t is not None]),
        "sample_pdf": lambda: torch.cat([t.reshape(-1) for t in E.sample_pdf_merge(z_c, w_c, 128, want_inds=False)[:2]]),
        "mlp_coarse": lambda: eng.mlp_forward(E.COARSE, rays, z_c).reshape(-1),
        "render_rays_rgb": lambda: eng.render_rays(rays, want=("rgb_fine",))["rgb_fine"].reshape(-1),
    }
    # the shard shapes of tests/test_gpu_render.py::test_full_frame_properties_640x480 (row tiles + a ragged cut):
    # small and odd launches exercise the kernels' tails; each must reproduce its slice of the full frame
    full = eng.render_rays(rays, want=("rgb_fine",))["rgb_fine"].clone()
    cuts = [0, 100 * W, 100 * W + 77, 300 * W, H * W]

    def shards():
        parts = [eng.render_rays(rays[a:b], want=("rgb_fine",))["rgb_fine"] for a, b in zip(cuts[:-1], cuts[1:])]
        return torch.cat(parts, 0).reshape(-1)
    cases["sharded_render_vs_full_frame"] = shards
    # the full set of outputs the render test asks for (weights, merged depths, maps, uint8 pixels), all compared
    want = ("rgb_fine", "acc_fine", "depth_fine", "z_vals_fine", "weights_fine", "rgb8_fine")

    def full_outputs():
        o = eng.render_rays(rays, want=want)
        return torch.cat([o[k].reshape(-1).view(torch.uint8) for k in want])
    cases["render_rays_all_outputs"] = full_outputs
    res = {}
    for name, fn in cases.items():
        ref = full.reshape(-1) if name.startswith("sharded") else fn().clone()
        bad, worst = 0, 0
        for _ in range(args.iters):
            cur = fn()
            diff = int((cur != ref).sum()) if cur.dtype == torch.uint8 else int((cur.view(torch.int32) != ref.view(torch.int32)).sum())
            if diff:
                bad += 1
                worst = max(worst, diff)
        res[name] = {"launches": args.iters, "launches_that_differ": bad, "max_differing_words": worst}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
