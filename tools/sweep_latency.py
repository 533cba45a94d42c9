#!/usr/bin/env python
"""BASELINE.json config 5: the GUI's camera sweep -- 36 yaw/pitch poses around one clicked spot
(application/app.py:389-413) at the product resolution 320x240 (office_*_config.yaml:2-3).

Measures, through the public handler API (pinned host poses in, uint8 frames out on the host):
  * per-click latency of render_coordinates-style single views (p50 / p99 over the 36 poses x reps)
  * latency of the whole sweep rendered as ONE batched launch sequence (render_poses, B = 36)
Prints one JSON line.  usage: python tools/sweep_latency.py [--reps 5] [--height 240 --width 320]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nerf-workspaces-explorer_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--height", type=int, default=240)
    ap.add_argument("--width", type=int, default=320)
    args = ap.parse_args()
    import nwx
    from nwx import synthetic
    H, W = args.height, args.width
    fx, fy, cx, cy = synthetic.intrinsics(H, W)
    h = nwx.NeRFReplicaInferenceHandler("office_tokyo", None)
    h._img_h, h._img_w, h._n_pix = H, W, H * W
    h._fx = h._fy = fx
    h._cx, h._cy = cx, cy
    h.load_state_dicts(*synthetic.random_state_dicts(0))
    h.max_rays_per_launch = 36 * H * W
    h.engine.reserve(36 * H * W)
    poses = synthetic.sweep_poses(36, 0).pin_memory()

    def clock(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, out

    for i in range(3):
        h.render_poses(poses[i:i + 1]); h.render_poses(poses)
    single, batch = [], []
    for _ in range(args.reps):
        for i in range(36):
            single.append(clock(lambda: h.render_poses(poses[i:i + 1]))[0])
        ms, imgs = clock(lambda: h.render_poses(poses))
        batch.append(ms)
    one = h.render_poses(poses[7:8])[0]
    same = bool(np.array_equal(one, imgs[7]))          # batching does not change a pixel
    pct = lambda a, q: float(np.percentile(np.asarray(a), q))
    print(json.dumps({
        "workload": f"36-pose GUI sweep, {W}x{H}, 64+128 samples, host poses in / uint8 frames out",
        "single_view_ms": {"p50": pct(single, 50), "p99": pct(single, 99), "n": len(single)},
        "batched_36_views_ms": {"p50": pct(batch, 50), "p99": pct(batch, 99), "n": len(batch)},
        "batched_ms_per_view": pct(batch, 50) / 36, "batched_equals_single_bitwise": same,
        "rays_per_s_batched": 36 * H * W / (pct(batch, 50) * 1e-3)}))


if __name__ == "__main__":
    main()
