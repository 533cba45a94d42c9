#!/usr/bin/env python
"""Device timeline of the training step (CUPTI through torch.profiler; no nsys in this image): per kernel the mean
duration inside a pipelined run of NeRFReplicaTrainingHandler.step, the stream it ran on, and the idle time of the
main stream -- what the per-launch ncu list (isolated replays) cannot show.  Writes gpurun_out/train_timeline.json.

    python tools/train_timeline.py [--steps 8]"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nerf-workspaces-explorer_b200"))

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=8)
    args = ap.parse_args()
    import bench
    import nwx
    from nwx import synthetic
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    n_img, h, w = bench.TRAIN_BANK
    fx, fy, cx, cy = synthetic.intrinsics(h, w)
    poses = synthetic.sweep_poses(36, 0).repeat(n_img // 36, 1, 1)[:n_img]
    bank = nwx.Engine(dev).raygen(poses, h, w, fx, fy, cx, cy, bench.NEAR, bench.FAR).view(n_img, h * w, 11)
    rgbs = torch.rand((n_img, h * w, 3), device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    cfg = nwx.config.default_config()
    cfg["rendering"]["n_rays"] = bench.TRAIN_RAYS_PER_GPU
    cfg["experiment"].update(image_height=h, image_width=w)
    th = nwx.NeRFReplicaTrainingHandler("office_tokyo", cfg, bank, rgbs, *synthetic.random_state_dicts(0), device=dev, seed=2)
    for i in range(30):
        th.step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        e0.record()
        for i in range(args.steps):
            th.step(30 + i)
        e1.record()
        torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / args.steps
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ev.sort(key=lambda e: e.time_range.start)
    t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
    # which stream is the main one: the stream with the most busy time
    busy = collections.Counter()
    for e in ev:
        busy[getattr(e, "stream", None) if hasattr(e, "stream") else e.device_index] += e.time_range.end - e.time_range.start
    per = collections.OrderedDict()
    for e in ev:
        k = e.name[:70]
        d = per.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += e.time_range.end - e.time_range.start
    # union of busy intervals over all streams -> device idle time
    spans = sorted((e.time_range.start, e.time_range.end) for e in ev)
    covered, cur_s, cur_e = 0.0, spans[0][0], spans[0][1]
    for s, t in spans[1:]:
        if s > cur_e:
            covered += cur_e - cur_s
            cur_s, cur_e = s, t
        else:
            cur_e = max(cur_e, t)
    covered += cur_e - cur_s
    # gaps larger than 1 us in the union timeline, attributed to the kernel that follows
    gaps = collections.Counter()
    cur_e = spans[0][1]
    ev_by_start = {e.time_range.start: e.name[:70] for e in ev}
    for s, t in spans[1:]:
        if s > cur_e:
            gaps[ev_by_start[s]] += s - cur_e
        cur_e = max(cur_e, t)
    rec = {"what": "torch.profiler (CUPTI) timeline of %d pipelined training steps, 4096 rays, one B200" % args.steps,
           "ms_per_step_events": ms_step, "window_ms_per_step": (t1 - t0) / 1e3 / args.steps,
           "device_busy_ms_per_step": covered / 1e3 / args.steps,
           "device_idle_ms_per_step": ((t1 - t0) - covered) / 1e3 / args.steps,
           "sum_of_kernel_ms_per_step": sum(v[1] for v in per.values()) / 1e3 / args.steps,
           "kernels_us_mean": {k: [v[0] / args.steps, round(v[1] / v[0], 2)] for k, v in per.items()},
           "idle_before_us_per_step": {k: round(v / args.steps, 2) for k, v in gaps.most_common(20)}}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "train_timeline.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps({k: rec[k] for k in ("ms_per_step_events", "window_ms_per_step", "device_busy_ms_per_step",
                                          "device_idle_ms_per_step", "sum_of_kernel_ms_per_step")}))


if __name__ == "__main__":
    main()
