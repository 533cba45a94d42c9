#!/usr/bin/env python
"""BASELINE.json config 4: data-parallel NeRF training step, 4096 rays per GPU, synthetic
Replica-shaped rays / pixels, gradient all-reduce over NCCL.  Prints one JSON line.

    python tools/train_bench.py [--steps 20] [--warmup 5] [--rays 4096]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py ..."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nerf-workspaces-explorer_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

FLOP_PER_POINT_TRAIN = 3489024      # fwd + bwd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--rays", type=int, default=4096)
    args = ap.parse_args()
    import nwx
    from nwx import synthetic
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    H, W = 240, 320
    fx, fy, cx, cy = synthetic.intrinsics(H, W)
    eng = nwx.Engine(dev)
    bank = eng.raygen(synthetic.sweep_poses(36, 0), H, W, fx, fy, cx, cy, 0.1, 10.0)        # [36*H*W, 11] ray bank
    gen = torch.Generator(device=dev).manual_seed(2 + rank)
    tr = nwx.Trainer(eng, *synthetic.random_state_dicts(0), seed=2 + rank)

    def batch():
        idx = torch.randint(0, bank.shape[0], (args.rays,), device=dev, generator=gen)
        return bank[idx], torch.rand((args.rays, 3), device=dev, generator=gen)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        tr.step(*batch(), i)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = nwx.engine.launch_count()
    e0.record()
    for i in range(args.steps):
        loss = tr.step(*batch(), args.warmup + i)
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms) / args.steps
    # data-parallel invariant: every rank holds bit-identical parameters after the same steps
    chk = tr.params.double().sum().reshape(1)
    same = True
    if world > 1:
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        same = all(bool(torch.equal(c, allc[0])) for c in allc)
    if rank == 0:
        pts = args.rays * 256
        print(json.dumps({
            "workload": f"data-parallel training step, {args.rays} rays/GPU, 64+128 samples, Adam, grad all-reduce",
            "n_gpus": world, "ms_per_step": ms_step, "rays_per_s": world * args.rays / (ms_step * 1e-3),
            "tflops_per_gpu": FLOP_PER_POINT_TRAIN * pts / (ms_step * 1e-3) / 1e12,
            "kernel_launches_per_step": (nwx.engine.launch_count() - launches0) / args.steps,
            "loss": [float(loss[0]), float(loss[1])], "params_identical_across_ranks": same}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
