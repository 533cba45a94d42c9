#!/usr/bin/env python
"""BASELINE.json config 4 on its own: the `train` record of bench.py (NeRFReplicaTrainingHandler.step, 4096 rays per
GPU, synthetic Replica-shaped banks, gradient all-reduce over NCCL) without the render benchmark around it -- the
command the ncu captures of the training kernels run.  Prints one JSON line.

    python tools/train_bench.py [--steps 20]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py ..."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nerf-workspaces-explorer_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20, help="bench.py's --steps: 5x as many training steps are timed (20..100)")
    args = ap.parse_args()
    import bench
    import nwx
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    rec = bench.run_train_bench(nwx, dev, world, rank, barrier, args)
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
