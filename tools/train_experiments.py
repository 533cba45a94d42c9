#!/usr/bin/env python
"""Where does the training step's time go?  Times the step (4096 rays, as bench.py's `train` record) with the
kernels' timing experiments switched on (nwx_debug_experiment; gradients are WRONG on purpose in 11 / 12):
   0  production
  11  forward / dX epilogues do not wait for the previous TMA store of their shared-memory tile
  12  no TMA stores of the activation / gradient tile images at all
  13  the forward neither builds nor stores the ReLU' bit masks
  14  no named barriers around the tile writes (the store's wait is skipped too)
  15  the forward does not store the views hidden
  18  the training forward saves nothing at all (null save pointers)
Prints one JSON line with ms per step and the per-kernel CUDA-event times of forward, dX and the rest.
Needs a library built with the experiments compiled in:  make -C nerf-workspaces-explorer_b200 clean && make -C
nerf-workspaces-explorer_b200 EXPERIMENTS=1  (the product build has no such branches and rejects the codes)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nerf-workspaces-explorer_b200"))

import torch  # noqa: E402


def main():
    import nwx
    from nwx import synthetic
    from nwx._lib import check
    dev = torch.device("cuda", 0)
    H, W = 240, 320
    fx, fy, cx, cy = synthetic.intrinsics(H, W)
    eng = nwx.Engine(dev)
    bank = eng.raygen(synthetic.sweep_poses(36, 0), H, W, fx, fy, cx, cy, 0.1, 10.0)
    tr = nwx.Trainer(eng, *synthetic.random_state_dicts(0), seed=2)
    gen = torch.Generator(device=dev).manual_seed(1)
    rays = bank[torch.randint(0, bank.shape[0], (4096,), device=dev, generator=gen)]
    gt = torch.rand((4096, 3), device=dev, generator=gen)
    out = {}
    codes = [int(c) for c in sys.argv[1:]] or [0, 11, 12, 13, 14, 15, 18, 0]
    for code in codes:
        check(nwx.lib().nwx_debug_experiment(eng._ctx, code))
        for i in range(3):
            tr.forward_backward(rays, gt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            tr.forward_backward(rays, gt)
        e1.record()
        torch.cuda.synchronize()
        out.setdefault(str(code), []).append(e0.elapsed_time(e1) / 20)
    check(nwx.lib().nwx_debug_experiment(eng._ctx, 0))
    print(json.dumps({"what": "ms per forward+backward of 4096 rays (no optimiser step) under the timing experiments", "ms": out}))


if __name__ == "__main__":
    main()
