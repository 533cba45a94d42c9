#!/usr/bin/env python
"""How much does saving the backward's operands cost the forward?  Times the render (inference) instantiation of the
fused MLP kernel on exactly the training step's points (4096 rays x 64 / 192 samples) next to the per-kernel times of
the training forward (tools/train_timeline.py), at the clocks of a short burst.  One JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nerf-workspaces-explorer_b200"))

import torch  # noqa: E402


def main():
    import nwx
    from nwx import engine as E
    from nwx import synthetic
    dev = torch.device("cuda", 0)
    eng = nwx.Engine(dev)
    sd_c, sd_f = synthetic.random_state_dicts(0)
    eng.load_weights(E.COARSE, sd_c); eng.load_weights(E.FINE, sd_f)
    H, W = 240, 320
    fx, fy, cx, cy = synthetic.intrinsics(H, W)
    bank = eng.raygen(synthetic.sweep_poses(36, 0), H, W, fx, fy, cx, cy, 0.1, 10.0)
    gen = torch.Generator(device=dev).manual_seed(1)
    rays = bank[torch.randint(0, bank.shape[0], (4096,), device=dev, generator=gen)].contiguous()
    out = {}
    for which, S in ((E.COARSE, 64), (E.FINE, 192)):
        z = torch.sort(torch.rand((4096, S), device=dev, generator=gen) * 9.9 + 0.1, -1)[0].contiguous()
        for _ in range(5):
            eng.mlp_forward(which, rays, z)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            eng.mlp_forward(which, rays, z)
        e1.record()
        torch.cuda.synchronize()
        out[f"render_kernel_ms_4096x{S}"] = e0.elapsed_time(e1) / 50     # includes the dirbias launch (0.006 ms)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
