#!/usr/bin/env python
"""Timeline of one CTA of the fused MLP kernel (debug instantiation): prints, per layer, when the MMA
issuer waited / issued and when the epilogue warps waited / finished, in SM cycles."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nerf-workspaces-explorer_b200"))
import torch
import nwx
from nwx import synthetic, engine as E
dev = torch.device("cuda:0")
eng = nwx.Engine(dev)
sd_c, sd_f = synthetic.random_state_dicts(0)
eng.load_weights(E.COARSE, sd_c); eng.load_weights(E.FINE, sd_f)
H, W = 480, 640
fx, fy, cx, cy = synthetic.intrinsics(H, W)
rays = eng.raygen(synthetic.sweep_poses(1, 0), H, W, fx, fy, cx, cy, 0.1, 10.0)
z = torch.sort(torch.rand(H * W, 192, device=dev) * 9.9 + 0.1, -1)[0]
eng.mlp_forward(E.FINE, rays, z)
buf = torch.zeros(5 * 4096 * 2, device=dev, dtype=torch.float32)
eng.debug_tap(-2, buf)
eng.mlp_forward(E.FINE, rays, z)
torch.cuda.synchronize()
eng.debug_tap(-1, None)
ev = buf.view(torch.int32).cpu().view(5, 4096, 2).numpy()
names = ["prod", "mma", "pe", "epi0", "epi1"]
import numpy as np
t0 = min(int(ev[r, 0, 1]) & 0xFFFFFFFF for r in (1, 3) if ev[r, 0, 0])
rows = []
for r in range(5):
    for tag, clk in ev[r]:
        if tag == 0: break
        rows.append(((int(clk) & 0xFFFFFFFF) - t0, names[r], int(tag) // 100000, (int(tag) // 1000) % 100, (int(tag) // 10) % 100, int(tag) % 10))
rows.sort()
it_show = int(sys.argv[1]) if len(sys.argv) > 1 else 3
out = [r for r in rows if r[3] == it_show and r[1] != "prod"]
for r in out[:400]:
    print(f"{r[0]:9d} {r[1]:5s} ev={r[2]:2d} it={r[3]} l={r[4]} t={r[5]}")
