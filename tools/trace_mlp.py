#!/usr/bin/env python
"""Timeline of one CTA of the fused MLP kernel (debug instantiation): per layer and tile, how long the MMA issuer
waited before issuing, how long the epilogue warps waited for the accumulator and how long they worked, in SM cycles.

    python tools/trace_mlp.py [iteration] [variant] [raw]"""
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
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng.set_mlp_variant(variant)
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
if len(sys.argv) > 3:
    for r in out[:400]:
        print(f"{r[0]:9d} {r[1]:5s} ev={r[2]:2d} it={r[3]} l={r[4]} t={r[5]}")
ev_at = {(r[1], r[2], r[4], r[5]): r[0] for r in out}
start = min(r[0] for r in out)
nxt = [r[0] for r in rows if r[3] == it_show + 1 and r[1] == "mma"]
print(f"variant {variant}, iteration {it_show}: {(min(nxt) - start) if nxt else -1} cycles from its first MMA event to the next iteration's")
print(" l t | mma: start  waited  issued | epi0: wait-from  waited  worked  done")
for l in range(10):
    for t in range(2):
        m1, m2, m3 = (ev_at.get(("mma", e, l, t)) for e in (1, 2, 3))
        e11, e12, e13 = (ev_at.get(("epi0", e, l, t)) for e in (11, 12, 13))
        if m1 is None and e11 is None:
            continue
        f = lambda v: "      -" if v is None else f"{v - start:7d}"
        d = lambda a, b: "     -" if a is None or b is None else f"{b - a:6d}"
        print(f"{l:2d} {t} | {f(m1)} {d(m1, m2)} {d(m2, m3)} | {f(e11)} {d(e11, e12)} {d(e12, e13)} {f(e13)}")
