#!/usr/bin/env python
"""Times the fused MLP kernel variants on the full-frame fine launch shape (58.98 M points)."""
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
res = {}
for v in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]:
    eng.set_mlp_variant(v)
    for _ in range(2):
        eng.mlp_forward(E.FINE, rays, z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.mlp_forward(E.FINE, rays, z)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res.setdefault(v, []).append({"ms": ms, "tflops": 1186816 * H * W * 192 / ms / 1e9})
if os.environ.get("NWX_VARIANTS_ONLY", "1") == "1":
    print(json.dumps(res)); sys.exit(0)
# timing experiments in the debug instantiation (results are wrong on purpose)
eng.set_mlp_variant(1)
dummy = torch.zeros(5 * 4096 * 2, device=dev)
for mode, name in ((-5, "debug instantiation, normal"), (-3, "no STS in hidden epilogues"), (-4, "no epilogue work"),
                   (-6, "MMA issuer never waits for weights (results wrong on purpose)"),
                   (-7, "hidden epilogues process half their columns (results wrong on purpose)"),
                   (-8, "epilogues only read TMEM: no ALU work, no smem stores (results wrong on purpose)")):
    eng.debug_tap(mode, dummy)
    for _ in range(2):
        eng.mlp_forward(E.FINE, rays, z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.mlp_forward(E.FINE, rays, z)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res[name] = {"ms": ms, "tflops": 1186816 * H * W * 192 / ms / 1e9}
eng.debug_tap(-1, None)
print(json.dumps(res))
