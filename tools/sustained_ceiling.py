import os, sys, json
sys.path.insert(0, "nerf-workspaces-explorer_b200")
import torch, nwx
from nwx import synthetic, engine as E
dev = torch.device("cuda:0")
eng = nwx.Engine(dev)
sd_c, sd_f = synthetic.random_state_dicts(0)
eng.load_weights(E.COARSE, sd_c); eng.load_weights(E.FINE, sd_f)
H, W = 480, 640
fx, fy, cx, cy = synthetic.intrinsics(H, W)
rays = eng.raygen(synthetic.sweep_poses(1, 0), H, W, fx, fy, cx, cy, 0.1, 10.0)
z = torch.sort(torch.rand(H * W, 192, device=dev) * 9.9 + 0.1, -1)[0]
dummy = torch.zeros(5 * 4096 * 2, device=dev)
res = {}
def run(name, n):
    for _ in range(3): eng.mlp_forward(E.FINE, rays, z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): eng.mlp_forward(E.FINE, rays, z)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res[name] = {"launches": n, "ms": round(ms, 2), "tflops": round(1186816 * H * W * 192 / ms / 1e9, 1)}
run("production, 5 launches", 5)
run("production, 60 launches (sustained ~3.3 s)", 60)
eng.debug_tap(-4, dummy)
run("no epilogue work, 5 launches", 5)
run("no epilogue work, 60 launches (sustained)", 60)
eng.debug_tap(-1, None)
print(json.dumps(res))
