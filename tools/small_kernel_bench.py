#!/usr/bin/env python
"""A/B timing of the HBM-side kernels of one 640x480 frame (K2 sample_pdf+merge, K4 compositing) -- the
production front ends against their fallbacks, each variant in its own process (the choice is read from the
environment once per process), CUDA events over `--iters` back-to-back launches on inputs larger than L2.

    python tools/small_kernel_bench.py [--iters 20]

Prints one JSON line: per variant the ms per launch, the algorithmic GB/s (SURVEY 8d bytes per ray) and a
checksum of the outputs, so that bit-identical results across variants can be read off."""
import argparse
import json
import os
import subprocess
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nerf-workspaces-explorer_b200"))

N = 640 * 480
BYTES_PER_RAY = {"composite_coarse": 1572, "composite_fine": 3880, "sample_pdf": 1280}


def worker(iters: int) -> dict:
    import torch
    import nwx
    from nwx import engine as E
    from nwx import synthetic
    dev = torch.device("cuda", 0)
    eng = nwx.Engine(dev)
    sd_c, sd_f = synthetic.random_state_dicts(0)
    eng.load_weights(E.COARSE, sd_c); eng.load_weights(E.FINE, sd_f)
    H, W = 480, 640
    fx, fy, cx, cy = synthetic.intrinsics(H, W)
    rays = eng.raygen(synthetic.sweep_poses(36, 0)[:1], H, W, fx, fy, cx, cy, 0.1, 10.0)
    out = eng.render_rays(rays, want=("raw_coarse", "raw_fine", "z_vals_coarse", "z_vals_fine", "weights_coarse", "rgb_fine"))
    raw_c, raw_f, z_c, z_f, w_c = (out[k] for k in ("raw_coarse", "raw_fine", "z_vals_coarse", "z_vals_fine", "weights_coarse"))
    rays_d = rays[:, 3:6].contiguous()

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        times = []                      # median of per-launch times: the wrappers allocate their outputs (hundreds of
        for _ in range(iters):          # MB), and an occasional cudaMalloc inside the loop would spoil a mean
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn()
            e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        times.sort()
        return times[len(times) // 2], r

    crc = lambda *ts: zlib.crc32(b"".join(t.detach().cpu().numpy().tobytes() for t in ts if t is not None))
    res = {}
    ms, r = timeit(lambda: E.composite(raw_c, z_c, rays_d, want_weights=True))
    # the production coarse pass puts two rays in a warp: its sums associate differently from the one-ray fallback, so
    # only the weights (a function of the fp64 scan) are compared bit for bit; the maps are compared by value
    res["composite_coarse"] = {"ms": ms, "crc": crc(r[3]), "maps_sample": [float(v) for v in torch.cat([r[0][:2].reshape(-1), r[4][:2]]).cpu()]}
    ms, r = timeit(lambda: E.composite(raw_f, z_f, rays_d, want_weights=False))
    res["composite_fine"] = {"ms": ms, "crc": crc(r[0], r[1], r[2], r[4])}
    ms, r = timeit(lambda: E.sample_pdf_merge(z_c, w_c, 128, want_inds=False))
    res["sample_pdf"] = {"ms": ms, "crc": crc(r[0], r[1])}
    ms, r = timeit(lambda: E.sample_pdf_merge(z_c, w_c, 128, want_inds=True))
    res["sample_pdf_with_inds"] = {"ms": ms, "crc": crc(r[0], r[1], r[2])}
    for k, v in res.items():
        b = BYTES_PER_RAY.get(k.replace("_with_inds", ""), 0) + (1024 if k.endswith("inds") else 0)
        v["algorithmic_gbs"] = b * N / (v["ms"] * 1e-3) / 1e9
    # what this box's HBM does for the three access mixes (torch library kernels, 4 GiB each): the training
    # forward / dX kernels are write-dominated, dW is read-only, the peak in MEASURED_PEAKS.json is a copy
    del out, raw_c, raw_f
    a = torch.empty(1 << 30, device=dev, dtype=torch.float32)
    b = torch.empty(1 << 30, device=dev, dtype=torch.float32)
    ms, _ = timeit(lambda: a.zero_())
    res["hbm_write_only_gbs"] = a.numel() * 4 / (ms * 1e-3) / 1e9
    ms, _ = timeit(lambda: torch.sum(a))
    res["hbm_read_only_gbs"] = a.numel() * 4 / (ms * 1e-3) / 1e9
    ms, _ = timeit(lambda: b.copy_(a))
    res["hbm_copy_gbs"] = 2 * a.numel() * 4 / (ms * 1e-3) / 1e9
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--worker", action="store_true")
    args = ap.parse_args()
    if args.worker:
        print("RESULT " + json.dumps(worker(args.iters)))
        return
    variants = {"production": {}, "fallbacks": {"NWX_COMPOSITE": "direct", "NWX_SAMPLE_PDF": "generic"}}
    out = {}
    for name, env in variants.items():
        proc = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", "--iters", str(args.iters)],
                              env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
        lines = [l for l in proc.stdout.splitlines() if l.startswith("RESULT ")]
        out[name] = json.loads(lines[-1][7:]) if lines else {"error": proc.stderr[-800:]}
    if all("error" not in v for v in out.values()):
        out["bit_identical"] = {k: out["production"][k]["crc"] == out["fallbacks"][k]["crc"]
                                for k, v in out["production"].items() if isinstance(v, dict)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
