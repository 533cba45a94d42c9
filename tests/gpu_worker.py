"""Runs one fused-MLP case in its own process and prints a JSON line.

A tcgen05 kernel with a broken barrier protocol does not fail, it spins; the kernel bounds every
wait and traps, which poisons the CUDA context.  Tests therefore run the MLP variants through
this worker (subprocess + timeout) before touching them in-process.

    python tests/gpu_worker.py mlp <variant> <n_points> [tap_layer]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nerf-workspaces-explorer_b200"))

from oracle import nerf_oracle as orc  # noqa: E402  (checker only)


def mlp_case(variant: int, n_points: int, tap_layer: int = -1) -> dict:
    import nwx
    from nwx import engine as E
    dev = torch.device("cuda:0")
    eng = nwx.Engine(dev)
    diag = torch.zeros(4, dtype=torch.int32).pin_memory()
    eng.debug_diag(diag)
    eng.set_mlp_variant(variant)
    sd = orc.init_state_dict(0)
    eng.load_weights(E.COARSE, sd)
    g = torch.Generator().manual_seed(5)
    pts = (torch.rand(n_points, 3, generator=g) - 0.5) * 12.0
    dirs = torch.nn.functional.normalize(torch.randn(n_points, 3, generator=g), dim=-1)
    pe3, pe2 = orc.positional_encoding(pts, 10, 10), orc.positional_encoding(dirs, 4, 1)
    ref = orc.mlp_forward(sd, torch.cat([pe3, pe2], -1))
    emu = orc.mlp_forward_bf16_emul(sd, pe3, pe2, fold_feature=variant <= 1)    # variants 2-4 keep the feature layer
    out = {"variant": variant, "n": n_points}
    tap = None
    if tap_layer >= 0:
        tap = torch.full((n_points, 256), float("nan"), device=dev)
        eng.debug_tap(tap_layer, tap)
    try:
        raw = eng.mlp_forward_points(E.COARSE, pts.to(dev), dirs.to(dev), 1)
        torch.cuda.synchronize()
    except Exception as exc:  # noqa: BLE001
        out.update(ok=False, error=str(exc)[:300], diag=[hex(int(v) & 0xFFFFFFFF) for v in diag])
        return out
    raw = raw.cpu()
    out.update(ok=True, finite=bool(torch.isfinite(raw).all()),
               max_vs_emul=float((raw - emu).abs().max()), mean_vs_emul=float((raw - emu).abs().mean()),
               max_vs_fp32=float((raw - ref).abs().max()), diag=[hex(int(v) & 0xFFFFFFFF) for v in diag],
               sample=[round(float(v), 5) for v in raw[0]], sample_ref=[round(float(v), 5) for v in ref[0]])
    if tap is not None:
        # layer-0 tap against the emulation: localises descriptor / layout errors
        W0, b0 = sd["_pts_linears.0.weight"], sd["_pts_linears.0.bias"]
        bf = lambda t: t.to(torch.bfloat16).float()
        h = torch.relu(torch.nn.functional.linear(bf(pe3), bf(W0)) + b0)
        for i in range(1, tap_layer + 1):
            if i > 7:
                break
            x = bf(h) if i != 5 else torch.cat([bf(pe3), bf(h)], -1)
            h = torch.relu(torch.nn.functional.linear(x, bf(sd[f"_pts_linears.{i}.weight"])) + sd[f"_pts_linears.{i}.bias"])
        t = tap.cpu()
        if tap_layer <= 7:
            out.update(tap_max_err=float((t - h).abs().max()), tap_nan=int(torch.isnan(t).sum()),
                       tap_row0=[round(float(v), 4) for v in t[0, :6]], tap_ref0=[round(float(v), 4) for v in h[0, :6]])
    return out


def train_case(n_rays: int) -> dict:
    """Forward+backward of one batch against the oracle's autograd (CPU)."""
    import nwx
    dev = torch.device("cuda:0")
    eng = nwx.Engine(dev)
    diag = torch.zeros(4, dtype=torch.int32).pin_memory()
    eng.debug_diag(diag)
    gen = torch.Generator().manual_seed(0)
    sd_c, sd_f = orc.init_state_dict(0, generator=gen), orc.init_state_dict(0, generator=gen)
    tr = nwx.Trainer(eng, sd_c, sd_f)
    g = torch.Generator().manual_seed(21)
    fx, fy, cx, cy = orc.intrinsics(24, 32)
    rays = orc.create_rays(1, orc.synthetic_poses(1, 3), 24, 32, fx, fy, cx, cy, 0.1, 10.0)[0]
    rays = rays[torch.randperm(768, generator=g)[:n_rays]].contiguous()
    gt = torch.rand(n_rays, 3, generator=g)
    t_rand, u = torch.rand(n_rays, 64, generator=g), torch.rand(n_rays, 128, generator=g)
    nc, nf = torch.randn(n_rays, 64, generator=g), torch.randn(n_rays, 192, generator=g)
    out = {"n_rays": n_rays}
    try:
        loss = tr.forward_backward(rays.to(dev), gt.to(dev), t_rand.to(dev), u.to(dev), nc.to(dev), nf.to(dev))
        torch.cuda.synchronize()
    except Exception as exc:  # noqa: BLE001
        out.update(ok=False, error=str(exc)[:300], diag=[hex(int(v) & 0xFFFFFFFF) for v in diag])
        return out
    lc, lf, gc, gf, _ = orc.training_loss_and_grads(rays, gt, sd_c, sd_f, orc.RenderConfig(), t_rand, u, nc, nf)
    out.update(ok=True, loss=[float(loss[0]), float(loss[1])], loss_ref=[float(lc), float(lf)], tensors={})
    for tag, which, ref in (("c", 0, gc), ("f", 1, gf)):
        mine = {k: v.cpu() for k, v in tr.grad_dict(which).items()}
        for k in orc.STATE_KEYS:
            a, b = mine[k].double().reshape(-1), ref[k].double().reshape(-1)
            cos = float((a @ b) / (a.norm() * b.norm() + 1e-300))
            out["tensors"][f"{tag}.{k}"] = [round(float(a.norm() / (b.norm() + 1e-300)), 4), round(cos, 5),
                                            bool(torch.isfinite(a).all())]
    return out


if __name__ == "__main__":
    kind = sys.argv[1]
    if kind == "mlp":
        res = mlp_case(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else -1)
    elif kind == "train":
        res = train_case(int(sys.argv[2]))
    else:
        raise SystemExit(f"unknown case {kind}")
    print("RESULT " + json.dumps(res))
