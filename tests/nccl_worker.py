"""Two (or more) NCCL ranks on real GPUs, launched by tests/test_gpu_multirank.py through torch.distributed.run:
the sharded render and the data-parallel training step must be BIT-IDENTICAL to what one rank computes alone.
Prints `RESULT {...}` on rank 0.  Uses the oracle only for synthetic weights / poses (checker side)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nerf-workspaces-explorer_b200"))

from oracle import nerf_oracle as orc  # noqa: E402


def main():
    import nwx
    from nwx import engine as E
    from nwx.dist import render_poses_sharded
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = {"world": world}
    gen = torch.Generator().manual_seed(0)
    sd_c, sd_f = orc.init_state_dict(0, generator=gen), orc.init_state_dict(0, generator=gen)

    # ---- render: one frame cut into `world` ray ranges + all-gather == the frame one rank renders alone ----
    for (H, W, B) in ((48, 64, 1), (25, 33, 1), (24, 32, 3)):       # even split, ragged split, multi-view batch
        h = nwx.NeRFReplicaInferenceHandler("office_tokyo", None, device=dev)
        h._img_h, h._img_w, h._n_pix = H, W, H * W
        fx, fy, cx, cy = orc.intrinsics(H, W)
        h._fx = h._fy = fx
        h._cx, h._cy = cx, cy
        h.load_state_dicts(sd_c, sd_f)
        poses = orc.synthetic_poses(36, 0)[4:4 + B]
        alone = h.render_poses(poses)                                # every rank renders the whole thing by itself
        gathered = render_poses_sharded(h, poses, to_host=True)      # ... and its 1/world share + all-gather
        same = bool((alone == gathered).all()) and gathered.shape == (B, H, W, 3)
        flag = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        res[f"render_{H}x{W}x{B}_bit_identical"] = bool(flag.item())

    # ---- training: 2 data-parallel steps == one rank that computes every rank's gradient and sums them ----
    fx, fy, cx, cy = orc.intrinsics(24, 32)
    bank = nwx.create_rays(4, orc.synthetic_poses(4, 1), 24, 32, fx, fy, cx, cy, 0.1, 10.0, device=dev).contiguous()
    rgbs = torch.rand(4, 24 * 32, 3, generator=torch.Generator().manual_seed(8)).to(dev).contiguous()
    n_rays, seed = 256, 5
    for overlap in (True, False):
        tr = nwx.Trainer(nwx.Engine(dev), sd_c, sd_f, seed=seed, overlap_allreduce=overlap)
        ref = nwx.Trainer(nwx.Engine(dev), sd_c, sd_f, seed=seed, data_parallel=False)
        assert tr.seed == seed + rank and ref.seed == seed
        batches_differ = True
        for step in range(2):
            rays, gt, idx = tr.engine.sample_training_batch(bank, rgbs, n_rays, tr.seed, tr.draws, want_indices=True)
            all_idx = [torch.empty_like(idx) for _ in range(world)]
            dist.all_gather(all_idx, idx)
            batches_differ &= all(not torch.equal(all_idx[0], a) for a in all_idx[1:])
            tr.step(rays, gt, step)
            # the single-rank equivalent: every rank's batch and draws, gradients summed, mean applied
            total = torch.zeros_like(ref.grads)
            for r in range(world):
                ref.seed, ref.draws = seed + r, step
                rr, gg = ref.engine.sample_training_batch(bank, rgbs, n_rays, ref.seed, ref.draws)
                ref.forward_backward(rr, gg)
                total += ref.grads
            ref.grads.copy_(total)
            ref.apply_optimizer(step, 1.0 / world)
        torch.cuda.synchronize()
        mine = tr.params.clone()
        everyone = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(everyone, mine)
        tag = "overlap" if overlap else "blocking"
        res[f"train_{tag}_params_identical_across_ranks"] = all(torch.equal(everyone[0], e) for e in everyone)
        # world == 2: a + b is order-independent, so the NCCL sum equals the local sum bit for bit; for larger worlds
        # the reduction tree's order is NCCL's, compare to rounding instead
        diff = float((mine - ref.params).abs().max())
        res[f"train_{tag}_vs_single_rank_max_abs_diff"] = diff
        res[f"train_{tag}_equals_single_rank"] = torch.equal(mine, ref.params) if world == 2 else diff <= 1e-6
        res[f"train_{tag}_batches_differ_across_ranks"] = bool(batches_differ)
    if rank == 0:
        print("RESULT " + json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
