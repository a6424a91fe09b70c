"""Data-parallel training step under torchrun (one rank per GPU, NCCL): every rank runs the fused step on its own
4096-ray shard with the GLOBAL mask_sum / ray count (DDP semantics, train.py:60 of the reference), then ONE all-reduce
of the 612 740 gradients as a flat buffer and Adam.  Prints one JSON line from rank 0: ms per step and the all-reduce
alone, both timed with CUDA events and taken as the max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench_multi.py
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import mipnerf_pl_b200 as mp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096, help="rays per rank")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16", "fp16"])
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = mp.MipNerf(precision=args.precision)
    model.load_state_dict(mp.make_state_dict(seed=0, kind="xavier"))
    model = model.to(dev)
    opt = mp.FusedAdam(model.parameters(), lr=5e-4)
    rays = mp.namedtuple_map(lambda t: t.to(dev), mp.random_ray_batch(args.rays, seed=rank, multiscale=True))
    rgbs = torch.rand(args.rays, 3, device=dev)
    mask_sum = torch.tensor(float(rays.lossmult.sum()) * world, device=dev)  # same per rank here (synthetic shards)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ar_ms = 0.0

    def step(timed):
        nonlocal ar_ms
        out = mp.forward_backward(model, rays, rgbs, True, True, mask_sum=mask_sum, global_rays=args.rays * world)
        if timed:
            ev[2].record()
        mp.allreduce_grads(list(model.parameters()), average=False)
        if timed:
            ev[3].record()
        opt.step()
        return out

    for _ in range(3):
        step(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev[0].record()
    for _ in range(args.steps):
        out = step(False)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    for _ in range(args.steps):
        step(True)
        torch.cuda.synchronize()
        ar_ms += ev[2].elapsed_time(ev[3]) / args.steps
    t = torch.tensor([ms, ar_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"what": f"{args.precision} data-parallel training step, {args.rays} rays per rank, randomized",
                          "n_gpus": world, "ms_per_step": float(t[0]), "allreduce_ms": float(t[1]),
                          "rays_per_s": args.rays * world / (float(t[0]) * 1e-3), "loss": float(out["loss"])}), flush=True)
    if world > 1:
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
