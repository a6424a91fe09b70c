"""BASELINE configs[4]: ray-batch sweep 1k..64k rays at 128+128 samples (and configs[2], the
multi-scale ray mix, with --multiscale).  Launch alone (1 GPU) or under torchrun (rays sharded: each
rank renders `batch` rays, one all_gather of fine RGB).  Device-timed with CUDA events, L2 flushed
between steps, max over ranks; prints one JSON object with a row per batch size.

    python tools/sweep.py [--multiscale] [--precision bf16] [--steps 30]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import mipnerf_pl_b200 as mp  # noqa: E402

FLOP_PER_RAY = 312_475_648


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--multiscale", action="store_true")
    ap.add_argument("--batches", default="1024,2048,4096,8192,16384,32768,65536")
    ap.add_argument("--peak-tflops", type=float, default=1679.2)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = mp.MipNerf(precision=args.precision)
    model.load_state_dict(mp.make_state_dict(seed=0, kind="xavier"))
    model = model.to(dev).eval()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for B in [int(b) for b in args.batches.split(",")]:
        rays = mp.namedtuple_map(lambda t: t.to(dev), mp.random_ray_batch(B, seed=rank, multiscale=args.multiscale))
        gathered = torch.empty(world * B, 3, device=dev) if world > 1 else None

        def step():
            ret = model(rays, False, True)
            if world > 1:
                dist.all_gather_into_tensor(gathered, ret[-1][0])

        for _ in range(args.warmup):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in ev:
            flush.zero_()
            a.record()
            step()
            b.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / args.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        rps = world * B / (ms * 1e-3)
        rows.append({"rays_per_gpu": B, "global_rays": world * B, "ms_per_step": round(ms, 4), "rays_per_s": rps,
                     "frac_of_mlp_roofline_per_gpu": rps / world * FLOP_PER_RAY / 1e12 / args.peak_tflops})
    if rank == 0:
        print(json.dumps({"what": "ray-batch sweep, 128+128 samples, " + ("multi-scale mix" if args.multiscale
                                                                         else "single-scale 800x800"),
                          "n_gpus": world, "precision": args.precision, "steps": args.steps,
                          "peak_tflops": args.peak_tflops, "rows": rows}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
