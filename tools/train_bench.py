"""Time the fp32 training step (forward + backward + Adam) on one GPU: ms per step for a 4096-ray batch of
BASELINE configs[1] shape, with the per-kernel breakdown from the library's launch accounting.

    python tools/train_bench.py [--rays 4096] [--steps 5]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402

import mipnerf_pl_b200 as mp  # noqa: E402
from mipnerf_pl_b200 import _cabi  # noqa: E402

FLOP_PER_RAY_FWD = 312_475_648


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16", "fp16"])
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    model = mp.MipNerf(precision=args.precision)
    model.load_state_dict(mp.make_state_dict(seed=0, kind="xavier"))
    model = model.to(dev)
    opt = mp.FusedAdam(model.parameters(), lr=5e-4)
    rays = mp.namedtuple_map(lambda t: t.to(dev), mp.random_ray_batch(args.rays, seed=0, multiscale=True))
    rgbs = torch.rand(args.rays, 3, device=dev)

    def step():
        out = mp.forward_backward(model, rays, rgbs, True, True)
        opt.step()
        return out

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    lib = _cabi.lib()
    # the step time: CUDA events around `steps` un-instrumented steps ...
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    # ... and the per-kernel breakdown from a second pass with an event pair around every launch (which itself
    # costs a few microseconds per launch, so its sum is not the step time)
    _cabi.profile_snapshot(reset=True)
    lib.mipnerf_b200_profile_enable(1)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    lib.mipnerf_b200_profile_enable(0)
    prof = _cabi.profile_snapshot(reset=True)
    flops = 3 * args.rays * FLOP_PER_RAY_FWD          # forward + dgrad + wgrad (dgrad of layer 0 is not needed)
    print(json.dumps({"what": f"{args.precision} training step (forward + backward + Adam), randomized, 128+128 samples",
                      "rays": args.rays, "ms_per_step": ms, "rays_per_s": args.rays / (ms * 1e-3),
                      "approx_tflops": flops / (ms * 1e-3) / 1e12, "loss": float(out["loss"]),
                      "kernel_ms_per_step": {k: round(v[1] / args.steps, 3) for k, v in prof.items() if v[2]},
                      "launches_per_step": {k: v[0] / args.steps for k, v in prof.items() if v[0]}}))


if __name__ == "__main__":
    main()
