"""Achieved HBM bandwidth of the stand-alone ray kernels (the non-MLP stages of SURVEY.md §8a), each
timed inside the library with CUDA events (mipnerf_b200_profile_enable) at a size far larger than L2.
Algorithmic bytes = inputs read once + outputs written once.

    python tools/ray_kernel_bw.py [--rays 262144] [--peak-gbs 6571.9]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402

import mipnerf_pl_b200 as mp  # noqa: E402
from mipnerf_pl_b200 import _cabi, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=262144)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--peak-gbs", type=float, default=6571.9)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    B, N = args.rays, 128
    rays = mp.namedtuple_map(lambda t: t.to(dev), mp.random_ray_batch(B, seed=0))
    g = torch.Generator(device=dev).manual_seed(0)
    w = torch.rand(B, N, device=dev, generator=g) ** 4
    rgb = torch.rand(B, N, 3, device=dev, generator=g)
    dens = torch.rand(B, N, 1, device=dev, generator=g) * 4
    lib = _cabi.lib()
    t, (means, covs) = ops.sample_along_rays(rays.origins, rays.directions, rays.radii, N, rays.near, rays.far,
                                             False, False, "cone")
    pose = mp.spheric_pose(0.3)
    side = int(B ** 0.5)
    cases = {
        # name: (callable, profile kernel names, algorithmic bytes per call)
        "sample_along_rays (fenceposts + cast_rays)": (
            lambda: ops.sample_along_rays(rays.origins, rays.directions, rays.radii, N, rays.near, rays.far, False,
                                          False, "cone"),
            B * (8 + 4 * (N + 1)) + B * (28 + 4 * (N + 1) + 2 * N * 12)),
        "resample_along_rays (blur-pool + inverse CDF + cast_rays)": (
            lambda: ops.resample_along_rays(rays.origins, rays.directions, rays.radii, t, w, False, "cone", True, 0.01),
            B * (4 * (N + 1) + 4 * N + 4 * (N + 1)) + B * (28 + 4 * (N + 1) + 2 * N * 12)),
        "sorted_piecewise_constant_pdf": (
            lambda: ops.sorted_piecewise_constant_pdf(t, w + 0.01, N + 1, False),
            B * (4 * (N + 1) + 4 * N + 4 * (N + 1))),
        "integrated_pos_enc": (
            lambda: ops.integrated_pos_enc((means, covs), 0, 16),
            B * N * (24 + 384)),
        "volumetric_rendering": (
            lambda: ops.volumetric_rendering(rgb, dens, t, rays.directions, True),
            B * (N * 16 + 4 * (N + 1) + 12 + 20 + 4 * N)),
        "distloss": (
            lambda: ops.distloss(w, t),
            B * (4 * N + 4 * (N + 1) + 4)),
        "generate_rays": (
            lambda: mp.generate_rays(pose, side, side, device=dev),
            side * side * 48),
    }
    # device-resident training rays: 20 images of 800x800 random pixels + poses, one random batch of B pixel ids
    import numpy as np
    from mipnerf_pl_b200.datasets import DeviceRayBank, Scene
    n_img, hw = 20, 800
    k_inv = np.array([[1 / 1111.111, 0, -0.5 * hw / 1111.111], [0, -1 / 1111.111, 0.5 * hw / 1111.111], [0, 0, -1]],
                     dtype=np.float32)
    scene = Scene([np.zeros((hw, hw, 3), dtype=np.float32) for _ in range(n_img)], np.broadcast_to(k_inv, (n_img, 3, 3)),
                  np.stack([mp.spheric_pose(0.3 * i) for i in range(n_img)]), 1.0, 2.0, 6.0)
    bank = DeviceRayBank(scene, dev)
    ids = torch.randint(0, bank.num_pixels, (B,), device=dev, generator=g)
    cases["rays_from_pixels (DeviceRayBank: training rays + target rgb from pixel ids)"] = (
        lambda: bank.rays(ids), B * (8 + 12 + 52 + 12))
    rows = []
    for name, (fn, nbytes) in cases.items():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        _cabi.profile_snapshot(reset=True)
        lib.mipnerf_b200_profile_enable(1)
        for _ in range(args.reps):
            fn()
        torch.cuda.synchronize()
        lib.mipnerf_b200_profile_enable(0)
        prof = {k: v for k, v in _cabi.profile_snapshot(reset=True).items() if v[2]}
        ms = sum(v[1] for v in prof.values()) / args.reps
        rows.append({"stage": name, "kernels": {k: round(v[1] / v[2], 4) for k, v in prof.items()},
                     "kernel_ms_per_call": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 2),
                     "achieved_GBps": round(nbytes / (ms * 1e-3) / 1e9, 1),
                     "frac_of_hbm_peak": round(nbytes / (ms * 1e-3) / 1e9 / args.peak_gbs, 3)})
    print(json.dumps({"rays": B, "samples": N, "peak_GBps": args.peak_gbs, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
