"""A short, quiet target for ncu: a few forwards of one precision / variant.

    python tools/ncu_target.py <precision> [rays] [iters]
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch  # noqa: E402

import mipnerf_pl_b200 as mp  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
rays_n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = "cuda:0"
model = mp.MipNerf(precision=precision)
model.load_state_dict(mp.make_state_dict(0))
model = model.to(dev).eval()
rays = mp.namedtuple_map(lambda t: t.to(dev), mp.random_ray_batch(rays_n, seed=0))
for _ in range(iters):
    model(rays, False, True)
torch.cuda.synchronize()
print("done", precision, rays_n, iters)
