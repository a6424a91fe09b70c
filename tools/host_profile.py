"""Host-side cost of one MipNerf.forward call (Python + ctypes + launches, no synchronisation inside the loop):
mean microseconds per call and the cProfile top of the call tree.  The GPU must not be the bottleneck of the
loop, so the batch is tiny (256 rays)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import mipnerf_pl_b200 as mp  # noqa: E402

dev = torch.device("cuda", 0)
model = mp.MipNerf(precision=os.environ.get("TRACE_PRECISION", "bf16"))
model.load_state_dict(mp.make_state_dict(seed=0))
model = model.to(dev).eval()
staging = mp.RayStaging(mp.random_ray_batch(256, seed=0))
rays = staging.to(dev)
for _ in range(20):
    model(rays, False, True)
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for _ in range(n):
    model(rays, False, True)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"forward host time: {(t1 - t0) / n * 1e6:.1f} us per call")
t0 = time.perf_counter()
for _ in range(n):
    staging.to(dev)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"RayStaging.to host time: {(t1 - t0) / n * 1e6:.1f} us per call")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    model(rays, False, True)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
