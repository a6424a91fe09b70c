"""Cumulative cycle counters of CTA 0 of the v4 (TS) level kernel (trace build).

    MIPNERF_B200_LIB=.../libmipnerf_b200.trace.so python tools/v4_counters.py [rays]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
os.environ["MIPNERF_B200_TC_VARIANT"] = "v4"
import mipnerf_pl_b200 as mp  # noqa: E402
from mipnerf_pl_b200 import _cabi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = "cuda:0"
model = mp.MipNerf(precision="bf16", num_levels=1)
model.load_state_dict(mp.make_state_dict(0))
model = model.to(dev).eval()
rays = mp.namedtuple_map(lambda t: t.to(dev), mp.random_ray_batch(n, seed=0))
for _ in range(5):
    model(rays, False, True)
buf = torch.zeros(64 + 2 * 5 * 15000, dtype=torch.int64, device=dev)
fn = C.CDLL(_cabi.LIB_PATH).mipnerf_b200_debug_set_trace_buffer
fn.argtypes = [C.c_void_p]
assert fn(buf.data_ptr()) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
model(rays, False, True)
e1.record()
torch.cuda.synchronize()
fn(None)
c = buf[16:48].cpu().tolist()
rays_per_slot = -(-n // 296)          # rounds: 74 pairs x 4 rays
print(f"launch {e0.elapsed_time(e1):.4f} ms, {rays_per_slot} rounds (rays per slot)")
m = c[0:8]
print(f"MMA thread per round (2 slots): total {m[4] / rays_per_slot:.0f} | wait f_ready {m[0] / rays_per_slot:.0f} | "
      f"wait a_ready {m[1] / rays_per_slot:.0f} | wait acc_drained {m[2] / rays_per_slot:.0f} | wait w_full {m[3] / rays_per_slot:.0f}")
w = c[8:16]
print(f"worker (slot 0) per ray: wait acc_full (first) {w[0] / rays_per_slot:.0f} | half-0 epilogue {w[1] / rays_per_slot:.0f} | "
      f"wait acc_full (second half) {w[2] / rays_per_slot:.0f} | half-1 epilogue + store + arrive {w[3] / rays_per_slot:.0f}")
