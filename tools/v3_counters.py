"""Cumulative cycle counters of CTA 0 of the v3 level kernel (trace build).

    MIPNERF_B200_LIB=.../libmipnerf_b200.trace.so MIPNERF_B200_TC_VARIANT=v3 python tools/v3_counters.py [rays]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
os.environ.setdefault("MIPNERF_B200_TC_VARIANT", "v3")
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
rounds = -(-n // 148)
print(f"launch {e0.elapsed_time(e1):.4f} ms, {rounds} rays per CTA")
mma = c[0:8]
print(f"MMA thread: total {mma[5]} cycles = {mma[5] / rounds:.0f} per ray | wait f_ready {mma[0] / rounds:.0f} | "
      f"wait epi (layer 0) {mma[1] / rounds:.0f} | wait epi q0-2 {mma[2] / rounds:.0f} | wait epi q3 {mma[3] / rounds:.0f} | "
      f"wait w_full {mma[4] / rounds:.0f} | rest (issue) {(mma[5] - sum(mma[0:5])) / rounds:.0f}")
for g in range(2):
    w = c[8 + 8 * g: 16 + 8 * g]
    print(f"worker group {g}: wait acc_full {w[0] / rounds:.0f} per ray | epilogues {w[1] / rounds:.0f} | composite {w[2] / rounds:.0f}"
          f" | of the epilogues: tmem ld+wait {w[3] / rounds:.0f}, convert+st issue {w[4] / rounds:.0f}, wait::st {w[5] / rounds:.0f}")
i = c[24:32]
print(f"IPE warp: wait f_free {i[0] / rounds:.0f} per ray | features {i[1] / rounds:.0f}")
