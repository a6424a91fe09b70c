"""Stress a kernel variant: many forwards with an L2 flush in between, eager and graph-replayed, checking every result
against the first one (bit-equal).  MIPNERF_B200_TC_VARIANT selects the variant."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch  # noqa: E402

import mipnerf_pl_b200 as mp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
mode = sys.argv[3] if len(sys.argv) > 3 else "both"
dev = "cuda:0"
model = mp.MipNerf(precision="bf16")
model.load_state_dict(mp.make_state_dict(0))
model = model.to(dev).eval()
host = mp.random_ray_batch(n, seed=0)
rays = mp.namedtuple_map(lambda t: t.to(dev), host)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
dbg = None
if "trace" in os.environ.get("MIPNERF_B200_LIB", ""):
    import ctypes as C
    from mipnerf_pl_b200 import _cabi
    dbg = torch.zeros(1024, dtype=torch.int64).pin_memory()          # mapped host memory: survives a device trap
    fn = C.CDLL(_cabi.LIB_PATH).mipnerf_b200_debug_set_trace_buffer
    fn.argtypes = [C.c_void_p]
    assert fn(dbg.data_ptr()) == 0


def report():
    if dbg is None:
        return
    n = int(dbg[48])
    print("timeout records:", n)
    for i in range(min(n, 64)):
        print("  site", int(dbg[49 + 3 * i]), "block", int(dbg[50 + 3 * i]), "extra", int(dbg[51 + 3 * i]))


import atexit
atexit.register(report)
ref = model(rays, False, True)
torch.cuda.synchronize()
ref_rgb = ref[1][0].clone()
if mode in ("both", "eager"):
    for i in range(iters):
        flush.zero_()
        out = model(rays, False, True)
        if i % 10 == 9 or os.environ.get("STRESS_SYNC"):
            torch.cuda.synchronize()
            assert torch.equal(out[1][0], ref_rgb), f"eager iteration {i} differs"
            print("eager iteration", i, "ok", flush=True)
    torch.cuda.synchronize()
    print("eager ok", iters)
if mode in ("both", "graph"):
    gf = mp.GraphedForward(model, mp.RayStaging(host), True, dev)
    for i in range(iters):
        flush.zero_()
        out = gf.replay()
        if i % 10 == 9:
            torch.cuda.synchronize()
            assert torch.equal(out[1][0], ref_rgb), f"graph iteration {i} differs"
    torch.cuda.synchronize()
    print("graph ok", iters)
