"""Timeline of CTA 0 of the fused level kernel (trace build): who waits for whom.

    MIPNERF_B200_LIB=.../libmipnerf_b200.trace.so python tools/tc_trace.py [pair|single|shared] [levels]

levels = 1 traces the coarse launch, 2 the fine launch (the buffer keeps the last launch).
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
variant = sys.argv[1] if len(sys.argv) > 1 else "pair"
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 1
os.environ["MIPNERF_B200_TC_VARIANT"] = variant
import mipnerf_pl_b200 as mp  # noqa: E402
from mipnerf_pl_b200 import _cabi  # noqa: E402

lib = _cabi.lib()
dev = "cuda:0"
model = mp.MipNerf(precision=os.environ.get("TRACE_PRECISION", "bf16"))
model.load_state_dict(mp.make_state_dict(0))
model = model.to(dev).eval()
rays = mp.namedtuple_map(lambda t: t.to(dev), mp.random_ray_batch(4096, seed=0))
for _ in range(3):
    model(rays, False, True)
torch.cuda.synchronize()
REG = 15000
NREG = 5                      # producer, MMA issuer, workers slot 0 / 1, IPE warp slot 0
buf = torch.zeros(8 + 2 * NREG * REG, dtype=torch.int64, device=dev)
fn = C.CDLL(_cabi.LIB_PATH).mipnerf_b200_debug_set_trace_buffer
fn.argtypes = [C.c_void_p]
assert fn(buf.data_ptr()) == 0
model.num_levels = levels     # the trace buffer keeps the LAST launch: 1 = coarse level, 2 = fine level
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(20):           # steady clocks
    model(rays, False, True)
buf.zero_()
torch.cuda.synchronize()
e0.record()
model(rays, False, True)
e1.record()
torch.cuda.synchronize()
launch_ms = e0.elapsed_time(e1)
fn(None)
raw = buf.cpu().numpy()
parts = []
for r in range(NREG):
    cnt = int(raw[r])
    parts.append(raw[8 + 2 * r * REG: 8 + 2 * r * REG + 2 * cnt].reshape(-1, 2))
ev = np.concatenate(parts)
n = len(ev)
t0 = ev[:, 0].min()
ev = ev[np.argsort(ev[:, 0])]
print(f"{n} events, span {(ev[-1, 0] - t0)} cycles; forward call {launch_ms:.4f} ms -> >= {(ev[-1, 0] - t0) / launch_ms / 1e3:.0f} MHz effective SM clock")
role = ev[:, 1] >> 24
kind = (ev[:, 1] >> 16) & 0xff
g = (ev[:, 1] >> 8) & 0xff
x = ev[:, 1] & 0xff
t = ev[:, 0] - t0

def sel(r, k, gg=None, xx=None):
    m = (role == r) & (kind == k)
    if gg is not None:
        m &= g == gg
    if xx is not None:
        m &= x == xx
    return t[m]

# steady-state statistics over the whole launch
for slot in (0, 1):
    acc = {gg: sel(2, 2, gg, slot) for gg in range(11)}          # worker: accumulator of group gg arrived
    epi = {gg: sel(2, 3, gg, slot) for gg in range(11)}          # worker: epilogue compute finished
    arr = {gg: sel(2, 4, gg, slot) for gg in range(11)}          # worker: arrived on a_ready
    mma_go = {gg: sel(1, 0, gg, slot) for gg in range(11)}       # MMA: a_ready observed
    mma_done = {gg: sel(1, 2, gg, slot) for gg in range(11)}     # MMA: all MMAs of the group issued + commit
    print(f"--- slot {slot}: mean cycles per step over {len(acc[1])} tiles")
    for gg in range(11):
        if len(acc[gg]) and len(epi[gg]) == len(acc[gg]):
            e = (epi[gg] - acc[gg]).mean()
            f = (arr[gg] - epi[gg]).mean() if len(arr[gg]) == len(epi[gg]) else float('nan')
        else:
            e = f = float('nan')
        # time from worker arrive (group gg done) to MMA observing it for group gg+1
        nxt = gg + 1
        if nxt < 11 and len(mma_go[nxt]) == len(arr[gg]) and len(arr[gg]):
            h1 = (mma_go[nxt] - arr[gg]).mean()
        else:
            h1 = float('nan')
        iss = (mma_done[gg] - mma_go[gg]).mean() if len(mma_done[gg]) == len(mma_go[gg]) and len(mma_go[gg]) else float('nan')
        # MMA issue end -> worker sees accumulator
        h2 = (acc[gg] - mma_done[gg]).mean() if len(acc[gg]) == len(mma_done[gg]) and len(acc[gg]) else float('nan')
        print(f"g{gg:2d}: mma issue span {iss:8.0f} | commit->worker wake {h2:8.0f} | epilogue {e:8.0f} | fence+arrive {f:6.0f} | arrive->mma sees (next g) {h1:8.0f}")
a, b, r0, r1 = sel(3, 0), sel(3, 1), sel(3, 2), sel(3, 3)
if len(b) > 3 and len(r0) == len(b) == len(r1) == len(a):
    print(f"IPE warp (slot 0), mean cycles per ray: prologue before f_free {(r1 - r0)[1:].mean():.0f} | "
          f"wait f_free {(a - r1)[1:].mean():.0f} | f_free -> features published {(b - a)[1:].mean():.0f} | "
          f"idle until next round starts {(r0[1:] - b[:-1]).mean():.0f}")
else:
    print("IPE warp: events", len(r0), len(r1), len(a), len(b))
tile_start = sel(2, 0, None, 1)
if len(tile_start) > 2:
    print("tile period slot 0:", np.diff(tile_start).mean())
# per-layer cadence of slot 1: accumulator-ready times relative to the previous layer's
accs = [sel(2, 2, l, 1) for l in range(10)]
if all(len(a) == len(accs[0]) for a in accs) and len(accs[0]) > 3:
    A = np.stack(accs, 1)[1:-1]          # drop first/last tile
    d = np.diff(A, axis=1)
    print("slot 1: acc(l) - acc(l-1), l=1..9:", np.round(d.mean(0)).astype(int).tolist())
    nxt = np.stack(accs, 1)[2:, 0] - A[:, 9][:len(np.stack(accs, 1)[2:, 0])]
    print("slot 1: acc(0, next ray) - acc(9):", int(nxt.mean()), " tile period:", int(np.diff(np.stack(accs,1)[:,0]).mean()))
    comp = sel(2, 5, None, 1); v_epi = sel(2, 3, 9, 1)
    if len(comp) == len(v_epi):
        print("view epilogue end -> composite done:", int((comp - v_epi).mean()))
    i0, i1 = sel(3, 0), sel(3, 1)
    f5 = sel(1, 2, 5, 0)
    print("IPE start after layer-5 commit (slot 0):", int((i0[1:len(f5)+1] - f5[:len(i0)-1]).mean()) if len(i0) > 1 else None)
    g0_ready = sel(1, 0, 0, 0)
    if len(g0_ready) == len(i1):
        print("MMA starts layer 0 (slot 0) after IPE done by:", np.round((g0_ready[1:] - i1[:-1])).astype(int).tolist()[:8])
# first 120 events of round 1 as a raw timeline
names = {(3, 0): "I start", (3, 1): "I done", (0, 0): "P issue", (1, 0): "M a_ready", (1, 1): "M w_full", (1, 2): "M commit", (2, 0): "W tile", (2, 1): "W feat",
         (2, 2): "W acc", (2, 3): "W epi", (2, 4): "W arrive", (2, 5): "W comp"}
start = np.searchsorted(t, tile_start[2]) if len(tile_start) > 2 else 0
for i in range(start, min(start + 150, len(t))):
    print(f"{t[i] - t[start]:8d} {names.get((role[i], kind[i]), '?'):10s} g={g[i]:2d} x={x[i]}")
