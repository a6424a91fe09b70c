"""The fused tcgen05 path (bf16 / fp16 operands, fp32 accumulate) on real hardware.

Two oracles:
  * the fp32 oracle (== the reference): measures what 16-bit operands cost; asserted against the
    bounds SURVEY.md §7.3 predicts (bf16 ~4e-4 relative on RGB with xavier weights, fp16 ~8x better);
  * the same oracle with the kernel's operand rounding emulated on the CPU: isolates kernel bugs
    (layout, barriers, heads, compositing) from rounding — must agree ~10x tighter.
"""
import numpy as np
import pytest
import torch

from helpers import (FLOORS, FINE_WEIGHTS_FLOOR, assert_close, exceed_stats, golden, golden_levels, golden_rays,
                     make_state_dict, oracle, oracle_rays, rel_err)

pytestmark = pytest.mark.gpu

import mipnerf_pl_b200 as mp  # noqa: E402

DEV = "cuda:0"
DT = {"bf16": torch.bfloat16, "fp16": torch.float16}


@pytest.fixture(autouse=True, params=["v4", "v3", "shared", "pair", "single"])
def tc_variant(request, monkeypatch):
    """All kernel variants: CTA pair + shared weight stream (v2), CTA pair (v1), 1-CTA kernel."""
    monkeypatch.setenv("MIPNERF_B200_TC_VARIANT", request.param)
    return request.param


def run(precision, kind, rays, seed=4, white=True):
    model = mp.MipNerf(precision=precision)
    model.load_state_dict(make_state_dict(seed=seed, kind=kind))
    model = model.to(DEV).eval()
    out = model(mp.namedtuple_map(lambda t: t.to(DEV), rays), False, white)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_tc_forward_vs_emulated_oracle(precision):
    rays = mp.random_ray_batch(300, seed=21, multiscale=True)       # 300: ragged vs 2 rays/CTA
    params = make_state_dict(seed=4, kind="xavier")
    want = oracle.forward(params, oracle_rays(rays), False, True, operand_dtype=DT[precision])
    got = run(precision, "xavier", rays)
    # same floors as the fp32 parity tests (helpers.FLOORS); bounds = 2x the errors measured on B200 (round 2):
    # the kernel and the emulation round the same operands but sum in different orders
    tol = 2e-3 if precision == "bf16" else 4e-4
    for lvl in range(2):
        for k, name in enumerate(("comp_rgb", "distance", "acc", "weights", "t_samples")):
            floor = FINE_WEIGHTS_FLOOR if (name == "weights" and lvl > 0) else FLOORS[name]
            e = rel_err(got[lvl][k].cpu().numpy(), want[lvl][k].numpy(), floor)
            print(f"{precision} level {lvl} {name}: rel err vs emulated oracle {e:.3e}")
            assert e <= tol, (precision, lvl, name, e)


@pytest.mark.parametrize("precision,bound", [("bf16", 1e-3), ("fp16", 1e-4)])
def test_tc_forward_vs_fp32_reference(precision, bound):
    rays = mp.random_ray_batch(512, seed=22)
    params = make_state_dict(seed=0, kind="xavier")
    want = oracle.forward(params, oracle_rays(rays), False, True)
    got = run(precision, "xavier", rays, seed=0)
    for lvl in range(2):
        err = float((got[lvl][0].cpu() - want[lvl][0]).abs().max())
        rel = float(((got[lvl][0].cpu() - want[lvl][0]).abs() / want[lvl][0].abs().clamp_min(1e-2)).max())
        print(f"{precision} level {lvl}: comp_rgb max abs err {err:.3e}, max rel err {rel:.3e} vs fp32 reference")
        assert rel <= bound, (precision, lvl, rel)


def test_tc_trained_like_and_black_background():
    rays = mp.random_ray_batch(256, seed=23)
    params = make_state_dict(seed=5, kind="trained_like")
    want = oracle.forward(params, oracle_rays(rays), False, False, operand_dtype=torch.bfloat16)
    got = run("bf16", "trained_like", rays, seed=5, white=False)
    for lvl in range(2):
        # x40 density head: the coarse level is well conditioned; at the fine level a bf16-sized change of one coarse
        # weight moves fine fenceposts across density edges, so kernel and emulation (same roundings, other summation
        # order) agree on almost every ray and differ by up to a few 1e-2 on the few that graze an edge
        frac, mx = exceed_stats(got[lvl][0], want[lvl][0], 0.2, rtol=4e-3)
        print(f"trained_like level {lvl}: comp_rgb vs emulated oracle: max rel err {mx:.3e}, {frac:.3%} of elements > 4e-3")
        assert mx <= (4e-3 if lvl == 0 else 5e-2) and frac <= 0.02
        assert torch.all(got[lvl][3] >= 0) and torch.all(got[lvl][2] <= 1 + 1e-4)


@pytest.mark.parametrize("precision,bound", [("bf16", (3e-2, 0.13)), ("fp16", (6e-3, 2e-2))])
@pytest.mark.parametrize("name,kind", [("forward_xavier.npz", "xavier"), ("forward_trained_like.npz", "trained_like")])
def test_tc_16bit_modes_vs_reference_goldens(precision, bound, name, kind):
    """The plain 16-bit modes against the COMMITTED outputs of the reference (the same files the fp32 and fp16x3 paths
    are held to at 1e-4): measured error printed, asserted at ~2x the round-2 measurement on the stress golden (fine
    level, relative with the 0.02 floor: bf16 6.1e-2 = 1.2e-3 absolute, fp16 9.0e-3 = 1.8e-4 absolute).  These modes
    do not claim the 1e-4 contract; fp16x3 does (test_gpu_x3.py)."""
    g = golden(name)
    seed, randomized, white = (int(v) for v in g["meta"])
    model = mp.MipNerf(precision=precision)
    model.load_state_dict(make_state_dict(seed=seed, kind=kind))
    model = model.to(DEV).eval()
    ret = model(golden_rays(g, device=DEV), bool(randomized), bool(white))
    want = golden_levels(g)
    for lvl in range(2):
        e = rel_err(ret[lvl][0].cpu().numpy(), want[lvl][0], FLOORS["comp_rgb"])
        a = float(np.abs(ret[lvl][0].cpu().numpy() - want[lvl][0]).max())
        print(f"{precision} {name} level {lvl}: comp_rgb max rel err {e:.3e} (max abs {a:.3e}) vs the reference golden")
        assert e <= bound[lvl], (precision, name, lvl, e)


def test_tc_batch_split_invariance_and_sizes():
    model = mp.MipNerf(precision="bf16")
    model.load_state_dict(make_state_dict(seed=1))
    model = model.to(DEV).eval()
    rays = mp.namedtuple_map(lambda t: t.to(DEV), mp.random_ray_batch(1000, seed=3))
    full = model(rays, False, True)
    a = model(mp.Rays(*[f[:333] for f in rays]), False, True)
    b = model(mp.Rays(*[f[333:] for f in rays]), False, True)
    for lvl in range(2):
        for k in range(5):
            assert torch.equal(torch.cat([a[lvl][k], b[lvl][k]]), full[lvl][k]), (lvl, k)
    one = model(mp.Rays(*[f[:1] for f in rays]), False, True)
    assert torch.equal(one[1][0], full[1][0][:1])
    assert model(mp.Rays(*[f[:0] for f in rays]), False, True)[1][0].shape == (0, 3)


def test_tc_full_batch_properties():
    model = mp.MipNerf(precision="bf16")
    model.load_state_dict(make_state_dict(seed=9, kind="trained_like"))
    model = model.to(DEV).eval()
    rays = mp.namedtuple_map(lambda t: t.to(DEV), mp.random_ray_batch(4096, seed=0))
    white, black = model(rays, False, True), model(rays, False, False)
    for lvl in range(2):
        rgb_w, dist, acc, w, t = white[lvl]
        assert torch.isfinite(rgb_w).all() and torch.all(w >= 0) and torch.all(acc <= 1 + 1e-4)
        assert torch.all(t[:, 1:] >= t[:, :-1])
        assert torch.allclose(rgb_w, black[lvl][0] + (1 - acc)[:, None], atol=1e-5)
        assert torch.allclose(w.sum(-1), acc, atol=1e-4)


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_tc_mlp_stage_entry(precision):
    """MLP.forward(x, view_enc, precision=...) alone (models/mip_nerf.py:75-111): the fused kernel in
    MLP-only mode (features from the caller, raw heads out) against the oracle with the kernel's
    operand rounding emulated, and against the fp32 oracle at the 16-bit bound."""
    g = torch.Generator().manual_seed(31)
    b = 37                                                            # odd: ragged last CTA pair
    x = torch.rand(b, 128, 96, generator=g) * 2 - 1                   # IPE features live in [-1, 1]
    venc = torch.randn(b, 27, generator=g)
    params = make_state_dict(seed=6, kind="xavier")
    mlp = mp.MLP(net_depth=8, net_width=256, net_depth_condition=1, net_width_condition=128,
                 skip_index=4, num_rgb_channels=3, num_density_channels=1, activation="relu",
                 xyz_dim=96, view_dim=27)
    mlp.load_state_dict({k[len("mlp."):]: v for k, v in params.items()})
    mlp = mlp.to(DEV).eval()
    rgb, dens = mlp(x.to(DEV), venc.to(DEV), precision=precision)
    torch.cuda.synchronize()
    assert rgb.shape == (b, 128, 3) and dens.shape == (b, 128, 1)
    want_rgb, want_dens = oracle.mlp_forward(params, x, venc, operand_dtype=DT[precision])
    f32_rgb, f32_dens = oracle.mlp_forward(params, x, venc)
    # Max-norm agreement with the emulation is NOT tight at the raw-head level: one 16-bit rounding that
    # lands on the other side (fp32 accumulation order differs) perturbs every next-layer input by
    # ~eps/16 and decorrelates that sample's later roundings, so a few samples differ by a full
    # 16-bit-rounding error.  The median is what isolates kernel bugs (layout / barrier / head
    # mistakes are O(1) everywhere); the max is held to the 16-bit bound.
    med_tol, max_tol = (1e-3, 2e-2) if precision == "bf16" else (1.5e-4, 3e-3)
    for name, got, want, ref in (("raw_rgb", rgb, want_rgb, f32_rgb), ("raw_density", dens, want_dens, f32_dens)):
        scale = float(ref.abs().max())
        d_emu = (got.cpu() - want).abs() / scale
        e_f32 = float((got.cpu() - ref).abs().max()) / scale
        print(f"{precision} {name}: err / max|ref|: median {float(d_emu.median()):.3e}, max {float(d_emu.max()):.3e} "
              f"vs emulated oracle; max {e_f32:.3e} vs fp32")
        assert float(d_emu.median()) <= med_tol, (name, float(d_emu.median()))
        assert float(d_emu.max()) <= max_tol, (name, float(d_emu.max()))
        assert e_f32 <= max_tol, (name, e_f32)
    # fp32 entry on the same inputs agrees with the fp32 oracle at the fp32 bar
    rgb32, dens32 = mlp(x.to(DEV), venc.to(DEV))
    assert float((rgb32.cpu() - f32_rgb).abs().max()) <= 1e-4 * float(f32_rgb.abs().max())
    assert float((dens32.cpu() - f32_dens).abs().max()) <= 1e-4 * float(f32_dens.abs().max())


def test_tc_mlp_stage_entry_rejects_other_shapes():
    mlp = mp.MLP(8, 256, 1, 128, 4, 3, 1, "relu", 96, 27).to(DEV)
    x = torch.zeros(4, 64, 96, device=DEV)
    with pytest.raises(NotImplementedError):
        mlp(x, torch.zeros(4, 27, device=DEV), precision="bf16")     # 64 samples/ray: fp32 path only


@pytest.mark.parametrize("randomized", [False, True])
def test_tc_fused_prologue_is_bit_exact(randomized):
    """The fenceposts the level kernels produce themselves (coarse: near/far (+ stratified jitter); fine:
    blur-pool + inverse CDF of the coarse weights) equal, bit for bit, what the stand-alone stage entry points
    (which are pinned bit-exactly to the reference) return for the same inputs — searchsorted indices included."""
    b = 203
    rays = mp.namedtuple_map(lambda t: t.to(DEV), mp.random_ray_batch(b, seed=29, multiscale=True))
    model = mp.MipNerf(precision="bf16")
    model.load_state_dict(make_state_dict(seed=5, kind="trained_like"))
    model = model.to(DEV).eval()
    g = torch.Generator(device=DEV).manual_seed(3)
    t_rand = torch.rand(b, 129, device=DEV, generator=g) if randomized else None
    u_jit = (torch.rand(b, 129, device=DEV, generator=g) * (1 / 129 - 1.2e-7)) if randomized else None
    (c_rgb, _, _, w0, t0, _), (_, _, _, _, t1, inds1) = model(rays, randomized, True, t_rand=t_rand, u_jitter=u_jit,
                                                              return_inds=True)
    want_t0, _ = mp.sample_along_rays(rays.origins, rays.directions, rays.radii, 128, rays.near, rays.far,
                                      randomized, False, "cone", t_rand=t_rand)
    want_t1, _, want_inds = mp.resample_along_rays(rays.origins, rays.directions, rays.radii, t0, w0, randomized,
                                                   "cone", True, 0.01, u_jitter=u_jit, return_inds=True)
    torch.cuda.synchronize()
    assert torch.equal(t0, want_t0)
    assert torch.equal(t1, want_t1)
    assert torch.equal(inds1, want_inds)
    assert torch.isfinite(c_rgb).all()
