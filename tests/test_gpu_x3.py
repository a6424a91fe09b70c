"""The split-operand tensor-core parity mode (`precision='fp16x3'`, MIPNERF_B200_FP16X3) on real hardware.

north_star: "rendered RGB within 1e-4 of the reference" with the MLP on tcgen05.  The reference computes every
nn.Linear in fp32 (models/mip_nerf.py:94-110); plain 16-bit operands miss that bar on trained-like weights
(bf16 1.9e-3, fp16 2.8e-4 absolute on RGB).  The split mode carries every operand as hi + lo fp16 halves and issues
hi.hi + lo.hi + hi.lo per K step (22 significant bits), and must pass THE SAME checks as the fp32 parity path:
the committed reference goldens through `helpers.assert_level_close` at RTOL = 1e-4 with the fp32 floors.
"""
import numpy as np
import pytest
import torch

from helpers import (FLOORS, RTOL, assert_fine_level_close, assert_level_close, golden, golden_levels, golden_rays,
                     make_state_dict, oracle, oracle_rays, rel_err)

pytestmark = pytest.mark.gpu

import mipnerf_pl_b200 as mp  # noqa: E402

DEV = "cuda:0"


def cuda(x):
    return torch.from_numpy(x).to(DEV) if isinstance(x, np.ndarray) else x.to(DEV)


def build_model(kind, seed, precision):
    model = mp.MipNerf(precision=precision)
    model.load_state_dict(make_state_dict(seed=seed, kind=kind))
    return model.to(DEV).eval()


@pytest.mark.parametrize("name,kind", [
    ("forward_xavier.npz", "xavier"),
    ("forward_trained_like.npz", "trained_like"),
    ("forward_randomized.npz", "trained_like"),
])
def test_fp16x3_forward_vs_reference_golden(name, kind):
    """The committed outputs of the reference's own forward, at the contract's tolerance."""
    g = golden(name)
    seed, randomized, white = (int(v) for v in g["meta"])
    model = build_model(kind, seed, "fp16x3")
    rays = golden_rays(g, device=DEV)
    ret = model(rays, bool(randomized), bool(white), t_rand=cuda(g["t_rand"]) if "t_rand" in g else None,
                u_jitter=cuda(g["u_jitter"]) if "u_jitter" in g else None, return_inds=True)
    want = golden_levels(g)
    assert len(ret) == len(want) == 2
    for lvl, (got, ref) in enumerate(zip(ret, want)):
        if lvl > 0 and kind == "trained_like":   # see helpers.assert_fine_level_close: per-element statement
            st = assert_fine_level_close(got[:5], ref, what=f"{name} level {lvl} ")
            print(f"fp16x3 {name} level {lvl}: " + ", ".join(f"{k} max {v[1]:.2e} ({v[0]} rays > 1e-4)" for k, v in st.items()))
        else:
            errs = assert_level_close(got[:5], ref, rtol=RTOL, what=f"{name} level {lvl} ", level=lvl)
            print(f"fp16x3 {name} level {lvl}: " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))
        if lvl > 0:
            mism = float((got[5].cpu().numpy() != g[f"l{lvl}_inds"]).mean())
            print(f"fp16x3 {name}: {mism:.3%} of the fine level's searchsorted indices differ from the reference's")
            assert mism < 5e-3
    assert torch.equal(ret[0][4].cpu(), torch.from_numpy(want[0][4])), "coarse fenceposts are bit-exact"


@pytest.mark.parametrize("kind", ["xavier", "trained_like"])
def test_fp16x3_forward_vs_oracle(kind):
    """Other rays (multi-scale mix, ragged count) against the fp32 oracle at 1e-4, and against the oracle with the
    split arithmetic emulated on the CPU (isolates kernel bugs from rounding) 4x tighter."""
    rays = mp.random_ray_batch(301, seed=11, multiscale=True)
    params = make_state_dict(seed=4, kind=kind)
    want = oracle.forward(params, oracle_rays(rays), False, True)
    emu = oracle.forward(params, oracle_rays(rays), False, True, operand_dtype=torch.float16, operand_split=True)
    got = build_model(kind, 4, "fp16x3")(mp.namedtuple_map(lambda t: t.to(DEV), rays), False, True)
    for lvl in range(2):
        if lvl > 0 and kind == "trained_like":
            st = assert_fine_level_close(got[lvl], want[lvl], what=f"{kind} level {lvl} ")
            print(f"fp16x3 {kind} level {lvl} vs fp32 oracle: " + ", ".join(f"{k} max {v[1]:.2e} ({v[0]} rays > 1e-4)" for k, v in st.items()))
            st = assert_fine_level_close(got[lvl], emu[lvl], what=f"{kind} level {lvl} (emulated) ")
            print(f"fp16x3 {kind} level {lvl} vs emulated split oracle: " + ", ".join(f"{k} max {v[1]:.2e}" for k, v in st.items()))
            continue
        errs = assert_level_close(got[lvl], want[lvl], rtol=RTOL, what=f"{kind} level {lvl} ", level=lvl)
        print(f"fp16x3 {kind} level {lvl} vs fp32 oracle: " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))
        errs = assert_level_close(got[lvl], emu[lvl], rtol=5e-5, what=f"{kind} level {lvl} (emulated) ", level=lvl)
        print(f"fp16x3 {kind} level {lvl} vs emulated split oracle: " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))


@pytest.mark.parametrize("precision,bound", [("fp16x3", 1e-5), ("bf16x3", 2e-4)])
def test_x3_mlp_stage_entry(precision, bound):
    """MLP.forward alone in the split modes (models/mip_nerf.py:75-111) vs the fp32 oracle, raw heads."""
    g = torch.Generator().manual_seed(31)
    b = 37
    x = torch.rand(b, 128, 96, generator=g) * 2 - 1
    venc = torch.randn(b, 27, generator=g)
    params = make_state_dict(seed=6, kind="xavier")
    mlp = mp.MLP(8, 256, 1, 128, 4, 3, 1, "relu", 96, 27)
    mlp.load_state_dict({k[len("mlp."):]: v for k, v in params.items()})
    mlp = mlp.to(DEV).eval()
    rgb, dens = mlp(x.to(DEV), venc.to(DEV), precision=precision)
    torch.cuda.synchronize()
    f32_rgb, f32_dens = oracle.mlp_forward(params, x, venc)
    for name, got, ref in (("raw_rgb", rgb, f32_rgb), ("raw_density", dens, f32_dens)):
        e = float((got.cpu() - ref).abs().max()) / float(ref.abs().max())
        print(f"{precision} {name}: max err / max|ref| = {e:.3e} vs fp32 oracle")
        assert e <= bound, (precision, name, e)


def test_bf16x3_forward_is_16_bit_accurate():
    """bf16 halves keep 16 significant bits with fp32's exponent range: on the stress golden its RGB is within 1e-4
    at the coarse level and 5e-4 at the fine level (measured 7.7e-5 / 2.5e-4) — 25x better than plain bf16, but not
    the mode the 1e-4 contract is claimed on (that is fp16x3)."""
    g = golden("forward_trained_like.npz")
    seed, randomized, white = (int(v) for v in g["meta"])
    ret = build_model("trained_like", seed, "bf16x3")(golden_rays(g, device=DEV), bool(randomized), bool(white))
    want = golden_levels(g)
    for lvl in range(2):
        e = rel_err(ret[lvl][0].cpu().numpy(), want[lvl][0], FLOORS["comp_rgb"])
        print(f"bf16x3 level {lvl} comp_rgb rel err {e:.3e}")
        assert e <= (1e-4 if lvl == 0 else 5e-4)


def test_x3_ragged_sizes_and_split_invariance():
    model = build_model("trained_like", 1, "fp16x3")
    rays = mp.namedtuple_map(lambda t: t.to(DEV), mp.random_ray_batch(599, seed=3))
    full = model(rays, False, True)
    for cut in (1, 2, 3, 297):
        a = model(mp.Rays(*[f[:cut] for f in rays]), False, True)
        b = model(mp.Rays(*[f[cut:] for f in rays]), False, True)
        for lvl in range(2):
            for k in range(5):
                assert torch.equal(torch.cat([a[lvl][k], b[lvl][k]]), full[lvl][k]), (cut, lvl, k)
    assert model(mp.Rays(*[f[:0] for f in rays]), False, True)[1][0].shape == (0, 3)


def test_x3_full_batch_properties():
    model = build_model("trained_like", 9, "fp16x3")
    rays = mp.namedtuple_map(lambda t: t.to(DEV), mp.random_ray_batch(4096, seed=0))
    white, black = model(rays, False, True), model(rays, False, False)
    f32 = build_model("trained_like", 9, "fp32")(rays, False, True)
    for lvl in range(2):
        rgb_w, dist, acc, w, t = white[lvl]
        assert torch.isfinite(rgb_w).all() and torch.all(w >= 0) and torch.all(acc <= 1 + 1e-5)
        assert torch.all(t[:, 1:] >= t[:, :-1])
        assert torch.allclose(rgb_w, black[lvl][0] + (1 - acc)[:, None], atol=1e-6)
        assert torch.allclose(w.sum(-1), acc, atol=1e-5)
        # at BASELINE configs[1] size: the tensor-core parity mode against the CUDA-core fp32 path
        e = rel_err(rgb_w.cpu().numpy(), f32[lvl][0].cpu().numpy(), FLOORS["comp_rgb"])
        print(f"fp16x3 vs fp32 path, 4096 rays, level {lvl}: comp_rgb rel err {e:.3e}")
        assert e <= RTOL


def test_x3_is_forward_only():
    model = build_model("xavier", 0, "fp16x3")
    rays = mp.namedtuple_map(lambda t: t.to(DEV), mp.random_ray_batch(8, seed=0))
    with pytest.raises(NotImplementedError):
        mp.forward_backward(model, rays, torch.zeros(8, 3, device=DEV), False, True)
