"""Tensor-core path for encoding degrees below the kernels' own 16 / 4 (`max_deg_point` 1..16, `deg_view` 1..4).

The level kernels always compute the 96 IPE features of degrees 0..15 (models/mip.py:322-341) and the 27 view features
of degrees 0..3 (:353-363); a model with fewer degrees has narrower layers.0 / layers.5 / view_layers.0, and its weights
are zero-padded to the kernels' widths when the operand image is packed (mlp_tc.cu: expand_encoding_columns_kernel).
Checked against the reference's own forward of such a model (tests/golden/forward_deg10_view2.npz), and bit for bit
against the default-degree model holding the same weights padded by hand.
"""
import numpy as np
import pytest
import torch

from helpers import (FLOORS, RTOL, assert_close, assert_fine_level_close, assert_level_close, golden, golden_levels,
                     golden_rays, make_state_dict, oracle, oracle_rays)

pytestmark = pytest.mark.gpu

import mipnerf_pl_b200 as mp  # noqa: E402

DEV = "cuda:0"
NAME = "forward_deg10_view2.npz"


def build_model(precision, seed, kind, max_deg_point, deg_view):
    model = mp.MipNerf(precision=precision, max_deg_point=max_deg_point, deg_view=deg_view)
    model.load_state_dict(make_state_dict(seed=seed, kind=kind, xyz_dim=6 * max_deg_point, view_dim=6 * deg_view + 3))
    return model.to(DEV).eval()


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_forward_vs_reference_golden(precision):
    """The reference's forward of MipNerf(max_deg_point=10, deg_view=2), at the contract's tolerance: the fp32 path
    (any shape) and the split-operand tensor-core mode (zero-padded operand image)."""
    g = golden(NAME)
    seed, randomized, white = (int(v) for v in g["meta"])
    model = build_model(precision, seed, "trained_like", 10, 2)
    ret = model(golden_rays(g, device=DEV), bool(randomized), bool(white), return_inds=True)
    want = golden_levels(g)
    for lvl, (got, ref) in enumerate(zip(ret, want)):
        if lvl > 0 and precision != "fp32":      # x40 density head: per-ray statement (helpers.assert_fine_level_close)
            st = assert_fine_level_close(got[:5], ref, what=f"{NAME} level {lvl} ")
            print(f"{precision} level {lvl}: " + ", ".join(f"{k} max {v[1]:.2e} ({v[0]} rays > 1e-4)" for k, v in st.items()))
        else:
            errs = assert_level_close(got[:5], ref, rtol=RTOL, what=f"{NAME} level {lvl} ", level=lvl)
            print(f"{precision} level {lvl}: " + ", ".join(f"{k} {v:.2e}" for k, v in errs.items()))
        if lvl > 0:
            mism = float((got[5].cpu().numpy() != g[f"l{lvl}_inds"]).mean())
            print(f"{precision}: {mism:.3%} of the fine level's searchsorted indices differ from the reference's")
            assert mism < 5e-3
    assert torch.equal(ret[0][4].cpu(), torch.from_numpy(want[0][4])), "coarse fenceposts are bit-exact"


def pad_to_default_degrees(sd, L, V):
    """The same network as a default-degree (16 / 4) state_dict: zero columns for the encoding degrees it lacks."""
    out = {k: v.clone() for k, v in sd.items()}

    def expand(w, prefix, half_src, half_dst):
        rows = w.shape[0]
        full = torch.zeros(rows, prefix + 2 * half_dst)
        full[:, :prefix] = w[:, :prefix]
        for h in range(2):
            full[:, prefix + h * half_dst: prefix + h * half_dst + half_src] = \
                w[:, prefix + h * half_src: prefix + (h + 1) * half_src]
        return full
    out["mlp.layers.0.0.weight"] = expand(sd["mlp.layers.0.0.weight"], 0, 3 * L, 48)
    out["mlp.layers.5.0.weight"] = expand(sd["mlp.layers.5.0.weight"], 256, 3 * L, 48)
    out["mlp.view_layers.0.0.weight"] = expand(sd["mlp.view_layers.0.0.weight"], 256 + 3, 3 * V, 12)
    return out


@pytest.mark.parametrize("precision", ["bf16", "fp16", "fp16x3"])
@pytest.mark.parametrize("L,V", [(10, 2), (1, 1), (15, 4), (16, 3)])
def test_padded_image_equals_hand_padded_default_model(precision, L, V):
    """Bit for bit: the narrower model through the padding packer == the default-degree model whose state_dict holds
    the same weights with explicit zero columns (the kernels are the same, so are their operand images)."""
    b = 200
    rays = mp.namedtuple_map(lambda t: t.to(DEV), mp.random_ray_batch(b, seed=13, multiscale=True))
    sd = make_state_dict(seed=8, kind="trained_like", xyz_dim=6 * L, view_dim=6 * V + 3)
    small = mp.MipNerf(precision=precision, max_deg_point=L, deg_view=V)
    small.load_state_dict(sd)
    small = small.to(DEV).eval()
    full = mp.MipNerf(precision=precision)
    full.load_state_dict(pad_to_default_degrees(sd, L, V))
    full = full.to(DEV).eval()
    a, c = small(rays, False, True), full(rays, False, True)
    for lvl in range(2):
        for k in range(5):
            assert torch.equal(a[lvl][k], c[lvl][k]), (precision, L, V, lvl, k)
    # randomized mode with the in-kernel generator takes the same path
    small.rng_seed, small.rng_offset, full.rng_seed, full.rng_offset = 3, 0, 3, 0
    assert torch.equal(small(rays, True, False)[1][0], full(rays, True, False)[1][0])


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_16bit_forward_vs_oracle_with_same_operand_rounding(precision):
    b = 96
    rays = mp.random_ray_batch(b, seed=22, multiscale=True)
    params = make_state_dict(seed=4, kind="xavier", xyz_dim=60, view_dim=15)
    dt = torch.bfloat16 if precision == "bf16" else torch.float16
    want = oracle.forward(params, oracle_rays(rays), False, True, dict(max_deg_point=10, deg_view=2), operand_dtype=dt)
    model = build_model(precision, 4, "xavier", 10, 2)
    got = model(mp.namedtuple_map(lambda t: t.to(DEV), rays), False, True)
    rtol = 2e-3 if precision == "bf16" else 4e-4
    for lvl in range(2):
        assert_close(got[lvl][0], want[lvl][0], FLOORS["comp_rgb"], rtol=rtol, what=f"{precision} level {lvl} comp_rgb")
        assert_close(got[lvl][2], want[lvl][2], FLOORS["acc"], rtol=rtol, what=f"{precision} level {lvl} acc")


def test_training_and_mlp_only_modes_keep_the_default_degrees():
    """What is NOT widened says so: the tensor-core training step and the MLP-only tensor-core entry point take the
    full 96 / 27 encodings; a narrower model gets NotImplementedError there (fp32 works for any degrees)."""
    model = build_model("bf16", 1, "xavier", 10, 2)
    model.train()
    rays = mp.namedtuple_map(lambda t: t.to(DEV), mp.random_ray_batch(32, seed=2))
    rgbs = torch.rand(32, 3, device=DEV)
    with pytest.raises(NotImplementedError):
        mp.forward_backward(model, rays, rgbs, False, True)
    model.precision = "fp32"
    out = mp.forward_backward(model, rays, rgbs, False, True)
    assert np.isfinite(float(out["loss"]))
