"""In-kernel Philox draws for randomized=True (replaces torch.rand / uniform_, models/mip.py:159, :201-202)."""
import numpy as np
import pytest
import torch

from helpers import make_state_dict

pytestmark = pytest.mark.gpu

import mipnerf_pl_b200 as mp  # noqa: E402

DEV = "cuda:0"
EPS = float(torch.finfo(torch.float32).eps)


def test_philox_uniform_distribution():
    b, n = 4096, 129
    u = mp.philox_uniform(1234, 0, 0, b, n, DEV).cpu().double().numpy()
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 2e-3 and abs(u.var() - 1 / 12) < 1e-3
    hist = np.histogram(u, bins=64, range=(0, 1))[0] / u.size
    assert np.abs(hist - 1 / 64).max() < 1.5e-3                         # flat
    assert abs(np.corrcoef(u[:, :-1].ravel(), u[:, 1:].ravel())[0, 1]) < 5e-3     # adjacent draws of a ray
    assert abs(np.corrcoef(u[:-1].ravel(), u[1:].ravel())[0, 1]) < 5e-3           # adjacent rays
    j = mp.philox_uniform(1234, 0, 2, b, n, DEV).cpu().double().numpy()            # a u_jitter stream
    top = 1.0 / n - EPS
    assert j.min() >= 0.0 and j.max() < top and abs(j.mean() - top / 2) < 2e-3 * top
    assert not np.array_equal(u, mp.philox_uniform(1234, 1, 0, b, n, DEV).cpu().numpy())   # offset advances
    assert not np.array_equal(u, mp.philox_uniform(1235, 0, 0, b, n, DEV).cpu().numpy())   # seed matters
    assert np.array_equal(u, mp.philox_uniform(1234, 0, 0, b, n, DEV).cpu().double().numpy())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_in_kernel_draws_equal_injected_draws(precision):
    """forward(randomized=True) with the in-kernel generator == forward with the same uniforms injected as arrays,
    bit for bit (fused prologue of the tensor-core kernels and the stand-alone kernels of the fp32 path alike); the
    result does not depend on how the batch is split."""
    b = 300
    model = mp.MipNerf(precision=precision)
    model.load_state_dict(make_state_dict(seed=3, kind="trained_like"))
    model = model.to(DEV).eval()
    rays = mp.namedtuple_map(lambda t: t.to(DEV), mp.random_ray_batch(b, seed=4, multiscale=True))
    model.rng_seed, model.rng_offset = 99, 7
    got = model(rays, True, True)
    assert model.rng_offset == 8
    t_rand = mp.philox_uniform(99, 7, 0, b, 129, DEV)
    u_jit = mp.philox_uniform(99, 7, 2, b, 129, DEV)                 # level 1's stream
    want = model(rays, True, True, t_rand=t_rand, u_jitter=u_jit)
    for lvl in range(2):
        for k in range(5):
            assert torch.equal(got[lvl][k], want[lvl][k]), (precision, lvl, k)
    model.rng_offset = 7
    again = model(rays, True, True)
    assert torch.equal(again[1][0], got[1][0])
    model.rng_offset = 8
    other = model(rays, True, True)
    assert not torch.equal(other[1][4], got[1][4])                   # fresh noise per call
    # stratified: every coarse fencepost stays inside its bin, fine fenceposts stay sorted
    det = model(rays, False, True)
    t_det, t_r = det[0][4], got[0][4]
    mids = 0.5 * (t_det[:, 1:] + t_det[:, :-1])
    lower = torch.cat([t_det[:, :1], mids], -1)
    upper = torch.cat([mids, t_det[:, -1:]], -1)
    assert torch.all(t_r >= lower - 1e-6) and torch.all(t_r <= upper + 1e-6)
    assert torch.all(got[1][4][:, 1:] >= got[1][4][:, :-1])


@pytest.mark.parametrize("precision,n", [("fp32", 64), ("bf16", 64), ("bf16", 4096 + 37)])
def test_training_step_draws_in_kernel(precision, n):
    """fp32 path and the fused tensor-core step (whose forward is the level kernel with its own in-kernel draws); 4133
    rays cross the 4096-ray chunk: the Philox counter is the ray index of the whole batch, not of the chunk."""
    model = mp.MipNerf(precision=precision)
    model.load_state_dict(make_state_dict(seed=1))
    model = model.to(DEV)
    rays = mp.namedtuple_map(lambda t: t.to(DEV), mp.random_ray_batch(n, seed=2))
    rgbs = torch.rand(n, 3, device=DEV)
    model.rng_seed, model.rng_offset = 5, 0
    a = float(mp.forward_backward(model, rays, rgbs, True, True)["loss"])
    g_a = [p.grad.clone() for p in model.parameters()]
    model.rng_offset = 0
    b_ = float(mp.forward_backward(model, rays, rgbs, True, True)["loss"])
    assert a == b_ and all(torch.equal(x, p.grad) for x, p in zip(g_a, model.parameters()))
    t_rand, u_jit = mp.philox_uniform(5, 0, 0, n, 129, DEV), mp.philox_uniform(5, 0, 2, n, 129, DEV)
    c = float(mp.forward_backward(model, rays, rgbs, True, True, t_rand=t_rand, u_jitter=u_jit)["loss"])
    assert a == c
