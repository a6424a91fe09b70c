"""Training-step pieces that run without a GPU: the oracle's loss/gradients and the LR schedule pinned
against tests/golden/training.npz (reference MipNerf + reference distloss + torch autograd, reference
MipLRDecay, torch.optim.Adam), and the host-side schedule / gradient all-reduce helper."""
import numpy as np
import pytest
import torch

from helpers import (GRAD_RTOL_TRUNK, assert_grads_match_golden, golden, make_state_dict, oracle, oracle_rays,
                     training_golden_case)

import mipnerf_pl_b200 as mp


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_training_loss_and_grads_match_reference(tag):
    g = golden("training.npz")
    rays, rgbs, randomized, white, disable_ms, t_rand, u_jit, seed = training_golden_case(g, tag)
    params = {k: v.clone().requires_grad_(True) for k, v in make_state_dict(seed=seed, kind="trained_like").items()}
    loss, mses, dls, _ = oracle.training_loss(params, oracle_rays(rays), rgbs, randomized, white,
                                              coarse_loss_mult=0.1, disable_multiscale_loss=disable_ms,
                                              t_rand=t_rand, u_jitter=u_jit)
    loss.backward()
    got = np.array([float(loss.detach())] + [float(x.detach()) for x in mses] + [float(x.detach()) for x in dls])
    np.testing.assert_allclose(got, g[f"{tag}_loss"], rtol=1e-5)
    worst = assert_grads_match_golden({k: v.grad for k, v in params.items()}, g, tag, rtol=1e-5)
    print(f"oracle vs reference autograd, case {tag}: worst per-tensor gradient error {worst:.2e}")


def test_reference_trunk_gradients_are_fp32_noise_limited():
    """Why helpers.GRAD_RTOL_TRUNK is 2e-3: the same autograd graph evaluated in float64 moves the fp32 trunk
    gradients by more than that bar (ReLU masks at round-off, ulp motion of the resampled fenceposts through the
    IPE), while the heads, which see no mask downstream, agree to ~1e-5."""
    b = 24
    rays = mp.random_ray_batch(b, seed=13, multiscale=True)
    rgbs = torch.rand(b, 3, generator=torch.Generator().manual_seed(5))
    grads = {}
    for dt in (torch.float32, torch.float64):
        params = {k: v.clone().to(dt).requires_grad_(True) for k, v in make_state_dict(seed=9, kind="xavier").items()}
        r = oracle.Rays(*[getattr(rays, k).to(dt) for k in oracle.Rays._fields])
        loss, _, _, _ = oracle.training_loss(params, r, rgbs.to(dt), False, True)
        loss.backward()
        grads[dt] = {k: v.grad.double() for k, v in params.items()}
    err = {k: float((grads[torch.float32][k] - grads[torch.float64][k]).norm() / grads[torch.float64][k].norm())
           for k in grads[torch.float32]}
    print({k.replace("mlp.", ""): float(f"{v:.1e}") for k, v in err.items() if k.endswith("weight")})
    assert err["mlp.layers.0.0.weight"] > 0.5 * GRAD_RTOL_TRUNK
    assert err["mlp.color_layer.weight"] < 1e-4 and err["mlp.density_layer.weight"] < 1e-4


def test_lr_schedule_matches_reference():
    g = golden("training.npz")
    for step, want in zip(g["lr_steps"], g["lr_values"]):
        for fn in (oracle.mip_lr, mp.mip_lr):
            assert fn(int(step), 5e-4, 5e-6, 1000000, 2500, 0.01) == pytest.approx(float(want), rel=1e-12)
    for step, want in zip((0, 50, 100), g["lr_nodelay_values"]):
        assert mp.mip_lr(step, 1e-3, 1e-5, 100, 0, 1.0) == pytest.approx(float(want), rel=1e-12)


def test_scheduler_class_tracks_reference_trajectory():
    """MipLRDecay driven like Lightning does (optimizer.step(); scheduler.step()) reproduces the lrs the
    reference scheduler produced next to torch.optim.Adam."""
    g = golden("training.npz")
    p = torch.nn.Parameter(torch.from_numpy(g["adam_p0"]).clone())
    opt = torch.optim.Adam([p], lr=5e-4)                   # CPU torch Adam: the class under test is the scheduler
    sched = mp.MipLRDecay(opt, 5e-4, 5e-6, 10, 4, 0.01)
    for i in range(4):
        assert opt.param_groups[0]["lr"] == pytest.approx(float(g["adam_lrs"][i]), rel=1e-12)
        p.grad = torch.from_numpy(g["adam_grads"][i]).clone()
        opt.step()
        sched.step()
        np.testing.assert_allclose(p.detach().numpy(), g["adam_traj"][i], rtol=1e-6, atol=1e-9)


def test_training_refuses_cpu_tensors_and_other_modes():
    model = mp.MipNerf()
    rays = mp.random_ray_batch(4, seed=0)
    with pytest.raises(RuntimeError):                       # no CPU fallback
        mp.forward_backward(model, rays, torch.zeros(4, 3), False, True)
    with pytest.raises(NotImplementedError):
        mp.forward_backward(mp.MipNerf(stop_resample_grad=False), rays, torch.zeros(4, 3), False, True)


def test_system_configure_optimizers_shapes():
    system = mp.MipNeRFSystem(mp.default_hparams())
    (opt,), (sched,) = system.configure_optimizers()
    assert isinstance(opt, mp.FusedAdam) and sched["interval"] == "step"
    assert opt.param_groups[0]["lr"] == pytest.approx(mp.mip_lr(0, 5e-4, 5e-6, 1000000, 2500, 0.01))
    assert sum(p.numel() for g_ in opt.param_groups for p in g_["params"]) == 612740


def test_fused_adam_resumes_from_a_torch_adam_checkpoint():
    """The reference resumes through Lightning, which restores torch.optim.Adam's state (float32 tensor `step`, no
    `grad_scale` in the param group): loading it into FusedAdam must leave a state step() can consume."""
    import mipnerf_pl_b200 as mp
    p_ref = [torch.nn.Parameter(torch.ones(5, 3)), torch.nn.Parameter(torch.zeros(3))]
    adam = torch.optim.Adam(p_ref, lr=5e-4)
    for _ in range(3):
        for p in p_ref:
            p.grad = torch.full_like(p, 0.1)
        adam.step()
    p_new = [torch.nn.Parameter(p.detach().clone()) for p in p_ref]
    fused = mp.FusedAdam(p_new, lr=5e-4)
    fused.load_state_dict(adam.state_dict())
    assert "grad_scale" not in fused.param_groups[0] or fused.param_groups[0]["grad_scale"] == 1.0
    for p in p_new:
        st = fused.state[p]
        assert mp.FusedAdam._step_count(st) == 3
        assert st["exp_avg"].shape == p.shape and st["exp_avg_sq"].shape == p.shape
