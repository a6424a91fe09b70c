"""CPU oracle for the Mip-NeRF per-ray hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mipnerf_pl_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker / the
CPU arm, never as the product path.

It is a restatement (fp32, torch-CPU element-wise math + explicit numpy for the
two order-sensitive reductions) of the algorithm in the reference repository
hjxwhy/mipnerf_pl:

  * ``models/mip.py``       (ray math)      -> functions below, each citing file:line
  * ``models/mip_nerf.py``  (MLP + level loop)

Parity pinning: the reference ships NO tests or golden vectors,
so the oracle is pinned against outputs of the reference itself, generated in
the build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference``) and committed as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks every function here against them.

Order-sensitive numerics that decide the resampler's *indices* are written out
explicitly so the oracle does not depend on the ISA torch dispatches to:

  * ``torch.sum(float32, -1)`` over a contiguous row == 32 strided fp32
    accumulators (4 ILP x 8 lanes), ILP partials added in order, lanes added in
    order  -> ``rowsum_f32``;
  * ``torch.cumsum(float32)`` on CPU == sequential float64 accumulation, each
    prefix rounded to float32 -> ``cumsum_f32``.
"""
from __future__ import annotations

import collections
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# Field order of the reference's ray container (datasets/datasets.py:13-16).
Rays = collections.namedtuple(
    "Rays", ("origins", "directions", "viewdirs", "radii", "lossmult", "near", "far"))

HALF_PI_F32 = torch.tensor(np.float32(0.5) * np.float32(np.pi))  # 0x3FC90FDB
F32_EPS = float(torch.finfo(torch.float32).eps)


# --------------------------------------------------------------------------
# order-sensitive reductions (see module docstring)
# --------------------------------------------------------------------------
def rowsum_f32(x: torch.Tensor) -> torch.Tensor:
    """torch.sum(x, -1, keepdim=True) for contiguous fp32 rows, n % 32 == 0,
    with the accumulation order written out (models/mip.py:182)."""
    a = x.detach().cpu().numpy().astype(np.float32, copy=False)
    n = a.shape[-1]
    if n % 32 != 0:
        raise ValueError("rowsum_f32 restates the vectorised order only for n % 32 == 0")
    acc = a[..., 0:32].copy()
    for i in range(32, n, 32):
        acc = acc + a[..., i:i + 32]
    acc = acc.reshape(a.shape[:-1] + (4, 8))
    lanes = acc[..., 0, :]
    for k in range(1, 4):
        lanes = lanes + acc[..., k, :]
    s = lanes[..., 0]
    for j in range(1, 8):
        s = s + lanes[..., j]
    return torch.from_numpy(np.ascontiguousarray(s[..., None]))


class _CumsumF32(torch.autograd.Function):
    """Forward: the bit pattern of torch.cumsum(float32) on CPU; backward: its adjoint (reversed cumsum),
    so the oracle's training_loss differentiates through the transmittance like autograd does in the
    reference (models/mip.py:385)."""

    @staticmethod
    def forward(ctx, x):
        a = x.detach().cpu().numpy().astype(np.float64)
        return torch.from_numpy(np.cumsum(a, axis=-1).astype(np.float32)).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return torch.flip(torch.cumsum(torch.flip(g, [-1]), dim=-1), [-1])


def cumsum_f32(x: torch.Tensor) -> torch.Tensor:
    """torch.cumsum(x, -1) on CPU fp32: float64 running sum, rounded per prefix
    (models/mip.py:189, :385)."""
    if x.dtype == torch.float64:        # float64 referee runs of the oracle (tests only)
        return torch.cumsum(x, dim=-1)
    return _CumsumF32.apply(x)


def conical_frustum_moments(t0, t1, radii):
    """Stable conical-frustum moments (models/mip.py:65-72).
    t0, t1: [B, N]; radii: [B, 1] -> t_mean, t_var, r_var each [B, N]."""
    mu = (t0 + t1) / 2
    hw = (t1 - t0) / 2
    denom = 3 * mu ** 2 + hw ** 2
    t_mean = mu + (2 * mu * hw ** 2) / denom
    t_var = (hw ** 2) / 3 - (4 / 15) * ((hw ** 4 * (12 * mu ** 2 - hw ** 2)) / denom ** 2)
    r_var = radii ** 2 * ((mu ** 2) / 4 + (5 / 12) * hw ** 2 - 4 / 15 * (hw ** 4) / denom)
    return t_mean, t_var, r_var


def lift_gaussian_diag(directions, t_mean, t_var, r_var):
    """Diagonal branch of lift_gaussian (models/mip.py:22-36).
    directions [B,3]; moments [B,N] -> mean [B,N,3], cov_diag [B,N,3]."""
    mean = directions[:, None, :] * t_mean[:, :, None]
    d_sq = directions ** 2
    d_norm = torch.sum(d_sq, dim=-1, keepdim=True) + 1e-10
    null_diag = 1 - d_sq / d_norm
    cov = t_var[:, :, None] * d_sq[:, None, :] + r_var[:, :, None] * null_diag[:, None, :]
    return mean, cov


def cast_rays(t_samples, origins, directions, radii, ray_shape="cone"):
    """Fenceposts -> per-interval Gaussians (models/mip.py:81-103)."""
    if ray_shape != "cone":
        raise NotImplementedError(ray_shape)  # models/mip.py:97-98
    t0, t1 = t_samples[..., :-1], t_samples[..., 1:]
    t_mean, t_var, r_var = conical_frustum_moments(t0, t1, radii)
    means, covs = lift_gaussian_diag(directions, t_mean, t_var, r_var)
    return means + origins[:, None, :], covs


def coarse_fenceposts(near, far, num_samples, batch, randomized=False, disparity=False,
                      t_rand: Optional[torch.Tensor] = None):
    """Level-0 fenceposts (models/mip.py:145-163).  ``t_rand`` [B, N+1] in [0,1)
    replaces the reference's ``torch.rand`` so tests can inject identical noise."""
    s = torch.linspace(0.0, 1.0, num_samples + 1)
    if disparity:
        t = 1.0 / (1.0 / near * (1.0 - s) + 1.0 / far * s)
    else:
        t = near + (far - near) * s
    if randomized:
        mids = 0.5 * (t[..., 1:] + t[..., :-1])
        upper = torch.cat([mids, t[..., -1:]], -1)
        lower = torch.cat([t[..., :1], mids], -1)
        if t_rand is None:
            t_rand = torch.rand(batch, num_samples + 1)
        t = lower + (upper - lower) * t_rand
    else:
        t = torch.broadcast_to(t, (batch, num_samples + 1))
    return t


def sample_along_rays(origins, directions, radii, num_samples, near, far, randomized,
                      disparity, ray_shape, t_rand=None):
    """models/mip.py:127-165."""
    t = coarse_fenceposts(near, far, num_samples, origins.shape[0], randomized, disparity, t_rand)
    return t, cast_rays(t, origins, directions, radii, ray_shape)


def integrated_pos_enc(means, covs, min_deg, max_deg):
    """Diagonal IPE, first output of expected_sin only
    (models/mip.py:335-341, :350, :286).  Feature order: scale-major then xyz,
    48 'sin' then 48 'cos' where cos is sin(fl32(y + fl32(pi/2)))."""
    scales = torch.tensor([2.0 ** i for i in range(min_deg, max_deg)], dtype=torch.float32)
    shp = means.shape[:-1] + (-1,)
    y = (means[..., None, :] * scales[:, None]).reshape(shp)
    yv = (covs[..., None, :] * scales[:, None] ** 2).reshape(shp)
    x = torch.cat([y, y + HALF_PI_F32], dim=-1)
    xv = torch.cat([yv, yv], dim=-1)
    return torch.exp(-0.5 * xv) * torch.sin(x)


def pos_enc(x, min_deg, max_deg, append_identity=True):
    """View-direction encoding (models/mip.py:353-363)."""
    scales = torch.tensor([2.0 ** i for i in range(min_deg, max_deg)], dtype=torch.float32)
    xb = (x[..., None, :] * scales[:, None]).reshape(x.shape[:-1] + (-1,))
    feat = torch.sin(torch.cat([xb, xb + HALF_PI_F32], dim=-1))
    return torch.cat([x, feat], dim=-1) if append_identity else feat


def volumetric_rendering(rgb, density, t_samples, dirs, white_bkgd):
    """Front-to-back alpha compositing (models/mip.py:366-401).  ``distance``
    is NOT divided by acc (unlike the JAX original)."""
    t_mids = 0.5 * (t_samples[..., :-1] + t_samples[..., 1:])
    delta = (t_samples[..., 1:] - t_samples[..., :-1]) * torch.linalg.norm(dirs[:, None, :], dim=-1)
    dd = density[..., 0] * delta
    alpha = 1 - torch.exp(-dd)
    trans = torch.exp(-torch.cat([torch.zeros_like(dd[..., :1]), cumsum_f32(dd[..., :-1])], dim=-1))
    weights = alpha * trans
    comp_rgb = (weights[..., None] * rgb).sum(dim=-2)
    acc = weights.sum(dim=-1)
    distance = (weights * t_mids).sum(dim=-1)
    distance = torch.clamp(torch.nan_to_num(distance), t_samples[:, 0], t_samples[:, -1])
    if white_bkgd:
        comp_rgb = comp_rgb + (1.0 - acc[..., None])
    return comp_rgb, distance, acc, weights


def distloss(weight, samples):
    """Distortion loss (models/mip.py:8-20): weight [B,N], samples [B,N+1] -> scalar."""
    interval = samples[:, 1:] - samples[:, :-1]
    mid = (samples[:, 1:] + samples[:, :-1]) * 0.5
    loss_uni = (1 / 3) * (interval * weight.pow(2)).sum(-1).mean()
    ww = weight.unsqueeze(-1) * weight.unsqueeze(-2)
    mm = (mid.unsqueeze(-1) - mid.unsqueeze(-2)).abs()
    return loss_uni + (ww * mm).sum((-1, -2)).mean()


def blurpool_weights(weights, resample_padding):
    """Max-then-average blur + constant (models/mip.py:252-257)."""
    wp = torch.cat([weights[..., :1], weights, weights[..., -1:]], dim=-1)
    wm = torch.maximum(wp[..., :-1], wp[..., 1:])
    return 0.5 * (wm[..., :-1] + wm[..., 1:]) + resample_padding


def deterministic_u(num_samples):
    """torch.linspace(0, 1-eps, n) on CPU: step = fl32(fl32(1-eps)/(n-1)); first half fl32(step*j), second half
    fl32(end - step*(n-1-j)) with one rounding (equal to fl32(j*step) when n-1 is a power of two)
    (models/mip.py:206-208)."""
    return torch.linspace(0.0, 1.0 - F32_EPS, num_samples)


def sorted_piecewise_constant_pdf(bins, weights, num_samples, randomized=False,
                                  u_jitter: Optional[torch.Tensor] = None,
                                  return_inds=False):
    """Inverse-CDF sampling from a histogram (models/mip.py:168-229).
    ``weights`` must be a fresh tensor (it is modified like the reference does).
    ``u_jitter`` [B, num_samples] in [0, 1/num_samples - eps) replaces the
    reference's ``uniform_`` draw."""
    eps = 1e-5
    weights = weights.clone()
    weight_sum = rowsum_f32(weights)                                   # :182
    padding = torch.maximum(torch.zeros_like(weight_sum), eps - weight_sum)
    weights += padding / weights.shape[-1]
    weight_sum = weight_sum + padding
    pdf = weights / weight_sum
    cdf = torch.minimum(torch.ones(()), cumsum_f32(pdf[..., :-1]))     # :189-190
    lead = list(cdf.shape[:-1])
    cdf = torch.cat([torch.zeros(lead + [1]), cdf, torch.ones(lead + [1])], dim=-1)
    if randomized:
        s = 1 / num_samples
        u = (torch.arange(num_samples) * s)[None, ...]
        if u_jitter is None:
            u_jitter = torch.empty(lead + [num_samples]).uniform_(to=(s - F32_EPS))
        u = u + u_jitter
        u = torch.minimum(u, torch.full_like(u, 1.0 - F32_EPS))
    else:
        u = torch.broadcast_to(deterministic_u(num_samples), lead + [num_samples])
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)                      # :211
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bins_b, bins_a = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    samples = bins_b + t * (bins_a - bins_b)
    return (samples, inds) if return_inds else samples


def resample_along_rays(origins, directions, radii, t_samples, weights, randomized, ray_shape,
                        stop_grad, resample_padding, u_jitter=None, return_inds=False):
    """models/mip.py:232-280 (stop_grad only changes autograd, not values)."""
    w = blurpool_weights(weights, resample_padding)
    out = sorted_piecewise_constant_pdf(t_samples, w, t_samples.shape[-1], randomized,
                                        u_jitter=u_jitter, return_inds=return_inds)
    new_t, inds = out if return_inds else (out, None)
    mc = cast_rays(new_t, origins, directions, radii, ray_shape)
    return (new_t, mc, inds) if return_inds else (new_t, mc)


# --------------------------------------------------------------------------
# models/mip_nerf.py
# --------------------------------------------------------------------------
DEFAULT_CONFIG = dict(
    num_samples=128, num_levels=2, resample_padding=0.01, stop_resample_grad=True,
    use_viewdirs=True, disparity=False, ray_shape="cone", min_deg_point=0, max_deg_point=16,
    deg_view=4, density_noise=0.0, density_bias=-1.0, rgb_padding=0.001,
    disable_integration=False, mlp_net_depth=8, mlp_net_width=256, mlp_net_depth_condition=1,
    mlp_net_width_condition=128, mlp_skip_index=4, mlp_num_rgb_channels=3,
    mlp_num_density_channels=1)


def _mlp_forward_16bit(params, x, view_enc, net_depth, skip_index, prefix, dt, split=False):
    """MLP.forward with the operand rounding of the tensor-core kernels emulated (fp32 accumulate):
    trunk / bottleneck / view-layer GEMM operands (activations AND weights) rounded to `dt`;
    density head, view-direction term and colour head stay fp32 — exactly the split documented in
    mipnerf_pl_b200/csrc/mlp_tc.cu.  `split=True` emulates the split-operand ("x3") modes: each operand is
    hi = fl16(x), lo = fl16(x - hi) and the product is hi.hi + lo.hi + hi.lo.  Test infrastructure only."""
    def r(t):
        return t.to(dt).to(torch.float32)

    def lin(a, w):
        if not split:
            return F.linear(r(a), r(w))
        ah, wh = r(a), r(w)
        al, wl = r(a - ah), r(w - wh)
        return F.linear(ah, wh) + F.linear(al, wh) + F.linear(ah, wl)
    inputs = x
    for i in range(net_depth):
        x = F.relu(lin(x, params[f"{prefix}layers.{i}.0.weight"]) + params[f"{prefix}layers.{i}.0.bias"])
        if i % skip_index == 0 and i > 0:
            x = torch.cat([x, inputs], dim=-1)
    raw_density = F.linear(x, params[f"{prefix}density_layer.weight"], params[f"{prefix}density_layer.bias"])
    bott = lin(x, params[f"{prefix}extra_layer.weight"]) + params[f"{prefix}extra_layer.bias"]
    wv, bv = params[f"{prefix}view_layers.0.0.weight"], params[f"{prefix}view_layers.0.0.bias"]
    k = bott.shape[-1]
    v = F.relu(lin(bott, wv[:, :k]) + (F.linear(view_enc, wv[:, k:]) + bv)[:, None, :])
    raw_rgb = F.linear(v, params[f"{prefix}color_layer.weight"], params[f"{prefix}color_layer.bias"])
    return raw_rgb, raw_density


def mlp_forward(params: Dict[str, torch.Tensor], x, view_enc, net_depth=8, skip_index=4,
                net_depth_condition=1, prefix="mlp.", operand_dtype=None, operand_split=False):
    """MLP.forward (models/mip_nerf.py:75-111) over a state_dict-style mapping
    with the reference's key names (``mlp.layers.{i}.0.weight`` ...)."""
    if operand_dtype is not None:
        assert view_enc is not None and net_depth_condition == 1
        return _mlp_forward_16bit(params, x, view_enc, net_depth, skip_index, prefix, operand_dtype, operand_split)
    inputs = x
    for i in range(net_depth):
        x = F.relu(F.linear(x, params[f"{prefix}layers.{i}.0.weight"], params[f"{prefix}layers.{i}.0.bias"]))
        if i % skip_index == 0 and i > 0:
            x = torch.cat([x, inputs], dim=-1)
    raw_density = F.linear(x, params[f"{prefix}density_layer.weight"], params[f"{prefix}density_layer.bias"])
    if view_enc is not None:
        bott = F.linear(x, params[f"{prefix}extra_layer.weight"], params[f"{prefix}extra_layer.bias"])
        v = view_enc[:, None, :].expand(-1, x.shape[1], -1)
        x = torch.cat([bott, v], dim=-1)
        for i in range(net_depth_condition):
            x = F.relu(F.linear(x, params[f"{prefix}view_layers.{i}.0.weight"],
                                params[f"{prefix}view_layers.{i}.0.bias"]))
    raw_rgb = F.linear(x, params[f"{prefix}color_layer.weight"], params[f"{prefix}color_layer.bias"])
    return raw_rgb, raw_density


def forward(params: Dict[str, torch.Tensor], rays: Rays, randomized: bool, white_bkgd: bool,
            config: Optional[dict] = None, t_rand=None, u_jitter=None,
            return_debug=False, operand_dtype=None, grad=False, operand_split=False,
            density_normal=None) -> List[Tuple[torch.Tensor, ...]]:
    """MipNerf.forward (models/mip_nerf.py:172-248): list over levels of
    (comp_rgb [B,3], distance [B], acc [B], weights [B,N], t_samples [B,N+1]).
    `density_normal`: per level, the [B,N] standard normals standing in for torch.randn of models/mip_nerf.py:233
    (used when randomized and config['density_noise'] > 0; drawn with torch.randn when omitted, like the reference).
    `grad=True` keeps the autograd graph (training); the resampler then runs under no_grad on detached
    weights, which is what stop_resample_grad=True does (models/mip.py:250-264)."""
    cfg = dict(DEFAULT_CONFIG)
    cfg.update(config or {})
    ret, debug = [], []
    t, weights = None, None
    if grad:
        assert cfg["stop_resample_grad"], "oracle restates the default stop_resample_grad=True only"
    with (torch.enable_grad() if grad else torch.no_grad()):
        for level in range(cfg["num_levels"]):
            inds = None
            if level == 0:
                t, (means, covs) = sample_along_rays(
                    rays.origins, rays.directions, rays.radii, cfg["num_samples"], rays.near,
                    rays.far, randomized, cfg["disparity"], cfg["ray_shape"], t_rand=t_rand)
            else:
                with torch.no_grad():
                    t, (means, covs), inds = resample_along_rays(
                        rays.origins, rays.directions, rays.radii, t.detach(), weights.detach(), randomized,
                        cfg["ray_shape"], cfg["stop_resample_grad"], cfg["resample_padding"],
                        u_jitter=u_jitter, return_inds=True)
            if cfg["disable_integration"]:
                covs = torch.zeros_like(covs)
            enc = integrated_pos_enc(means, covs, cfg["min_deg_point"], cfg["max_deg_point"])
            view_enc = pos_enc(rays.viewdirs, 0, cfg["deg_view"], True) if cfg["use_viewdirs"] else None
            raw_rgb, raw_density = mlp_forward(params, enc, view_enc, cfg["mlp_net_depth"],
                                               cfg["mlp_skip_index"], cfg["mlp_net_depth_condition"],
                                               operand_dtype=operand_dtype, operand_split=operand_split)
            if randomized and cfg["density_noise"] > 0:  # models/mip_nerf.py:232-233
                z = (density_normal[level].reshape(raw_density.shape) if density_normal is not None
                     else torch.randn(raw_density.shape, dtype=raw_density.dtype))
                raw_density = raw_density + cfg["density_noise"] * z
            rgb = torch.sigmoid(raw_rgb) * (1 + 2 * cfg["rgb_padding"]) - cfg["rgb_padding"]
            density = F.softplus(raw_density + cfg["density_bias"])
            comp_rgb, distance, acc, weights = volumetric_rendering(rgb, density, t, rays.directions, white_bkgd)
            ret.append((comp_rgb, distance, acc, weights, t))
            debug.append(dict(means=means, covs=covs, enc=enc, raw_rgb=raw_rgb,
                              raw_density=raw_density, inds=inds))
    return (ret, debug) if return_debug else ret


def training_loss(params: Dict[str, torch.Tensor], rays: Rays, rgbs, randomized: bool, white_bkgd: bool,
                  coarse_loss_mult=0.1, disable_multiscale_loss=False, config=None, t_rand=None, u_jitter=None,
                  density_normal=None):
    """MipNeRFSystem.training_step's loss (models/nerf_system.py:95-111) with the autograd graph over
    `params` kept: (loss, [mse per level], [distloss per level], ret)."""
    ret = forward(params, rays, randomized, white_bkgd, config, t_rand=t_rand, u_jitter=u_jitter, grad=True,
                  density_normal=density_normal)
    mask = torch.ones_like(rays.lossmult) if disable_multiscale_loss else rays.lossmult
    losses, dls = [], []
    for (rgb, _, _, weights, t_samples) in ret:
        losses.append((mask * (rgb - rgbs[..., :3]) ** 2).sum() / mask.sum())
        dls.append(distloss(weights, t_samples))
    loss = coarse_loss_mult * (sum(losses[:-1]) + 0.01 * sum(dls[:-1])) + losses[-1] + 0.01 * dls[-1]
    return loss, losses, dls, ret


def gaussian_window(window_size=11, sigma=1.5):
    """utils/metrics.py:10-17."""
    g = torch.stack([torch.exp(torch.tensor(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)))
                     for x in range(window_size)])
    return g / g.sum()


def eval_errors(pred_color, batch_pixels, window_size=11, max_val=1.0):
    """(psnr, ssim) of [1,H,W,3] images: utils/metrics.py:190-197 (calc_psnr :182-188; SSIM :44-126 with the Gaussian
    window, zero padding and reduction='mean')."""
    psnr = -10.0 * torch.log10(torch.mean((pred_color - batch_pixels) ** 2))
    a, b = pred_color.permute(0, 3, 1, 2), batch_pixels.permute(0, 3, 1, 2)
    c = a.shape[1]
    g = gaussian_window(window_size, 1.5)
    kernel = torch.matmul(g.unsqueeze(-1), g.unsqueeze(-1).t()).repeat(c, 1, 1, 1)
    pad = (window_size - 1) // 2
    f = lambda x: F.conv2d(x, kernel, padding=pad, groups=c)  # noqa: E731
    mu1, mu2 = f(a), f(b)
    s1, s2, s12 = f(a * a) - mu1.pow(2), f(b * b) - mu2.pow(2), f(a * b) - mu1 * mu2
    c1, c2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    ssim_map = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1.pow(2) + mu2.pow(2) + c1) * (s1 + s2 + c2))
    return psnr, ssim_map.mean()


def mip_lr(step, lr_init, lr_final, max_steps, lr_delay_steps=0, lr_delay_mult=1.0):
    """MipLRDecay.get_lr (utils/lr_schedule.py:51-60) at `last_epoch == step`."""
    import numpy as np
    if lr_delay_steps > 0:
        delay_rate = lr_delay_mult + (1 - lr_delay_mult) * np.sin(0.5 * np.pi * np.clip(step / lr_delay_steps, 0, 1))
    else:
        delay_rate = 1.0
    tt = np.clip(step / max_steps, 0, 1)
    return float(delay_rate * np.exp(np.log(lr_init) * (1 - tt) + np.log(lr_final) * tt))


def render_image(params, rays: Rays, height: int, width: int, chunk_size: int, randomized=False,
                 white_bkgd=True, config=None):
    """Restatement of MipNeRFSystem.render_image's compute
    (models/nerf_system.py:151-177 + models/mip.py:404-421), without the logger."""
    flat = [getattr(rays, k).reshape(-1, getattr(rays, k).shape[-1]) for k in Rays._fields]
    n = flat[0].shape[0]
    coarse, fine, dist = [], [], []
    for s in range(0, n, chunk_size):
        chunk = Rays(*[f[s:s + chunk_size] for f in flat])
        (c_rgb, _, _, _, _), (f_rgb, d, _, _, _) = forward(params, chunk, randomized, white_bkgd, config)
        coarse.append(c_rgb), fine.append(f_rgb), dist.append(d)
    coarse = torch.cat(coarse).reshape(1, height, width, 3)
    fine = torch.cat(fine).reshape(1, height, width, 3)
    return coarse, fine, torch.cat(dist).reshape(1, height, width), rays.lossmult
