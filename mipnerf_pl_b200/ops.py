"""Host-side mirror of the reference's ray-math free functions (models/mip.py).

Same names, argument meaning and error behaviour as the reference; each call
marshals CUDA tensors into the matching C-ABI entry point of
libmipnerf_b200.so.  CPU tensors are rejected: there is no fallback path.

Differences that the C ABI forces and that are visible here:
  * random draws are explicit optional arguments (`t_rand`, `u_jitter`, `density_normal`) so a
    caller (or a parity test) can inject the noise; when omitted they are
    drawn with torch's CUDA generator;
  * everything is fp32; other dtypes are cast at the boundary.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _cabi

F32_EPS = float(torch.finfo(torch.float32).eps)


def _dev(t: torch.Tensor) -> torch.device:
    if not isinstance(t, torch.Tensor) or t.device.type != "cuda":
        raise RuntimeError("mipnerf_pl_b200 runs on CUDA tensors only (B200 path, no CPU fallback); "
                           f"got {getattr(t, 'device', type(t))}")
    return t.device


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(dtype=torch.float32).contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _rays_struct(origins, directions, radii, near=None, far=None, viewdirs=None):
    n = origins.shape[0]
    keep = [_f32(origins), _f32(directions), _f32(radii).reshape(-1)]
    near_t = _f32(near).reshape(-1) if near is not None else torch.zeros(n, device=origins.device)
    far_t = _f32(far).reshape(-1) if far is not None else torch.zeros(n, device=origins.device)
    vd = _f32(viewdirs) if viewdirs is not None else None
    keep += [near_t, far_t, vd]
    s = _cabi.RaysStruct(keep[0].data_ptr(), keep[1].data_ptr(), _ptr(vd), keep[2].data_ptr(),
                         near_t.data_ptr(), far_t.data_ptr(), n)
    return s, keep


def draw_t_rand(batch: int, num_samples: int, device) -> torch.Tensor:
    """torch.rand(batch, N+1) of models/mip.py:159."""
    return torch.rand(batch, num_samples + 1, device=device, dtype=torch.float32)


def draw_u_jitter(batch: int, num_draws: int, device) -> torch.Tensor:
    """uniform_(to=1/num_draws - eps) of models/mip.py:201-202."""
    return torch.empty(batch, num_draws, device=device, dtype=torch.float32).uniform_(
        0.0, 1.0 / num_draws - F32_EPS)


def philox_uniform(seed: int, offset: int, stream_id: int, batch: int, num_draws: int, device) -> torch.Tensor:
    """The uniforms the kernels draw in-kernel for (seed, offset): stream 0 = t_rand in [0,1) (models/mip.py:159),
    stream 1 + level = that level's u_jitter in [0, 1/num_draws - eps) (models/mip.py:201-202).  Passing them as
    `t_rand` / `u_jitter` reproduces the in-kernel randomized forward bit for bit."""
    dev = torch.device(device)
    out = torch.empty(batch, num_draws, device=dev)
    rng = _cabi.Rng(seed & 0xFFFFFFFFFFFFFFFF, offset)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().mipnerf_b200_philox_uniform(C.byref(rng), stream_id, batch, num_draws, out.data_ptr(),
                                                            _stream(dev)), "philox_uniform")
    return out


def draw_density_normal(batch: int, num_samples: int, device) -> torch.Tensor:
    """torch.randn(raw_density.shape) of models/mip_nerf.py:233 (one [B,N] draw per level)."""
    return torch.randn(batch, num_samples, device=device, dtype=torch.float32)


def philox_normal(seed: int, offset: int, level: int, batch: int, num_samples: int, device) -> torch.Tensor:
    """The standard normals the kernels draw in-kernel for (seed, offset) as the density noise of `level`
    (models/mip_nerf.py:232-233).  Passed as `density_normal[level]` they reproduce the in-kernel randomized
    forward bit for bit."""
    dev = torch.device(device)
    out = torch.empty(batch, num_samples, device=dev)
    rng = _cabi.Rng(seed & 0xFFFFFFFFFFFFFFFF, offset)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().mipnerf_b200_philox_normal(C.byref(rng), level, batch, num_samples, out.data_ptr(),
                                                           _stream(dev)), "philox_normal")
    return out


def cast_rays(t_samples, origins, directions, radii, ray_shape, diagonal=True):
    """models/mip.py:81-103."""
    if ray_shape == "cylinder":
        raise NotImplementedError
    assert ray_shape == "cone"
    if not diagonal:
        raise NotImplementedError("full-covariance branch is dead code in the reference (SURVEY §2)")
    dev = _dev(t_samples)
    t = _f32(t_samples)
    b, n = t.shape[0], t.shape[1] - 1
    rs, keep = _rays_struct(origins, directions, radii)
    means = torch.empty(b, n, 3, device=dev)
    covs = torch.empty(b, n, 3, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().mipnerf_b200_cast_rays(C.byref(rs), t.data_ptr(), n, means.data_ptr(),
                                                       covs.data_ptr(), _stream(dev)), "cast_rays")
    return means, covs


def sample_along_rays(origins, directions, radii, num_samples, near, far, randomized, disparity, ray_shape,
                      t_rand: Optional[torch.Tensor] = None):
    """models/mip.py:127-165 -> (t_samples [B,N+1], (means, covs))."""
    if ray_shape == "cylinder":
        raise NotImplementedError
    assert ray_shape == "cone"
    dev = _dev(origins)
    b = origins.shape[0]
    rs, keep = _rays_struct(origins, directions, radii, near, far)
    if randomized and t_rand is None:
        t_rand = draw_t_rand(b, num_samples, dev)
    tr = _f32(t_rand) if randomized else None
    t = torch.empty(b, num_samples + 1, device=dev)
    means = torch.empty(b, num_samples, 3, device=dev)
    covs = torch.empty(b, num_samples, 3, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().mipnerf_b200_sample_along_rays(
            C.byref(rs), num_samples, int(bool(randomized)), int(bool(disparity)), _ptr(tr), t.data_ptr(),
            means.data_ptr(), covs.data_ptr(), _stream(dev)), "sample_along_rays")
    return t, (means, covs)


def sorted_piecewise_constant_pdf(bins, weights, num_samples, randomized,
                                  u_jitter: Optional[torch.Tensor] = None, return_inds: bool = False):
    """models/mip.py:168-229.  `weights` is NOT modified (the reference pads it in place)."""
    dev = _dev(bins)
    bn, w = _f32(bins), _f32(weights)
    b, nb = w.shape
    if randomized and u_jitter is None:
        u_jitter = draw_u_jitter(b, num_samples, dev)
    uj = _f32(u_jitter) if randomized else None
    out = torch.empty(b, num_samples, device=dev)
    inds = torch.empty(b, num_samples, device=dev, dtype=torch.int64) if return_inds else None
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().mipnerf_b200_sorted_piecewise_constant_pdf(
            bn.data_ptr(), w.data_ptr(), b, nb, num_samples, int(bool(randomized)), _ptr(uj), out.data_ptr(),
            _ptr(inds), _stream(dev)), "sorted_piecewise_constant_pdf")
    return (out, inds) if return_inds else out


def resample_along_rays(origins, directions, radii, t_samples, weights, randomized, ray_shape, stop_grad,
                        resample_padding, u_jitter: Optional[torch.Tensor] = None, return_inds: bool = False):
    """models/mip.py:232-280 -> (new_t [B,N+1], (means, covs)).  Forward only, so `stop_grad`
    (which only changes autograd in the reference) has no effect on the values."""
    if ray_shape == "cylinder":
        raise NotImplementedError
    assert ray_shape == "cone"
    dev = _dev(t_samples)
    t, w = _f32(t_samples), _f32(weights)
    b, n = w.shape
    rs, keep = _rays_struct(origins, directions, radii)
    if randomized and u_jitter is None:
        u_jitter = draw_u_jitter(b, n + 1, dev)
    uj = _f32(u_jitter) if randomized else None
    new_t = torch.empty(b, n + 1, device=dev)
    means = torch.empty(b, n, 3, device=dev)
    covs = torch.empty(b, n, 3, device=dev)
    inds = torch.empty(b, n + 1, device=dev, dtype=torch.int64) if return_inds else None
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().mipnerf_b200_resample_along_rays(
            C.byref(rs), t.data_ptr(), w.data_ptr(), n, int(bool(randomized)), _ptr(uj),
            float(resample_padding), new_t.data_ptr(), means.data_ptr(), covs.data_ptr(), _ptr(inds),
            _stream(dev)), "resample_along_rays")
    return (new_t, (means, covs), inds) if return_inds else (new_t, (means, covs))


def integrated_pos_enc(means_covs, min_deg, max_deg, diagonal=True):
    """models/mip.py:322-350 (diagonal): ([..,3],[..,3]) -> [.., 6*(max-min)]."""
    if not diagonal:
        raise NotImplementedError("full-covariance branch is dead code in the reference (SURVEY §2)")
    means, covs = means_covs
    dev = _dev(means)
    m, c = _f32(means), _f32(covs)
    lead = m.shape[:-1]
    npts = m.numel() // 3
    out = torch.empty(*lead, 6 * (max_deg - min_deg), device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().mipnerf_b200_integrated_pos_enc(
            m.data_ptr(), c.data_ptr(), npts, int(min_deg), int(max_deg), out.data_ptr(), _stream(dev)),
            "integrated_pos_enc")
    return out


def pos_enc(x, min_deg, max_deg, append_identity=True):
    """models/mip.py:353-363."""
    dev = _dev(x)
    xx = _f32(x)
    lead = xx.shape[:-1]
    width = 6 * (max_deg - min_deg) + (3 if append_identity else 0)
    out = torch.empty(*lead, width, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().mipnerf_b200_pos_enc(
            xx.data_ptr(), xx.numel() // 3, int(min_deg), int(max_deg), int(bool(append_identity)),
            out.data_ptr(), _stream(dev)), "pos_enc")
    return out


def volumetric_rendering(rgb, density, t_samples, dirs, white_bkgd):
    """models/mip.py:366-401 -> (comp_rgb [B,3], distance [B], acc [B], weights [B,N])."""
    dev = _dev(rgb)
    r, d, t, dd = _f32(rgb), _f32(density), _f32(t_samples), _f32(dirs)
    b, n = r.shape[0], r.shape[1]
    comp = torch.empty(b, 3, device=dev)
    dist = torch.empty(b, device=dev)
    acc = torch.empty(b, device=dev)
    w = torch.empty(b, n, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().mipnerf_b200_volumetric_rendering(
            r.data_ptr(), d.data_ptr(), t.data_ptr(), dd.data_ptr(), b, n, int(bool(white_bkgd)),
            comp.data_ptr(), dist.data_ptr(), acc.data_ptr(), w.data_ptr(), _stream(dev)),
            "volumetric_rendering")
    return comp, dist, acc, w


def distloss(weight, samples):
    """Distortion loss value (models/mip.py:8-20): weight [B,N], samples [B,N+1] -> scalar.
    Value only (its gradient is part of `train.forward_backward`); O(N) per ray instead of the
    reference's two [B,N,N] temporaries."""
    dev = _dev(weight)
    w, t = _f32(weight), _f32(samples)
    b, n = w.shape
    out = torch.empty(b, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().mipnerf_b200_distloss(w.data_ptr(), t.data_ptr(), b, n, out.data_ptr(), _stream(dev)),
                    "distloss")
    return out.mean()
