"""Ray container and host-side ray helpers.

`Rays` keeps the reference's field order (datasets/datasets.py:13-16);
`rearrange_render_image` mirrors models/mip.py:404-421.
`blender_rays` / `spheric_pose` synthesise Blender-shape rays the way the
reference's loaders do (datasets/datasets.py:214-263, utils/vis.py:159-198,
render_video.py:29-105) because no dataset is reachable offline.
"""
from __future__ import annotations

import collections
from typing import List, Tuple

import numpy as np
import torch

Rays = collections.namedtuple(
    "Rays", ("origins", "directions", "viewdirs", "radii", "lossmult", "near", "far"))
Rays_keys = Rays._fields


def namedtuple_map(fn, tup):
    """Apply `fn` to every field (datasets/datasets.py:19-21)."""
    return type(tup)(*map(fn, tup))


_FIELD_WIDTHS = (3, 3, 3, 1, 1, 1, 1)  # floats per ray of each Rays field: 13 in total (52 B/ray)


class RayStaging:
    """A ray batch in ONE pinned, field-major host buffer ([origins | directions | viewdirs | radii | lossmult |
    near | far], 13*B floats) so that a step moves it to the GPU with a single host-to-device copy instead of
    seven, and hands the kernels the same seven contiguous row-major tensors as views of one device buffer.
    The reference's DataLoader does the equivalent with `pin_memory=True` + per-field `.to(device)`."""

    def __init__(self, rays: Rays, pin: bool = True):
        b = rays.origins.shape[0]
        self.num_rays = b
        self.host = torch.empty(13 * b, dtype=torch.float32, pin_memory=pin and torch.cuda.is_available())
        self.fill(rays)
        self._dev = {}

    def _views(self, flat: torch.Tensor) -> Rays:
        b, off, out = self.num_rays, 0, []
        for w in _FIELD_WIDTHS:
            out.append(flat[off:off + w * b].view(b, w))
            off += w * b
        return Rays(*out)

    def fill(self, rays: Rays) -> None:
        """Overwrite the host buffer with another batch of the same size."""
        for dst, src in zip(self._views(self.host), rays):
            dst.copy_(src.reshape(dst.shape))

    @property
    def host_rays(self) -> Rays:
        return self._views(self.host)

    def to(self, device, non_blocking: bool = True) -> Rays:
        """One H2D copy on the current stream; returns Rays whose fields are views of the device buffer
        (reused between calls: consume the rays before the next `to`)."""
        device = torch.device(device)
        buf = self._dev.get(device)
        if buf is None:
            buf = self._dev[device] = torch.empty(13 * self.num_rays, dtype=torch.float32, device=device)
        buf.copy_(self.host, non_blocking=non_blocking)
        return self._views(buf)


def rearrange_render_image(rays: Rays, chunk_size: int = 4096) -> Tuple[List[Rays], torch.Tensor]:
    """[1,H,W,C] ray fields -> list of flat chunks + the lossmult mask
    (models/mip.py:404-421)."""
    val_mask = rays.lossmult
    flat = [getattr(rays, k).reshape(-1, getattr(rays, k).shape[-1]) for k in Rays_keys]
    n = flat[0].shape[0]
    chunks = [Rays(*[f[s:s + chunk_size] for f in flat]) for s in range(0, n, chunk_size)]
    return chunks, val_mask


# ---------------------------------------------------------------------------
# synthetic Blender-shape rays
# ---------------------------------------------------------------------------
BLENDER_CAMERA_ANGLE_X = 0.6911112070083618  # render_video.py:191-192


def spheric_pose(theta: float, phi: float = -np.pi / 5, radius: float = 4.0) -> np.ndarray:
    """One camera-to-world [3,4] on the reference's circular path (utils/vis.py:169-193)."""
    trans = np.eye(4)
    trans[2, 3] = radius
    cp, sp = np.cos(phi), np.sin(phi)
    rot_phi = np.array([[1, 0, 0, 0], [0, cp, -sp, 0], [0, sp, cp, 0], [0, 0, 0, 1.0]])
    ct, st = np.cos(theta), np.sin(theta)
    rot_theta = np.array([[ct, 0, -st, 0], [0, 1, 0, 0], [st, 0, ct, 0], [0, 0, 0, 1.0]])
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1.0]])
    return (flip @ rot_theta @ rot_phi @ trans)[:3].astype(np.float32)


def blender_rays(c2w: np.ndarray, height: int = 800, width: int = 800, scale: int = 1,
                 near: float = 2.0, far: float = 6.0,
                 camera_angle_x: float = BLENDER_CAMERA_ANGLE_X) -> Rays:
    """All H*W rays of one pose as numpy fp32 arrays shaped [H, W, C]
    (datasets/datasets.py:214-263).  `scale` = 2**j gives the j-th level of the
    multi-scale set: H, W and focal divided by it, lossmult = scale**2
    (datasets/convert_blender_data.py:66-81)."""
    h, w = height // scale, width // scale
    focal = np.float32(0.5 * width / np.tan(0.5 * camera_angle_x)) / np.float32(scale)
    x, y = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32), indexing="xy")
    cam = np.stack([(x - w * 0.5 + 0.5) / focal, -(y - h * 0.5 + 0.5) / focal, -np.ones_like(x)], -1)
    c2w = np.asarray(c2w, dtype=np.float32)
    directions = (cam @ c2w[:3, :3].T).astype(np.float32)
    origins = np.broadcast_to(c2w[:3, -1], directions.shape).astype(np.float32)
    viewdirs = (directions / np.linalg.norm(directions, axis=-1, keepdims=True)).astype(np.float32)
    dx = np.sqrt(np.sum((directions[:-1] - directions[1:]) ** 2, -1))
    dx = np.concatenate([dx, dx[-2:-1]], 0)
    radii = (dx[..., None] * 2 / np.sqrt(12)).astype(np.float32)  # kernel ABI is fp32
    ones = np.ones_like(origins[..., :1])
    return Rays(origins, directions, viewdirs, radii, (ones * float(scale * scale)).astype(np.float32),
                (ones * near).astype(np.float32), (ones * far).astype(np.float32))


def rays_to_torch(rays: Rays, device="cpu", flatten: bool = True) -> Rays:
    def conv(a):
        t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
        if flatten:
            t = t.reshape(-1, t.shape[-1])
        return t.to(device=device, dtype=torch.float32).contiguous()
    return namedtuple_map(conv, rays)


def random_ray_batch(num_rays: int, seed: int = 0, multiscale: bool = False, device="cpu") -> Rays:
    """`num_rays` Blender-shape rays drawn without replacement from one pose on the
    radius-4 spheric path.  multiscale=True mixes the four
    resolutions 800/400/200/100 1:1:1:1 (BASELINE config 3)."""
    rng = np.random.RandomState(seed)
    c2w = spheric_pose(float(rng.uniform(0, 2 * np.pi)))
    scales = (1, 2, 4, 8) if multiscale else (1,)
    parts = []
    per = [num_rays // len(scales)] * len(scales)
    per[0] += num_rays - sum(per)
    for sc, n in zip(scales, per):
        full = blender_rays(c2w, scale=sc)
        flat = [f.reshape(-1, f.shape[-1]) for f in full]
        total = flat[0].shape[0]
        idx = rng.choice(total, size=n, replace=n > total)
        parts.append([f[idx] for f in flat])
    merged = [np.concatenate([p[i] for p in parts], 0) for i in range(len(Rays_keys))]
    return rays_to_torch(Rays(*merged), device=device, flatten=False)
