"""`MLP` and `MipNerf` with the reference's constructor signatures, submodule
tree and state_dict keys (models/mip_nerf.py:14-248), whose `forward` runs on
the sm_100a kernels of libmipnerf_b200.so through the C ABI.

The modules own ordinary fp32 `torch.nn.Linear` parameters, so Lightning
checkpoints of the reference (`mip_nerf.mlp.layers.{i}.0.weight`, ...) load with
`load_state_dict` unchanged.  `forward` builds no autograd graph; training goes
through `mipnerf_pl_b200.train` (`fused_loss` / `forward_backward`), where one library call runs forward and
backward and hands the gradients to autograd or `param.grad`.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import torch

from . import _cabi
from .ops import _dev, _f32, _ptr, _stream, draw_density_normal, draw_t_rand, draw_u_jitter
from .rays import Rays


def _xavier_init(linear):
    torch.nn.init.xavier_uniform_(linear.weight.data)


class _Workspace:
    """Per-(device, stream) scratch handed to the library (torch owns every byte).  Keyed by the stream
    the call is enqueued on, so forwards issued on different streams never share a scratch buffer."""
    _bufs = {}

    @classmethod
    def get(cls, device: torch.device, nbytes: int) -> torch.Tensor:
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device(),
               torch.cuda.current_stream(device).cuda_stream)
        buf = cls._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = None
            cls._bufs.pop(key, None)
            buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
            cls._bufs[key] = buf
        return buf


class MLP(torch.nn.Module):
    """models/mip_nerf.py:14-111 — same constructor, same parameter names."""

    def __init__(self, net_depth: int, net_width: int, net_depth_condition: int, net_width_condition: int,
                 skip_index: int, num_rgb_channels: int, num_density_channels: int, activation: str,
                 xyz_dim: int, view_dim: int):
        super().__init__()
        if activation != "relu":
            raise NotImplementedError  # models/mip_nerf.py:49-50
        self.net_depth, self.net_width = net_depth, net_width
        self.net_depth_condition, self.net_width_condition = net_depth_condition, net_width_condition
        self.skip_index = skip_index
        self.num_rgb_channels, self.num_density_channels = num_rgb_channels, num_density_channels
        self.xyz_dim, self.view_dim = xyz_dim, view_dim
        layers = []
        for i in range(net_depth):
            if i == 0:
                dim_in = xyz_dim
            elif (i - 1) % skip_index == 0 and i > 1:
                dim_in = net_width + xyz_dim
            else:
                dim_in = net_width
            linear = torch.nn.Linear(dim_in, net_width)
            _xavier_init(linear)
            layers.append(torch.nn.Sequential(linear, torch.nn.ReLU(True)))
        self.layers = torch.nn.ModuleList(layers)
        self.density_layer = torch.nn.Linear(net_width, num_density_channels)
        _xavier_init(self.density_layer)
        self.extra_layer = torch.nn.Linear(net_width, net_width)
        _xavier_init(self.extra_layer)
        layers = []
        for i in range(net_depth_condition):
            dim_in = net_width + view_dim if i == 0 else net_width_condition
            linear = torch.nn.Linear(dim_in, net_width_condition)
            _xavier_init(linear)
            layers.append(torch.nn.Sequential(linear, torch.nn.ReLU(True)))
        self.view_layers = torch.nn.Sequential(*layers)
        self.color_layer = torch.nn.Linear(net_width_condition, num_rgb_channels)
        self._packed = {}  # precision -> (versions, tensor)

    def __getstate__(self):  # ctypes marshalling caches are per-process
        d = self.__dict__.copy()
        d.pop("_ws_cache", None)
        d.pop("_lin_cache", None)
        d["_packed"] = {}
        return d

    # ---- marshalling --------------------------------------------------------------------------
    def linears(self) -> List[torch.nn.Linear]:
        """state_dict order expected by mipnerf_b200_weights (cached: the submodule tree is fixed after __init__)."""
        lins = self.__dict__.get("_lin_cache")
        if lins is None:
            lins = ([seq[0] for seq in self.layers] + [self.density_layer, self.extra_layer] +
                    [seq[0] for seq in self.view_layers] + [self.color_layer])
            self.__dict__["_lin_cache"] = lins
        return lins

    def _weights_struct(self, cfg: "_cabi.Config", precision: int, device):
        """(struct, keep-alive list), cached until a parameter is modified / moved / re-typed."""
        lins = self.linears()
        key = (precision, str(device), tuple((l.weight.data_ptr(), l.weight._version, l.bias.data_ptr(), l.bias._version,
                                               l.weight.dtype) for l in lins))
        hit = self.__dict__.get("_ws_cache")
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        ws, keep = self._weights_struct_uncached(cfg, precision, device, lins)
        self.__dict__["_ws_cache"] = (key, ws, keep)
        return ws, keep

    def _weights_struct_uncached(self, cfg, precision, device, lins):
        arr = (_cabi.Linear * len(lins))()
        keep = []
        for i, l in enumerate(lins):
            w, b = _f32(l.weight), _f32(l.bias)
            if w.device != device:
                raise RuntimeError(f"MLP parameters live on {w.device}, rays on {device}")
            keep += [w, b]
            arr[i] = _cabi.Linear(w.data_ptr(), b.data_ptr(), l.in_features, l.out_features)
        ws = _cabi.Weights(arr, len(lins), -1, None, 0)
        if precision != _cabi.FP32:
            packed = self._packed_image(cfg, ws, precision, device, lins)
            ws.packed, ws.packed_bytes, ws.packed_precision = packed.data_ptr(), packed.numel(), precision
            keep.append(packed)
        keep.append(arr)
        return ws, keep

    def _packed_image(self, cfg, ws, precision, device, lins):
        versions = tuple((p.data_ptr(), p._version) for l in lins for p in (l.weight, l.bias))
        hit = self._packed.get((precision, str(device)))
        if hit is not None and hit[0] == versions:
            return hit[1]
        lib = _cabi.lib()
        nbytes = lib.mipnerf_b200_packed_weights_bytes(C.byref(cfg), precision)
        if nbytes == 0:
            raise NotImplementedError("tensor-core path: the 8x256 / 1x128 MLP with num_samples=128, min_deg_point=0, "
                                      "max_deg_point 1..16 and deg_view 1..4 is implemented; use precision='fp32'")
        packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            _cabi.check(lib.mipnerf_b200_pack_weights(C.byref(cfg), C.byref(ws), precision, packed.data_ptr(),
                                                      nbytes, _stream(device)), "pack_weights")
        self._packed[(precision, str(device))] = (versions, packed)
        return packed

    def _config(self, num_samples=128, **over) -> "_cabi.Config":
        deg_pts = self.xyz_dim // 6
        deg_view = (self.view_dim - 3) // 6
        vals = dict(num_samples=num_samples, num_levels=1, min_deg_point=0, max_deg_point=deg_pts,
                    deg_view=deg_view, use_viewdirs=1, disparity=0, disable_integration=0,
                    resample_padding=0.01, density_bias=-1.0, rgb_padding=0.001, net_depth=self.net_depth,
                    net_width=self.net_width, net_depth_condition=self.net_depth_condition,
                    net_width_condition=self.net_width_condition, skip_index=self.skip_index,
                    num_rgb_channels=self.num_rgb_channels, num_density_channels=self.num_density_channels)
        vals.update(over)
        return _cabi.Config(**vals)

    def forward(self, x, view_direction=None, precision: str = "fp32"):
        """models/mip_nerf.py:75-111: x [B,N,xyz_dim], view_direction [B,view_dim] ->
        (raw_rgb [B,N,3], raw_density [B,N,1])."""
        dev = _dev(x)
        xx = _f32(x)
        b, n = xx.shape[0], xx.shape[1]
        vd = _f32(view_direction) if view_direction is not None else None
        prec = _cabi.PRECISIONS[precision]
        cfg = self._config(use_viewdirs=int(vd is not None))
        ws, keep = self._weights_struct(cfg, prec, dev)
        lib = _cabi.lib()
        raw_rgb = torch.empty(b, n, 3, device=dev)
        raw_density = torch.empty(b, n, 1, device=dev)
        nbytes = lib.mipnerf_b200_mlp_workspace_bytes(C.byref(cfg), b, n, prec)
        scratch = _Workspace.get(dev, nbytes)
        with torch.cuda.device(dev):
            _cabi.check(lib.mipnerf_b200_mlp_forward(
                C.byref(cfg), C.byref(ws), xx.data_ptr(), _ptr(vd), b, n, prec, raw_rgb.data_ptr(),
                raw_density.data_ptr(), scratch.data_ptr(), scratch.numel(), _stream(dev)), "MLP.forward")
        return raw_rgb, raw_density


class LevelOutputs(list):
    """What `MipNerf.forward` returns: the reference's list of per-level 5-tuples (models/mip_nerf.py:246), plus
    `.pixels` — comp_rgb | distance | acc of every level as one contiguous [levels, 5*B] tensor (a view of the same
    memory), for reading rendered pixels back to the host with a single copy."""
    pixels: Optional[torch.Tensor] = None


class MipNerf(torch.nn.Module):
    """models/mip_nerf.py:114-248 — same constructor (plus `precision`), same forward contract."""

    def __init__(self, num_samples: int = 128, num_levels: int = 2, resample_padding: float = 0.01,
                 stop_resample_grad: bool = True, use_viewdirs: bool = True, disparity: bool = False,
                 ray_shape: str = 'cone', min_deg_point: int = 0, max_deg_point: int = 16, deg_view: int = 4,
                 density_activation: str = 'softplus', density_noise: float = 0., density_bias: float = -1.,
                 rgb_activation: str = 'sigmoid', rgb_padding: float = 0.001,
                 disable_integration: bool = False, append_identity: bool = True, mlp_net_depth: int = 8,
                 mlp_net_width: int = 256, mlp_net_depth_condition: int = 1, mlp_net_width_condition: int = 128,
                 mlp_skip_index: int = 4, mlp_num_rgb_channels: int = 3, mlp_num_density_channels: int = 1,
                 mlp_net_activation: str = 'relu', precision: Optional[str] = None):
        super().__init__()
        self.num_levels = num_levels
        self.num_samples = num_samples
        self.disparity = disparity
        self.ray_shape = ray_shape
        self.disable_integration = disable_integration
        self.min_deg_point = min_deg_point
        self.max_deg_point = max_deg_point
        self.use_viewdirs = use_viewdirs
        self.deg_view = deg_view
        self.density_noise = density_noise
        self.density_bias = density_bias
        self.resample_padding = resample_padding
        self.stop_resample_grad = stop_resample_grad
        mlp_xyz_dim = (max_deg_point - min_deg_point) * 3 * 2
        mlp_view_dim = deg_view * 3 * 2
        mlp_view_dim = mlp_view_dim + 3 if append_identity else mlp_view_dim
        self.mlp = MLP(mlp_net_depth, mlp_net_width, mlp_net_depth_condition, mlp_net_width_condition,
                       mlp_skip_index, mlp_num_rgb_channels, mlp_num_density_channels, mlp_net_activation,
                       mlp_xyz_dim, mlp_view_dim)
        if rgb_activation != 'sigmoid':
            raise NotImplementedError  # models/mip_nerf.py:162-165
        self.rgb_padding = rgb_padding
        if density_activation != 'softplus':
            raise NotImplementedError  # models/mip_nerf.py:167-170
        # forward() always encodes viewdirs with append_identity=True (models/mip_nerf.py:221-226);
        # a model built with append_identity=False has a 24-wide view input and fails in the reference.
        self._append_identity = bool(append_identity)
        # 'fp32' | 'bf16' | 'fp16' | 'fp16x3' | 'bf16x3'; None -> $MIPNERF_B200_PRECISION or 'fp32'
        self.precision = precision or os.environ.get("MIPNERF_B200_PRECISION", "fp32")
        # randomized=True without injected noise draws its uniforms inside the kernels (Philox4x32-10, counter-based):
        # seed from torch's global generator at first use, offset advanced by one per call.
        self.rng_seed: Optional[int] = None
        self.rng_offset = 0

    def next_rng(self) -> "_cabi.Rng":
        """(seed, offset) of the next randomized call; advances the offset (the role of torch's generator offset)."""
        if self.rng_seed is None:
            self.rng_seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        rng = _cabi.Rng(self.rng_seed, self.rng_offset)
        self.rng_offset += 1
        return rng

    def _config(self) -> "_cabi.Config":
        return self.mlp._config(
            num_samples=self.num_samples, num_levels=self.num_levels, min_deg_point=self.min_deg_point,
            max_deg_point=self.max_deg_point, deg_view=self.deg_view, use_viewdirs=int(bool(self.use_viewdirs)),
            disparity=int(bool(self.disparity)), disable_integration=int(bool(self.disable_integration)),
            resample_padding=float(self.resample_padding), density_bias=float(self.density_bias),
            rgb_padding=float(self.rgb_padding), density_noise=float(self.density_noise))

    def forward(self, rays: Rays, randomized: bool, white_bkgd: bool, *, t_rand: Optional[torch.Tensor] = None,
                u_jitter: Optional[torch.Tensor] = None, density_normal=None, return_inds: bool = False):
        """rays -> [(comp_rgb [B,3], distance [B], acc [B], weights [B,N], t_samples [B,N+1])] * levels
        (models/mip_nerf.py:172-248).  `t_rand` / `u_jitter` / `density_normal` (one [B,N] tensor of standard normals
        per level, models/mip_nerf.py:233) inject the noise of randomized mode; without any of them the kernels draw
        in-kernel.  With `return_inds` a sixth element (searchsorted indices, None for level 0) is appended."""
        if self.ray_shape == 'cylinder':
            raise NotImplementedError  # models/mip.py:97-98
        assert self.ray_shape == 'cone'
        if self.use_viewdirs and not self._append_identity:
            raise RuntimeError("append_identity=False: view encoding width does not match view_layers "
                               "(same failure as the reference)")
        dev = _dev(rays.origins)
        b = rays.origins.shape[0]
        n = self.num_samples
        prec = _cabi.PRECISIONS[self.precision]
        cfg = self._config()
        keep = [_f32(rays.origins), _f32(rays.directions), _f32(rays.viewdirs), _f32(rays.radii).reshape(-1),
                _f32(rays.near).reshape(-1), _f32(rays.far).reshape(-1)]
        rs = _cabi.RaysStruct(keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr(), keep[3].data_ptr(),
                              keep[4].data_ptr(), keep[5].data_ptr(), b)
        rng = None
        noisy = bool(randomized) and self.density_noise > 0     # models/mip_nerf.py:232
        if randomized and t_rand is None and u_jitter is None and density_normal is None:
            rng = self.next_rng()                       # in-kernel Philox: no torch.rand launch, no [B,N+1] arrays
            normals = [None] * self.num_levels
        elif randomized:
            t_rand = _f32(t_rand) if t_rand is not None else draw_t_rand(b, n, dev)
            u_jitter = _f32(u_jitter) if u_jitter is not None else draw_u_jitter(b, n + 1, dev)
            normals = list(density_normal) if density_normal is not None else [None] * self.num_levels
            if len(normals) != self.num_levels:
                raise ValueError(f"density_normal: expected {self.num_levels} tensors (one per level)")
            normals = [(_f32(x).reshape(b, n) if x is not None else draw_density_normal(b, n, dev)) if noisy else None
                       for x in normals]
        else:
            t_rand = u_jitter = None
            normals = [None] * self.num_levels
        ws, wkeep = self.mlp._weights_struct(cfg, prec, dev)
        outs = (_cabi.LevelOut * self.num_levels)()
        # One allocation for everything: the per-level pixel outputs (comp_rgb | distance | acc = 5 floats/ray) of
        # all levels first, so that a caller that only wants pixels reads them back with ONE contiguous copy
        # (`ret.pixels`, [levels, 5*B]); then weights / fenceposts per level.
        levels = self.num_levels
        flat = torch.empty(levels * b * (5 + n + n + 1), device=dev)
        ret = LevelOutputs()
        ret.pixels = flat[:levels * 5 * b].view(levels, 5 * b) if b > 0 else flat[:0].view(levels, 0)
        tail = levels * 5 * b
        for lvl in range(levels):
            o = lvl * 5 * b
            comp = flat[o:o + 3 * b].view(b, 3)
            dist = flat[o + 3 * b:o + 4 * b]
            acc = flat[o + 4 * b:o + 5 * b]
            q = tail + lvl * (2 * n + 1) * b
            w = flat[q:q + n * b].view(b, n)
            t = flat[q + n * b:q + (2 * n + 1) * b].view(b, n + 1)
            inds = torch.empty(b, n + 1, device=dev, dtype=torch.int64) if (return_inds and lvl > 0) else None
            outs[lvl] = _cabi.LevelOut(comp.data_ptr(), dist.data_ptr(), acc.data_ptr(), w.data_ptr(),
                                       t.data_ptr(), _ptr(inds), _ptr(normals[lvl]))
            ret.append((comp, dist, acc, w, t, inds) if return_inds else (comp, dist, acc, w, t))
        lib = _cabi.lib()
        nbytes = lib.mipnerf_b200_workspace_bytes(C.byref(cfg), b, prec)
        if nbytes == 0 and b > 0:
            _cabi.check(lib.mipnerf_b200_forward(C.byref(cfg), C.byref(ws), C.byref(rs), 0, None, None, 0, prec,
                                                 outs, None, 0, None), "MipNerf.forward")
        scratch = _Workspace.get(dev, nbytes)
        if rng is not None:
            fn = lib.mipnerf_b200_forward_rng
            args = (C.byref(cfg), C.byref(ws), C.byref(rs), C.byref(rng), int(bool(white_bkgd)), prec, outs,
                    scratch.data_ptr(), scratch.numel(), _stream(dev))
        else:
            fn = lib.mipnerf_b200_forward
            args = (C.byref(cfg), C.byref(ws), C.byref(rs), int(bool(randomized)), _ptr(t_rand), _ptr(u_jitter),
                    int(bool(white_bkgd)), prec, outs, scratch.data_ptr(), scratch.numel(), _stream(dev))
        if dev.index is None or dev.index == torch.cuda.current_device():
            _cabi.check(fn(*args), "MipNerf.forward")
        else:
            with torch.cuda.device(dev):
                _cabi.check(fn(*args), "MipNerf.forward")
        return ret
