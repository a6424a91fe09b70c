"""Deterministic, machine-independent MLP parameters for tests and benchmarks.

No checkpoint is reachable offline, so parity and throughput are measured on
random-init weights of the reference architecture (models/mip_nerf.py:19-73):
xavier-uniform for every layer except `color_layer` (torch default init), drawn
from numpy's RandomState so every box generates bit-identical tensors.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch


def layer_shapes(net_depth=8, net_width=256, net_depth_condition=1, net_width_condition=128, skip_index=4,
                 xyz_dim=96, view_dim=27, num_rgb=3, num_density=1):
    """[(state_dict prefix, out_features, in_features)] in the reference's registration order."""
    shapes = []
    for i in range(net_depth):
        if i == 0:
            k = xyz_dim
        elif (i - 1) % skip_index == 0 and i > 1:
            k = net_width + xyz_dim
        else:
            k = net_width
        shapes.append((f"mlp.layers.{i}.0", net_width, k))
    shapes.append(("mlp.density_layer", num_density, net_width))
    shapes.append(("mlp.extra_layer", net_width, net_width))
    for i in range(net_depth_condition):
        k = net_width + view_dim if i == 0 else net_width_condition
        shapes.append((f"mlp.view_layers.{i}.0", net_width_condition, k))
    shapes.append(("mlp.color_layer", num_rgb, net_width_condition))
    return shapes


def make_state_dict(seed: int = 0, kind: str = "xavier", **shape_kwargs) -> "OrderedDict[str, torch.Tensor]":
    """kind='xavier': the reference's init.  kind='trained_like': same tensors with the density
    head scaled up and shifted so densities span empty -> opaque along a ray (spiky compositing
    weights, strongly non-uniform fine sampling) — the stress set of SURVEY.md §8d."""
    rng = np.random.RandomState(seed)
    sd = OrderedDict()
    for name, out_f, in_f in layer_shapes(**shape_kwargs):
        if name.endswith("color_layer"):
            bound_w = 1.0 / math.sqrt(in_f)          # kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in))
        else:
            bound_w = math.sqrt(6.0 / (in_f + out_f))  # xavier_uniform
        bound_b = 1.0 / math.sqrt(in_f)
        w = rng.uniform(-bound_w, bound_w, size=(out_f, in_f)).astype(np.float32)
        b = rng.uniform(-bound_b, bound_b, size=(out_f,)).astype(np.float32)
        if kind == "trained_like" and name.endswith("density_layer"):
            w = w * 40.0
            b = b - 2.0
        elif kind == "trained_like" and name.endswith("color_layer"):
            w = w * 8.0
        elif kind not in ("xavier", "trained_like"):
            raise ValueError(kind)
        sd[name + ".weight"] = torch.from_numpy(w)
        sd[name + ".bias"] = torch.from_numpy(b)
    return sd
