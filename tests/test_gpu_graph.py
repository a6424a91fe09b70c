"""CUDA-graph replay of the forward (`GraphedForward`): bit-identical to the eager call, re-readable inputs."""
import pytest
import torch

from helpers import make_state_dict

pytestmark = pytest.mark.gpu

import mipnerf_pl_b200 as mp  # noqa: E402

DEV = "cuda:0"


@pytest.mark.parametrize("precision", ["bf16", "fp16x3"])
def test_graph_replay_equals_eager_and_tracks_new_rays(precision):
    model = mp.MipNerf(precision=precision)
    model.load_state_dict(make_state_dict(seed=2, kind="trained_like"))
    model = model.to(DEV).eval()
    a, b = mp.random_ray_batch(700, seed=1), mp.random_ray_batch(700, seed=2, multiscale=True)
    staging = mp.RayStaging(a)
    gf = mp.GraphedForward(model, staging, white_bkgd=True, device=DEV)
    for rays in (a, b, a):
        got = gf(rays)                                      # fill pinned buffer, ONE H2D copy, ONE graph launch
        torch.cuda.synchronize()
        want = model(mp.namedtuple_map(lambda t: t.to(DEV), rays), False, True)
        for lvl in range(2):
            for k in range(5):
                assert torch.equal(got[lvl][k], want[lvl][k]), (precision, lvl, k)
        assert torch.equal(got.pixels, torch.cat([torch.cat([want[l][0].reshape(-1), want[l][1], want[l][2]])[None]
                                                  for l in range(2)]))


def test_graph_sees_updated_weights_only_after_recapture():
    """The packed operand image is baked into the captured launches: a weight update needs a new GraphedForward
    (documented behaviour; inference weights are static)."""
    model = mp.MipNerf(precision="bf16")
    model.load_state_dict(make_state_dict(seed=0))
    model = model.to(DEV).eval()
    rays = mp.random_ray_batch(64, seed=0)
    gf = mp.GraphedForward(model, mp.RayStaging(rays), device=DEV)
    first = gf().pixels.clone()
    model.load_state_dict(make_state_dict(seed=1))
    gf2 = mp.GraphedForward(model, mp.RayStaging(rays), device=DEV)
    assert not torch.equal(gf2().pixels, first)
    want = model(mp.namedtuple_map(lambda t: t.to(DEV), rays), False, True)
    assert torch.equal(gf2().pixels[1, :192].view(64, 3), want[1][0])
