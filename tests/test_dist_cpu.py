"""Multi-rank host logic on CPU (gloo, world_size 2): ray/row sharding and the single all-gather of
rendered pixels must reproduce the single-process result bit for bit (rays are independent).
The per-ray compute is a stand-in CPU function here; the CUDA kernels are exercised by the -m gpu tests."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp_

from mipnerf_pl_b200.rays import Rays, random_ray_batch
from mipnerf_pl_b200.render import gather_rows, render_sharded, shard_bounds, shard_rows


def fake_forward(rays: Rays):
    """Deterministic per-ray function standing in for MipNerf.forward (3 'rgb' + 1 'distance')."""
    rgb = torch.sin(rays.origins * 3.0 + rays.directions) * rays.radii * 1e3
    dist_ = (rays.near + rays.far)[:, 0] * rays.viewdirs[:, 2]
    return [rgb, dist_]


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rays = random_ray_batch(n, seed=5)
        got = render_sharded(fake_forward, rays, world, rank)
        want = fake_forward(rays)
        ok = all(torch.equal(g, w) for g, w in zip(got, want))
        # ragged gather with explicit counts
        counts = [3, 0, 5][:world] if world == 3 else [4, 1]
        local = torch.full((counts[rank], 2), float(rank))
        full = gather_rows(local, counts)
        ok = ok and full.shape == (sum(counts), 2) and torch.equal(full[:counts[0]], torch.zeros(counts[0], 2))
        torch.save(ok, os.path.join(out_dir, f"ok{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 37])
def test_sharded_render_matches_single_process(tmp_path, n):
    world = 2
    port = 29500 + (os.getpid() % 2000) + n
    mp_.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    assert all(torch.load(tmp_path / f"ok{r}.pt") for r in range(world))


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 800, 4096, 640000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_rows(800, 8, 3) == (300, 400)


def _grad_worker(rank, world, port, out_dir):
    """allreduce_grads: DDP's mean over ranks of every parameter gradient, as ONE flat all-reduce."""
    from mipnerf_pl_b200.train import allreduce_grads
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        shapes = [(256, 96), (256,), (1, 256), (3, 128), (3,)]
        per_rank = [[torch.randn(*s, generator=g) for s in shapes] for _ in range(world)]
        params = [torch.nn.Parameter(torch.zeros(*s)) for s in shapes] + [torch.nn.Parameter(torch.zeros(2))]
        for p, gr in zip(params, per_rank[rank]):
            p.grad = gr.clone()                              # the last parameter has no gradient: skipped
        allreduce_grads(params)
        want = [sum(per_rank[r][i] for r in range(world)) / world for i in range(len(shapes))]
        ok = all(torch.allclose(p.grad, w, rtol=0, atol=1e-6) for p, w in zip(params, want)) and params[-1].grad is None
        allreduce_grads(params, average=False)
        ok = ok and all(torch.allclose(p.grad, w * world, rtol=0, atol=1e-5) for p, w in zip(params, want))
        # gradients that are consecutive views of one buffer (what forward_backward sets up): reduced in place,
        # no concatenate / scatter — same values, and the views still alias the buffer afterwards
        flat = torch.cat([t.reshape(-1) for t in per_rank[rank]])
        off = 0
        for p, sh in zip(params, shapes):
            n = int(torch.tensor(sh).prod())
            p.grad = flat[off:off + n].view(*sh)
            off += n
        from mipnerf_pl_b200.train import _flat_view
        ok = ok and _flat_view([p.grad for p in params[:-1]]) is not None
        allreduce_grads(params)
        ok = ok and all(torch.allclose(p.grad, w, rtol=0, atol=1e-6) for p, w in zip(params, want))
        ok = ok and params[0].grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()
        torch.save(ok, os.path.join(out_dir, f"gok{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_allreduce_grads_world2(tmp_path):
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp_.spawn(_grad_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(torch.load(tmp_path / f"gok{r}.pt") for r in range(world))
