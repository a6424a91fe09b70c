"""Real multi-GPU checks (NCCL, one process per GPU; skipped on a single-GPU box — the gloo tests in
test_dist_cpu.py cover the host logic everywhere):
  * a frame rendered row-sharded over 2 GPUs + one all_gather is BIT-identical to the single-GPU frame
    (SURVEY.md §8e determinism check);
  * ray-sharded training: per-rank forward_backward with the global mask sum / ray count, then ONE all-reduce of
    the flat gradient buffer, reproduces the full-batch gradients."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp_

from helpers import make_state_dict

pytestmark = pytest.mark.gpu

import mipnerf_pl_b200 as mp  # noqa: E402


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        ok = {}
        # ---- frame: sharded == single
        model = mp.MipNerf(precision="bf16")
        model.load_state_dict(make_state_dict(seed=0, kind="trained_like"))
        model = model.to(dev).eval()
        pose = mp.spheric_pose(0.7)
        sharded = mp.render_frame(model, pose, 96, 80, True, world=world, rank=rank)
        single = mp.render_frame(model, pose, 96, 80, True, world=1, rank=0)
        ok["frame"] = all(torch.equal(a, b) for a, b in zip(sharded, single))
        # ---- training: shard gradients + all-reduce == full-batch gradients
        b = 300
        rays = mp.namedtuple_map(lambda t: t.to(dev), mp.random_ray_batch(b, seed=4, multiscale=True))
        rgbs = torch.rand(b, 3, generator=torch.Generator().manual_seed(1)).to(dev)
        tm = mp.MipNerf()
        tm.load_state_dict(make_state_dict(seed=2, kind="xavier"))
        tm = tm.to(dev)
        full = mp.forward_backward(tm, rays, rgbs, False, True)
        g_full = [p.grad.clone() for p in tm.parameters()]
        lo, hi = mp.shard_bounds(b, world, rank)
        shard = mp.namedtuple_map(lambda t: t[lo:hi], rays)
        part = mp.forward_backward(tm, shard, rgbs[lo:hi], False, True, mask_sum=rays.lossmult.sum(), global_rays=b)
        mp.allreduce_grads(tm.parameters(), average=False)
        loss = part["loss"].clone()
        dist.all_reduce(loss)
        worst = max(float((p.grad - g).norm() / g.norm()) for p, g in zip(tm.parameters(), g_full))
        ok["train"] = worst <= 1e-5 and abs(float(loss) - float(full["loss"])) <= 1e-5 * abs(float(full["loss"]))
        ok["worst"] = worst
        torch.save(ok, os.path.join(out_dir, f"ok{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL)")
def test_sharded_render_and_training_on_two_gpus(tmp_path):
    world = 2
    port = 30500 + (os.getpid() % 2000)
    mp_.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        ok = torch.load(tmp_path / f"ok{r}.pt")
        print(f"rank {r}: {ok}")
        assert ok["frame"], "sharded frame differs from the single-GPU frame"
        assert ok["train"], f"sharded gradients differ: {ok['worst']:.2e}"
