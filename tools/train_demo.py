"""End-to-end loop on the GPU with no host data path: synthetic Blender-format scene on disk -> DeviceRayBank ->
MipNeRFSystem.training_step / FusedAdam / MipLRDecay -> render_image of a validation view.  Prints the training
PSNR every few steps and the step rate.

    python tools/train_demo.py [--steps 200] [--batch 1024] [--size 32]
"""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402

import mipnerf_pl_b200 as mp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--size", type=int, default=32)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    root = tempfile.mkdtemp()
    mp.write_synthetic_blender_scene(root, n_images=8, height=args.size, width=args.size, seed=1)
    mp.convert_blender_to_multiscale(root, root + "_ms", 3)
    hp = mp.default_hparams(**{"dataset_name": "multi_blender", "data_path": root + "_ms", "train.batch_size": args.batch,
                               "optimizer.lr_delay_steps": 50, "optimizer.max_steps": args.steps, "val.chunk_size": 4096})
    torch.manual_seed(0)
    system = mp.MipNeRFSystem(hp).to(dev)
    (opt,), (sched,) = system.configure_optimizers()
    bank = mp.DeviceRayBank(mp.load_multicam_scene(root + "_ms", "train"), dev)
    log = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for step in range(args.steps):
        batch = bank.sample(args.batch)
        loss = system.training_step(batch, step)
        opt.zero_grad(set_to_none=False)
        loss.backward()
        opt.step()
        sched["scheduler"].step()
        if step % max(1, args.steps // 10) == 0 or step == args.steps - 1:
            log.append({"step": step, "loss": float(loss.detach()), "psnr": float(system._logged["train/psnr"]),
                        "lr": opt.param_groups[0]["lr"]})
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # validation view through render_image (models/nerf_system.py:151-177)
    val = mp.Multicam(root + "_ms", "val", batch_type="single_image")
    rays, img = val[0]
    to = lambda a: torch.from_numpy(a)[None].to(dev)  # noqa: E731
    coarse, fine, _ = system.render_image((mp.namedtuple_map(to, rays), to(img)))
    val_psnr = float(mp.calc_psnr(fine, to(img)))
    print(json.dumps({"what": "train_demo: DeviceRayBank -> training_step (fp32) -> FusedAdam, synthetic random-pixel scene",
                      "steps": args.steps, "batch": args.batch, "steps_per_s": args.steps / dt,
                      "rays_per_s": args.steps * args.batch / dt, "log": log, "val_psnr_view0": val_psnr}))


if __name__ == "__main__":
    main()
