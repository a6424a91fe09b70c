"""ncu target: two fused training steps (bf16 by default) of a 4096-ray batch.  Usage: ncu ... python tools/ncu_train_target.py [precision] [rays]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import mipnerf_pl_b200 as mp  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda", 0)
model = mp.MipNerf(precision=precision)
model.load_state_dict(mp.make_state_dict(seed=0, kind="xavier"))
model = model.to(dev)
rays = mp.namedtuple_map(lambda t: t.to(dev), mp.random_ray_batch(n, seed=0, multiscale=True))
rgbs = torch.rand(n, 3, device=dev)
for _ in range(2):
    mp.forward_backward(model, rays, rgbs, True, True)
torch.cuda.synchronize()
print("done")
