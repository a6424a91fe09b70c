#!/bin/bash
# compute-sanitizer memcheck / synccheck over every kernel family at small sizes: the three
# tensor-core variants (deterministic + randomized: in-kernel prologue and resampler), the MLP-only stage entry,
# the fp32 path, a frame, the distloss kernel, one fp32 training step (backward + Adam) and the tensor-core
# training mode (tcgen05 linear / wgrad kernels).
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
import mipnerf_pl_b200 as mp
dev = "cuda:0"
rays = mp.namedtuple_map(lambda t: t.to(dev), mp.random_ray_batch(37, seed=2, multiscale=True))
g = torch.Generator(device=dev).manual_seed(0)
t_rand = torch.rand(37, 129, device=dev, generator=g)
u_jit = torch.rand(37, 129, device=dev, generator=g) * (1 / 129 - 1.2e-7)
for variant in ("pair", "single", "shared"):
    os.environ["MIPNERF_B200_TC_VARIANT"] = variant
    for prec in ("bf16", "fp16", "fp32"):
        m = mp.MipNerf(precision=prec); m.load_state_dict(mp.make_state_dict(1)); m = m.to(dev).eval()
        out = m(rays, False, True); torch.cuda.synchronize()
        outr = m(rays, True, False, t_rand=t_rand, u_jitter=u_jit, return_inds=True); torch.cuda.synchronize()
        print(variant, prec, float(out[1][0].sum()), float(outr[1][0].sum()))
os.environ["MIPNERF_B200_TC_VARIANT"] = "pair"
x = torch.rand(5, 128, 96, device=dev); venc = torch.randn(5, 27, device=dev)
rgb, dens = m.mlp(x, venc, precision="bf16"); torch.cuda.synchronize(); print("mlp stage", float(rgb.sum()), float(dens.sum()))
f = mp.render_frame(m, mp.spheric_pose(0.3), 16, 16); torch.cuda.synchronize(); print("frame", float(f[1].sum()))
print("distloss", float(mp.distloss(out[1][3], out[1][4])))
tm = mp.MipNerf(); tm.load_state_dict(mp.make_state_dict(2)); tm = tm.to(dev)
opt = mp.FusedAdam(tm.parameters(), lr=5e-4)
o = mp.forward_backward(tm, rays, torch.rand(37, 3, device=dev), True, True, t_rand=t_rand, u_jitter=u_jit); opt.step()
torch.cuda.synchronize(); print("train", float(o["loss"]))
for prec in ("bf16", "fp16"):                      # tensor-core training mode: linear_tc + wgrad_tc kernels
    tm = mp.MipNerf(precision=prec); tm.load_state_dict(mp.make_state_dict(2)); tm = tm.to(dev)
    o = mp.forward_backward(tm, rays, torch.rand(37, 3, device=dev), False, True)
    torch.cuda.synchronize(); print("train", prec, float(o["loss"]))
PY
for tool in memcheck synccheck; do
  echo "== $tool"; timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py 2>&1 | tail -20 | tee gpurun_out/sanitizer_$tool.txt
done
