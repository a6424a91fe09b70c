#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
import mipnerf_pl_b200 as mp
dev="cuda:0"
for variant in ("pair","single"):
    os.environ["MIPNERF_B200_TC_VARIANT"]=variant
    for prec in ("bf16","fp32"):
        m=mp.MipNerf(precision=prec); m.load_state_dict(mp.make_state_dict(1)); m=m.to(dev).eval()
        rays=mp.namedtuple_map(lambda t:t.to(dev), mp.random_ray_batch(37, seed=2))
        out=m(rays, False, True); torch.cuda.synchronize()
        print(variant, prec, float(out[1][0].sum()))
f=mp.render_frame(m, mp.spheric_pose(0.3), 16, 16); torch.cuda.synchronize(); print("frame", float(f[1].sum()))
PY
for tool in memcheck synccheck; do
  echo "== $tool"; timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py 2>&1 | tail -12 | tee gpurun_out/sanitizer_$tool.txt
done
