#!/bin/bash
mkdir -p gpurun_out
export MIPNERF_B200_TC_VARIANT=v4
echo "== v4 quick correctness (smoke-like) + stress"
timeout 300 python - <<'PY'
import torch, mipnerf_pl_b200 as mp
dev=torch.device("cuda",0)
for prec in ("bf16","fp16"):
    m=mp.MipNerf(precision=prec); m.load_state_dict(mp.make_state_dict(seed=0)); m=m.to(dev).eval()
    ref=mp.MipNerf(precision="fp32"); ref.load_state_dict(mp.make_state_dict(seed=0)); ref=ref.to(dev).eval()
    for n in (4, 301, 4096):
        rays=mp.namedtuple_map(lambda t:t.to(dev), mp.random_ray_batch(n, seed=3))
        a=m(rays, False, True); b=ref(rays, False, True)
        torch.cuda.synchronize()
        print(prec, n, [float((x[0]-y[0]).abs().max()) for x,y in zip(a,b)])
    for i in range(100):
        m(rays, False, True)
    torch.cuda.synchronize(); print(prec, "stress ok")
PY
echo "== v4 tests"
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -k "v4" 2>&1 | tail -4
echo "== bench v4 vs default"
for V in v4 pair; do
MIPNERF_B200_TC_VARIANT=$V timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-frame --no-train --no-parity-mode 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$V', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['launch_ms'], (d.get('parity') or {}).get('trained_like'))
"
done
