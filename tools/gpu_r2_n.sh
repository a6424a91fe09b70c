#!/bin/bash
# training step, third pass: enc tile image, per-slab view encodings; launch list of one step; ncu of the training level kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q 2>&1 | tail -3
timeout 600 python tools/train_bench.py --precision bf16 > gpurun_out/r2_train_bf16_fused4.json 2> gpurun_out/r2_train_bf16_fused4.err; tail -c 1300 gpurun_out/r2_train_bf16_fused4.json; tail -3 gpurun_out/r2_train_bf16_fused4.err
echo "== launch list of one training step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_train_launches.csv python tools/ncu_train_target.py bf16 4096 > gpurun_out/r2_train_launches.log 2>&1; tail -1 gpurun_out/r2_train_launches.log
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_train_launches.csv')) if len(r)>5 and r[0].isdigit()]
# second step only: kernels after the first half
half=len(rows)//2
agg=collections.OrderedDict()
for r in rows[half:]:
    name=r[4].split('(')[0][-60:]
    t=float(r[-1].replace(',',''))
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=t
unit=rows[0][-2]
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:16]: print(f"{t:12.1f} {unit} {n:4d}  {k}")
print('total', sum(t for n,t in agg.values()), unit, len(rows)-half, 'launches')
PY
echo "== ncu of the training-mode level kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mlp_level_kernel" -s 2 -c 2 -f -o gpurun_out/r2_prof_level_train python tools/ncu_train_target.py bf16 4096 2>&1 | tail -1
