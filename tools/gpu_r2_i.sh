#!/bin/bash
# single-GPU evidence pass: x3 / rng / metrics tests, ncu of v3, ncu launch list of a bench run, ncu dram metrics of the ray kernels
mkdir -p gpurun_out
echo "== tests (x3, tensorcore 16-bit goldens, rng, metrics)"
timeout 900 python -m pytest tests/test_gpu_x3.py tests/test_gpu_rng.py tests/test_metrics.py tests/test_gpu_graph.py -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|rays >|Error" | tail -20
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q 2>&1 | tail -4
echo "== ncu v3"
MIPNERF_B200_TC_VARIANT=v3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_level_kernel_v3 -s 4 -c 2 -f -o gpurun_out/r2_prof_v3 python tools/ncu_target.py bf16 4096 4 2>&1 | tail -1
echo "== ncu fp16x3 (rolled epilogue)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_level -s 4 -c 1 -f -o gpurun_out/r2_prof_fp16x3_rolled python tools/ncu_target.py fp16x3 4096 4 2>&1 | tail -1
echo "== ncu launch list of one bench run"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-frame --no-parity --no-graph > gpurun_out/r2_ncu_list.log 2>&1; tail -1 gpurun_out/r2_ncu_list.log | cut -c1-200
echo "== ncu dram metrics of the stand-alone ray kernels"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"coarse_t|cast_rays|resample|ipe_kernel|composite|distloss|generate_rays" -c 40 --csv --log-file gpurun_out/r2_ray_kernels_ncu.csv python tools/ray_kernel_bw.py --reps 2 > gpurun_out/r2_ray_kernels.log 2>&1; tail -2 gpurun_out/r2_ray_kernels.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches.csv gpurun_out/r2_ray_kernels_ncu.csv
