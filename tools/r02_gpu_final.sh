#!/bin/bash
# round-2 final measurement call: full GPU tests, smoke, default bench line, reference arm, HBM probe, launch list of one
# graphed forward, ncu --set full of the GroupNorm / update kernels (raw CSV exported on the box)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02f_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02f_pytest_gpu.log; tail -2 gpurun_out/r02f_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02f_smoke.log 2>&1; tail -2 gpurun_out/r02f_smoke.log
timeout 1200 python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; cut -c1-300 gpurun_out/r02f_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err; cut -c1-300 gpurun_out/r02f_bench_reference.json
timeout 300 python tests/gpu_probes/hbm_probe.py > gpurun_out/r02f_hbm_probe.txt 2>&1
NWARM=1 NTIMED=1 timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02f_fwd_plain.log 2>&1 &&
NWARM=1 NTIMED=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/r02f_fwd_launches.csv python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02f_ncu_fwd.log 2>&1
tail -2 gpurun_out/r02f_fwd_plain.log
for k in gn_apply_fast gn_stats sampler_update blur_down_tiled; do
  timeout 300 ncu --set full --clock-control none -k regex:$k -s 4 -c 1 -o /tmp/r02f_$k python tests/gpu_probes/hbm_probe.py > gpurun_out/r02f_ncu_$k.log 2>&1
  ncu -i /tmp/r02f_$k.ncu-rep --page raw --csv > gpurun_out/r02f_${k}_raw.csv 2>/dev/null
done
du -sh gpurun_out
echo done
