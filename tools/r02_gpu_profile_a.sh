#!/bin/bash
# round-2 profile call A: full GPU tests, smoke, default bench line, image-mode sweep, HBM probe (+ GN chunk cap A/B)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu.log; tail -2 gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 1200 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; cut -c1-300 gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --workload fwd-sweep > gpurun_out/r02_bench_fwd_sweep.json 2> gpurun_out/r02_bench_fwd_sweep.err
timeout 300 python tests/gpu_probes/hbm_probe.py > gpurun_out/r02_hbm_probe.txt 2>&1
for c in 592 1184; do echo "FLAIR_GN_CHUNKS=$c" >> gpurun_out/r02_hbm_probe_chunks.txt; FLAIR_GN_CHUNKS=$c timeout 300 python tests/gpu_probes/hbm_probe.py 2>&1 | grep -i "gn_\|stats\|apply" >> gpurun_out/r02_hbm_probe_chunks.txt; done
cat gpurun_out/r02_hbm_probe.txt gpurun_out/r02_hbm_probe_chunks.txt
timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video 512 > gpurun_out/r02_fwd_512.log 2>&1; tail -2 gpurun_out/r02_fwd_512.log
timeout 300 python tests/gpu_probes/sr3_perf_probe.py > gpurun_out/r02_fwd_sr3.log 2>&1; tail -1 gpurun_out/r02_fwd_sr3.log
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
echo done
