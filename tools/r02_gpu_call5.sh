#!/bin/bash
# round-2 GPU call 5 (re-entry): validate HEAD — all GPU tests, smoke, conv / forward timings, default bench line
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02i_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02i_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02i_smoke.log 2>&1
for sh in "1 256 64 64" "1 128 128 128" "1 256 128 64" "1 256 64 432" "10 256 64 64" "1 128 256 128" "1 128 128 432" "10 128 128 128" "10 64 256 256" "10 32 256 256" "10 16 512 512"; do
  timeout 120 python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02i_conv.log 2>&1
done
timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02i_fwd.log 2>&1
timeout 300 python tests/gpu_probes/unet_perf_probe.py 64 image > gpurun_out/r02i_fwd_image64.log 2>&1
timeout 300 python tests/gpu_probes/sr3_perf_probe.py > gpurun_out/r02i_fwd_sr3.log 2>&1
timeout 300 python tests/gpu_probes/r02_step_probe.py gaussian > gpurun_out/r02i_step.log 2>&1
timeout 900 python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
tail -3 gpurun_out/r02i_pytest.log; cat gpurun_out/r02i_conv.log gpurun_out/r02i_fwd.log; cat gpurun_out/r02i_bench.json
echo done
