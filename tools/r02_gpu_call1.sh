#!/bin/bash
# round-2 GPU call 1: correctness of the new paths + A/B timings (run from the repo root under gpurun)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.log 2>&1
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_round2.py 2>&1 | tail -15 > gpurun_out/r02a_pytest_old.log
python -m pytest tests/test_gpu_round2.py -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r02a_pytest_new.log
python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02a_fwd.log 2>&1
FLAIR_PDL=1 python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02a_fwd_pdl.log 2>&1
for sh in "1 256 64 64" "1 128 128 128" "1 256 128 64" "1 256 64 432" "10 256 64 64" "1 128 256 128" "1 128 128 432"; do
  python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02a_conv.log 2>&1
  FLAIR_PDL=1 python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02a_conv_pdl.log 2>&1
done
python tests/gpu_probes/r02_step_probe.py gaussian > gpurun_out/r02a_step.log 2>&1
FLAIR_PDL=1 python tests/gpu_probes/r02_step_probe.py gaussian > gpurun_out/r02a_step_pdl.log 2>&1
python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1
echo done
