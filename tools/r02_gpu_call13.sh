#!/bin/bash
# round-2 GPU call 13: fused GroupNorm statistics (conv epilogue + finalize) — tests and forward A/B
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02r_pytest_kernels.log; tail -4 gpurun_out/r02r_pytest_kernels.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_kernels.py 2>&1 | tail -15 > gpurun_out/r02r_pytest_rest.log; tail -4 gpurun_out/r02r_pytest_rest.log
for f in 1 0; do
  echo "FLAIR_FUSED_GN=$f" >> gpurun_out/r02r_fwd.log
  FLAIR_FUSED_GN=$f timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video 2>&1 | tail -2 >> gpurun_out/r02r_fwd.log
  FLAIR_FUSED_GN=$f timeout 300 python tests/gpu_probes/unet_perf_probe.py 64 image 2>&1 | tail -2 >> gpurun_out/r02r_fwd.log
done
cat gpurun_out/r02r_fwd.log
for sh in "10 256 64 64" "1 256 64 64" "10 128 128 128"; do timeout 100 python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02r_conv.log 2>&1; done; cat gpurun_out/r02r_conv.log
