#!/bin/bash
# round-2 GPU call 18: fast resampling GroupNorm apply kernels: parity + forward time
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py tests/test_gpu_unet_sr3.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02x_pytest.log; cat gpurun_out/r02x_pytest.log
timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video 2>&1 | tail -2 | head -1 > gpurun_out/r02x_fwd.log; cat gpurun_out/r02x_fwd.log
timeout 300 python tests/gpu_probes/sr3_perf_probe.py 2>&1 | tail -1 >> gpurun_out/r02x_fwd.log; tail -1 gpurun_out/r02x_fwd.log
