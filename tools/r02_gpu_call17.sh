#!/bin/bash
# round-2 GPU call 17: pipelined GN loads, floor'd gn_apply grid, update kernel without 64-bit divisions, blur_down
# whole-tile staging: parity + HBM probe + forward time
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py tests/test_gpu_ops.py tests/test_gpu_sampler.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02w_pytest.log; cat gpurun_out/r02w_pytest.log
timeout 300 python tests/gpu_probes/hbm_probe.py > gpurun_out/r02w_hbm.log 2>&1; cat gpurun_out/r02w_hbm.log
timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video 2>&1 | tail -2 | head -1 > gpurun_out/r02w_fwd.log; cat gpurun_out/r02w_fwd.log
