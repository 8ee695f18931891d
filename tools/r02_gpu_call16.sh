#!/bin/bash
# round-2 GPU call 16: aux face-prior warps (f3) parity, Up(q) with a float4 tap table / PRE template, swizzled blur_down
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_aux.py tests/test_gpu_ops.py tests/test_gpu_sampler.py tests/test_gpu_round2.py -m gpu -q -x -s -k "not 256_T10" 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02v_pytest.log; cat gpurun_out/r02v_pytest.log
timeout 300 python tests/gpu_probes/hbm_probe.py 2>&1 | tail -6 > gpurun_out/r02v_hbm.log; cat gpurun_out/r02v_hbm.log
