#!/bin/bash
# round-2 last call: default bench line at HEAD (fused GroupNorm statistics off) + the tests touched since the full run
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r02g_pytest.log; cat gpurun_out/r02g_pytest.log
timeout 1200 python bench.py > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; cut -c1-300 gpurun_out/r02g_bench_n1.json
timeout 300 python tests/gpu_probes/hbm_probe.py > gpurun_out/r02g_hbm_probe.txt 2>&1
