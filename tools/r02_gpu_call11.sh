#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "deform" 2>&1 | tail -8 > gpurun_out/r02o_pytest.log
tail -3 gpurun_out/r02o_pytest.log
for c in 0 1; do echo "coop=$c" >> gpurun_out/r02o_vsr.log; FLAIR_DEFORM_COOP=$c timeout 300 python tests/gpu_probes/vsr_probe.py 128 128 2>&1 | grep -v -i warn | grep "BasicVSRPP\|deform" >> gpurun_out/r02o_vsr.log; done
cat gpurun_out/r02o_vsr.log
for c in 0 1; do FLAIR_DEFORM_COOP=$c timeout 100 python tests/gpu_probes/deform_one.py 128 128 128 2>&1 | tail -3 >> gpurun_out/r02o_deform_one.log; done
cat gpurun_out/r02o_deform_one.log
