#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02q_pytest.log
tail -3 gpurun_out/r02q_pytest.log
for a in "64 256" "128 128"; do timeout 300 python tests/gpu_probes/vsr_probe.py $a 2>&1 | grep -v -i warn | grep "BasicVSRPP\|deform" >> gpurun_out/r02q_vsr.log; done
FLAIR_DEFORM_COOP=0 timeout 300 python tests/gpu_probes/vsr_probe.py 128 128 2>&1 | grep -v -i warn | grep "BasicVSRPP\|deform" >> gpurun_out/r02q_vsr.log
cat gpurun_out/r02q_vsr.log
for a in "64 256 256" "128 128 128"; do timeout 100 python tests/gpu_probes/deform_one.py $a 2>&1 | tail -1 >> gpurun_out/r02q_deform_one.log; done
cat gpurun_out/r02q_deform_one.log
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_kernels.py 2>&1 | tail -8 > gpurun_out/r02q_pytest_rest.log
tail -3 gpurun_out/r02q_pytest_rest.log
timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video 2>&1 | tail -2 > gpurun_out/r02q_fwd.log; cat gpurun_out/r02q_fwd.log


timeout 300 ncu --set full --clock-control none --import-source on -k regex:deform_conv -s 3 -c 1 -o gpurun_out/r02q_deform64 python tests/gpu_probes/deform_one.py 64 256 256 > gpurun_out/r02q_ncu.log 2>&1
