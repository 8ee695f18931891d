#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02g_pytest_kernels.log
if grep -q "failed\|error" gpurun_out/r02g_pytest_kernels.log; then echo "KERNEL TESTS FAILED"; cat gpurun_out/r02g_pytest_kernels.log; exit 1; fi
for sh in "1 256 64 64" "1 128 128 128" "1 256 128 64" "1 256 64 432" "10 256 64 64" "1 128 256 128" "1 128 128 432" "10 128 128 128" "10 64 256 256" "10 32 256 256" "10 16 512 512"; do
  python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02g_conv.log 2>&1
done
for d in 1; do for sh in "10 256 64 64" "10 128 128 128"; do echo -n "debug=$d: " >> gpurun_out/r02g_conv.log; FLAIR_CONV_DEBUG=$d python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02g_conv.log 2>&1; done; done
python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02g_fwd.log 2>&1
FLAIR_PDL=1 python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02g_fwd_pdl.log 2>&1
python -m pytest tests -m gpu -q --deselect tests/test_gpu_kernels.py 2>&1 | tail -12 > gpurun_out/r02g_pytest_rest.log
echo done
