#!/bin/bash
# round-2 GPU call 3: single-slab halo conv — correctness, then timings and the CTA-0 timeline (trace build)
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02d_pytest_kernels.log
if grep -q "failed\|error" gpurun_out/r02d_pytest_kernels.log; then echo "KERNEL TESTS FAILED"; cat gpurun_out/r02d_pytest_kernels.log; exit 1; fi
python -m pytest tests -m gpu -q --deselect tests/test_gpu_kernels.py 2>&1 | tail -25 > gpurun_out/r02d_pytest_rest.log
for sh in "1 256 64 64" "1 128 128 128" "1 256 128 64" "1 256 64 432" "10 256 64 64" "1 128 256 128" "1 128 128 432" "10 128 128 128" "10 64 256 256" "10 32 256 256" "10 16 512 512"; do
  python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02d_conv.log 2>&1
  FLAIR_PDL=1 python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02d_conv_pdl.log 2>&1
done
python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02d_fwd.log 2>&1
FLAIR_PDL=1 python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02d_fwd_pdl.log 2>&1
NOVSR=1 python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02d_fwd_novsr.log 2>&1
python tests/gpu_probes/unet_perf_probe.py 64 image > gpurun_out/r02d_fwd_image64.log 2>&1
# CTA-0 timeline with the trace build (rebuilt on the box, then restored)
cp flair_b200/libflair_b200.so /tmp/lib_keep.so
touch flair_b200/csrc/conv_igemm.cu; FLAIR_BUILD_TRACE=1 python -m flair_b200.build > /dev/null 2>&1
for sh in "1 256 64 64" "1 128 128 128" "10 256 64 64" "1 256 64 432"; do python tests/gpu_probes/conv_trace.py $sh >> gpurun_out/r02d_conv_trace.txt 2>&1; done
cp /tmp/lib_keep.so flair_b200/libflair_b200.so
echo done
