#!/bin/bash
# round-2 GPU call 14: fused GN statistics with the parallel finalize — tests + same-box A/B
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r02t_pytest.log; tail -3 gpurun_out/r02t_pytest.log
REF=flair_b200/_ab/lib_f0c93b1.so
for sh in "10 256 64 64" "1 256 64 64" "10 128 128 128"; do
  echo -n "ref: " >> gpurun_out/r02t_conv.log; FLAIR_FUSED_GN=0 FLAIR_B200_LIB=$REF timeout 100 python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02t_conv.log 2>&1
  echo -n "new: " >> gpurun_out/r02t_conv.log; timeout 100 python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02t_conv.log 2>&1
done; cat gpurun_out/r02t_conv.log
echo "ref lib (f0c93b1, no fused GN):" >> gpurun_out/r02t_fwd.log
FLAIR_FUSED_GN=0 FLAIR_B200_LIB=$REF timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video 2>&1 | tail -2 | head -1 >> gpurun_out/r02t_fwd.log
for f in 1 0; do
  echo "FLAIR_FUSED_GN=$f" >> gpurun_out/r02t_fwd.log
  FLAIR_FUSED_GN=$f timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video 2>&1 | tail -2 | head -1 >> gpurun_out/r02t_fwd.log
  FLAIR_FUSED_GN=$f timeout 300 python tests/gpu_probes/unet_perf_probe.py 64 image 2>&1 | tail -2 | head -1 >> gpurun_out/r02t_fwd.log
done
cat gpurun_out/r02t_fwd.log
