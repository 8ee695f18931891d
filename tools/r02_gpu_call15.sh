#!/bin/bash
# round-2 GPU call 15: vectorised blur_down staging + 4-wide Up(q) in the fused update (parity + HBM probe),
# ABBA ordering for the conv / forward A/B against the f0c93b1 library (is the +2 ms an order effect?)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sampler.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02u_pytest.log; tail -3 gpurun_out/r02u_pytest.log
timeout 300 python tests/gpu_probes/hbm_probe.py > gpurun_out/r02u_hbm.log 2>&1; cat gpurun_out/r02u_hbm.log
REF=flair_b200/_ab/lib_f0c93b1.so
for sh in "1 256 64 64" "10 256 64 64"; do
  for v in new ref slim0 ref new slim0; do
    echo -n "$v: " >> gpurun_out/r02u_conv.log
    case $v in
      ref) FLAIR_FUSED_GN=0 FLAIR_B200_LIB=$REF timeout 100 python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02u_conv.log 2>&1;;
      new) timeout 100 python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02u_conv.log 2>&1;;
      slim0) FLAIR_CONV_SLIM=0 timeout 100 python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02u_conv.log 2>&1;;
    esac
  done
done; cat gpurun_out/r02u_conv.log
for v in new ref ref new; do
  echo -n "$v: " >> gpurun_out/r02u_fwd.log
  if [ $v = ref ]; then FLAIR_FUSED_GN=0 FLAIR_B200_LIB=$REF timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video 2>&1 | tail -2 | head -1 >> gpurun_out/r02u_fwd.log
  else timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video 2>&1 | tail -2 | head -1 >> gpurun_out/r02u_fwd.log; fi
done
cat gpurun_out/r02u_fwd.log
