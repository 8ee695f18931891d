#!/bin/bash
# round-2 GPU call 10: compact conv epilogue (fast path + addend prefetch) — kernel tests, A/B against 06455bd
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02n_pytest.log
tail -3 gpurun_out/r02n_pytest.log
REF=flair_b200/_ab/lib_06455bd.so
for sh in "10 256 64 64" "1 256 64 64" "1 256 64 432" "1 128 128 128" "10 128 128 128" "10 64 256 256"; do
  echo -n "ref: " >> gpurun_out/r02n_ab.log; FLAIR_B200_LIB=$REF timeout 100 python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02n_ab.log 2>&1
  echo -n "new: " >> gpurun_out/r02n_ab.log; timeout 100 python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02n_ab.log 2>&1
done
cat gpurun_out/r02n_ab.log
for a in "64 256" "128 128"; do
echo "ref:" >> gpurun_out/r02n_vsr.log; FLAIR_B200_LIB=$REF timeout 300 python tests/gpu_probes/vsr_probe.py $a 2>&1 | grep -v -i warn | head -1 >> gpurun_out/r02n_vsr.log
echo "new:" >> gpurun_out/r02n_vsr.log; timeout 300 python tests/gpu_probes/vsr_probe.py $a 2>&1 | grep -v -i warn >> gpurun_out/r02n_vsr.log
done
cat gpurun_out/r02n_vsr.log
timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video 2>&1 | tail -2 > gpurun_out/r02n_fwd.log
FLAIR_PDL=1 timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video 2>&1 | tail -2 > gpurun_out/r02n_fwd_pdl.log
cat gpurun_out/r02n_fwd.log gpurun_out/r02n_fwd_pdl.log
FLAIR_PDL=1 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02n_pytest_pdl.log
tail -3 gpurun_out/r02n_pytest_pdl.log
echo done
