#!/bin/bash
# round-2 GPU call 8: epilogue addend prefetch + 32-bit warp indexing + tcgen05 spatial attention — tests, VSR step
# breakdown, forward time; MMA-rate probes and the CTA-0 timeline (trace build)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02l_pytest.log
tail -3 gpurun_out/r02l_pytest.log
for a in "64 256" "128 128"; do timeout 300 python tests/gpu_probes/vsr_probe.py $a 2>&1 | grep -v -i warn >> gpurun_out/r02l_vsr.log; done
cat gpurun_out/r02l_vsr.log
timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02l_fwd.log 2>&1; tail -2 gpurun_out/r02l_fwd.log
for sh in "1 256 64 64" "10 256 64 64" "1 256 64 432" "1 128 128 128" "10 128 128 128"; do
  for d in 0 1 4 5; do echo -n "debug=$d: " >> gpurun_out/r02l_conv_debug.log; FLAIR_CONV_DEBUG=$d timeout 100 python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02l_conv_debug.log 2>&1; done
done
cat gpurun_out/r02l_conv_debug.log
cd tests/gpu_probes/cu && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_conv_like mma_conv_like.cu && timeout 100 /tmp/mma_conv_like > ../../../gpurun_out/r02l_mma_conv_like.log 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate mma_rate.cu && timeout 100 /tmp/mma_rate > ../../../gpurun_out/r02l_mma_rate.log 2>&1
cd ../../..
cat gpurun_out/r02l_mma_conv_like.log gpurun_out/r02l_mma_rate.log
cp flair_b200/libflair_b200.so /tmp/lib_keep.so
touch flair_b200/csrc/conv_igemm.cu; FLAIR_BUILD_TRACE=1 python -m flair_b200.build > /dev/null 2>&1
for sh in "1 256 64 64" "10 256 64 64" "1 256 64 432" "1 128 128 128"; do timeout 100 python tests/gpu_probes/conv_trace.py $sh >> gpurun_out/r02l_conv_trace.txt 2>&1; done
cp /tmp/lib_keep.so flair_b200/libflair_b200.so
cat gpurun_out/r02l_conv_trace.txt
echo done
