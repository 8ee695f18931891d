#!/bin/bash
# round-2 GPU call 9: A/B of the conv epilogue change against the library of commit 06455bd on the SAME box, and the
# per-tile overhead of the MMA pipeline (debug bits)
set -u
mkdir -p gpurun_out
REF=flair_b200/_ab/lib_06455bd.so
for sh in "10 256 64 64" "1 256 64 64" "1 256 64 432" "1 128 128 128" "10 128 128 128"; do
  echo -n "ref: " >> gpurun_out/r02m_ab.log; FLAIR_B200_LIB=$REF timeout 100 python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02m_ab.log 2>&1
  echo -n "new: " >> gpurun_out/r02m_ab.log; timeout 100 python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02m_ab.log 2>&1
done
cat gpurun_out/r02m_ab.log
for sh in "10 256 64 64" "10 128 128 128"; do
for d in 0 2 1 9 17 33 57 61; do echo -n "debug=$d: " >> gpurun_out/r02m_debug.log; FLAIR_CONV_DEBUG=$d timeout 100 python tests/gpu_probes/conv_graph.py $sh >> gpurun_out/r02m_debug.log 2>&1; done
done
cat gpurun_out/r02m_debug.log
echo "ref:" >> gpurun_out/r02m_vsr.log; FLAIR_B200_LIB=$REF timeout 300 python tests/gpu_probes/vsr_probe.py 64 256 2>&1 | grep -v -i warn | head -1 >> gpurun_out/r02m_vsr.log
echo "new:" >> gpurun_out/r02m_vsr.log; timeout 300 python tests/gpu_probes/vsr_probe.py 64 256 2>&1 | grep -v -i warn >> gpurun_out/r02m_vsr.log
cat gpurun_out/r02m_vsr.log
echo done
