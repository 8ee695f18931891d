#!/bin/bash
# round-2 GPU call 20: full GPU test suite at HEAD + how much the fused deformable conv depends on the coherence of the
# offsets (random offsets = what the synthetic weights produce; smooth = spatially coherent offsets / flows)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02h_pytest_gpu.log; cat gpurun_out/r02h_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 >> gpurun_out/r02h_pytest_gpu.log
for a in "64 256 256" "64 256 256 smooth" "128 128 128" "128 128 128 smooth"; do
  timeout 100 python tests/gpu_probes/deform_one.py $a 2>&1 | tail -1 >> gpurun_out/r02h_deform.log
done
cat gpurun_out/r02h_deform.log
