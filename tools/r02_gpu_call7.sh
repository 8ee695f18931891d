#!/bin/bash
set -u
mkdir -p gpurun_out
for a in "64 256" "128 128"; do timeout 300 python tests/gpu_probes/vsr_probe.py $a >> gpurun_out/r02k_vsr.log 2>&1; done
for v in 0 1 2; do
  FLAIR_ATTN_VMODE=$v timeout 120 python tests/gpu_probes/attn_probe.py 4 16 16 10 >> gpurun_out/r02k_attn.log 2>&1
done
FLAIR_ATTN_TC=0 timeout 120 python tests/gpu_probes/attn_probe.py 4 16 16 10 >> gpurun_out/r02k_attn.log 2>&1
grep -v Warning gpurun_out/r02k_vsr.log; grep attn gpurun_out/r02k_attn.log; tail -3 gpurun_out/r02k_attn.log
