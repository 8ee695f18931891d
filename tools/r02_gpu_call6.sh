#!/bin/bash
# round-2 GPU call 6: where does the forward's time go now? BasicVSR++ module timings + ncu launch lists
set -u
mkdir -p gpurun_out
for a in "64 256" "128 128"; do
  timeout 200 python tests/gpu_probes/vsr_probe.py $a >> gpurun_out/r02j_vsr.log 2>&1
  FLAIR_PDL=1 timeout 200 python tests/gpu_probes/vsr_probe.py $a >> gpurun_out/r02j_vsr_pdl.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02j_vsr64_launches.csv python tests/gpu_probes/vsr_probe.py 64 256 > gpurun_out/r02j_ncu_vsr64.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02j_vsr128_launches.csv python tests/gpu_probes/vsr_probe.py 128 128 > gpurun_out/r02j_ncu_vsr128.log 2>&1
NWARM=1 NTIMED=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/r02j_fwd_launches.csv python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02j_ncu_fwd.log 2>&1
cat gpurun_out/r02j_vsr.log gpurun_out/r02j_vsr_pdl.log
wc -l gpurun_out/*.csv
echo done
