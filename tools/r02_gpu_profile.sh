#!/bin/bash
# round-2 profile call: full GPU tests, default bench line, launch list of one forward, conv traffic, ncu --set full of
# the top kernels (each only after its command ran clean without ncu)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu.log; tail -2 gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 1200 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; cut -c1-400 gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --workload fwd-sweep > gpurun_out/r02_bench_fwd_sweep.json 2> gpurun_out/r02_bench_fwd_sweep.err; cut -c1-300 gpurun_out/r02_bench_fwd_sweep.json
timeout 300 python tests/gpu_probes/hbm_probe.py > gpurun_out/r02_hbm_probe.txt 2>&1
# launch list of one graphed forward (the last forward of the probe)
NWARM=1 NTIMED=1 timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02_fwd_plain.log 2>&1 &&
NWARM=1 NTIMED=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/r02_fwd_launches.csv python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02_ncu_fwd.log 2>&1
# DRAM / L2->SM traffic of every conv launch of one eager forward
NOGRAPH=1 NWARM=1 NTIMED=1 timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv_igemm -c 6000 --csv --log-file gpurun_out/r02_conv_traffic.csv python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02_ncu_traffic.log 2>&1
# ncu --set full: conv shapes, deform, attention
for sh in "10 256 64 64" "1 256 64 64" "10 128 128 128" "1 256 64 432" "10 64 256 256"; do
  tag=$(echo $sh | tr ' ' '_')
  timeout 100 python tests/gpu_probes/conv_graph.py $sh 8 > gpurun_out/r02_conv_${tag}_plain.log 2>&1 &&
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 10 -c 1 -o gpurun_out/r02_conv_${tag} python tests/gpu_probes/conv_graph.py $sh 8 > gpurun_out/r02_ncu_conv_${tag}.log 2>&1
done
timeout 100 python tests/gpu_probes/deform_one.py 64 256 256 > gpurun_out/r02_deform64_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:deform_conv -s 3 -c 1 -o gpurun_out/r02_deform64 python tests/gpu_probes/deform_one.py 64 256 256 > gpurun_out/r02_ncu_deform64.log 2>&1
timeout 100 python tests/gpu_probes/deform_one.py 128 128 128 > gpurun_out/r02_deform128_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:deform_conv -s 3 -c 1 -o gpurun_out/r02_deform128 python tests/gpu_probes/deform_one.py 128 128 128 > gpurun_out/r02_ncu_deform128.log 2>&1
timeout 100 python tests/gpu_probes/attn_probe.py 4 16 16 10 > gpurun_out/r02_attn_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_spatial_tc -s 3 -c 1 -o gpurun_out/r02_attn_tc python tests/gpu_probes/attn_probe.py 4 16 16 10 > gpurun_out/r02_ncu_attn.log 2>&1
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
echo done
