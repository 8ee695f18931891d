#!/bin/bash
# round-2 profile call B: launch list of one graphed forward, conv traffic over one eager forward, ncu --set full of the
# top kernels (raw CSV exported on the box; the .ncu-rep files are not copied back: 64 MiB limit)
set -u
mkdir -p gpurun_out
NWARM=1 NTIMED=1 timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02_fwd_plain.log 2>&1 &&
NWARM=1 NTIMED=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/r02_fwd_launches.csv python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02_ncu_fwd.log 2>&1
NOGRAPH=1 NWARM=1 NTIMED=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum --clock-control none -k regex:conv_igemm -c 4000 --csv --log-file gpurun_out/r02_conv_traffic.csv python tests/gpu_probes/unet_perf_probe.py 10 video > gpurun_out/r02_ncu_traffic.log 2>&1
for sh in "10 256 64 64" "1 256 64 64" "10 128 128 128" "1 256 64 432"; do
  tag=$(echo $sh | tr ' ' '_')
  timeout 100 python tests/gpu_probes/conv_graph.py $sh 8 > gpurun_out/r02_conv_${tag}_plain.log 2>&1 &&
  timeout 300 ncu --set full --clock-control none -k regex:conv_igemm -s 10 -c 1 -o /tmp/r02_conv_${tag} python tests/gpu_probes/conv_graph.py $sh 8 > gpurun_out/r02_ncu_conv_${tag}.log 2>&1
  ncu -i /tmp/r02_conv_${tag}.ncu-rep --page raw --csv > gpurun_out/r02_conv_${tag}_raw.csv 2>/dev/null
done
for a in "64 256 256" "128 128 128"; do
  tag=$(echo $a | cut -d' ' -f1)
  timeout 100 python tests/gpu_probes/deform_one.py $a > gpurun_out/r02_deform${tag}_plain.log 2>&1 &&
  timeout 300 ncu --set full --clock-control none -k regex:deform_conv -s 3 -c 1 -o /tmp/r02_deform${tag} python tests/gpu_probes/deform_one.py $a > gpurun_out/r02_ncu_deform${tag}.log 2>&1
  ncu -i /tmp/r02_deform${tag}.ncu-rep --page raw --csv > gpurun_out/r02_deform${tag}_raw.csv 2>/dev/null
done
timeout 100 python tests/gpu_probes/attn_probe.py 4 16 16 10 > gpurun_out/r02_attn_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none -k regex:attn_spatial_tc -s 3 -c 1 -o /tmp/r02_attn_tc python tests/gpu_probes/attn_probe.py 4 16 16 10 > gpurun_out/r02_ncu_attn.log 2>&1
ncu -i /tmp/r02_attn_tc.ncu-rep --page raw --csv > gpurun_out/r02_attn_tc_raw.csv 2>/dev/null
du -sh gpurun_out
echo done
