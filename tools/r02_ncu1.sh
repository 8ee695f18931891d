#!/bin/bash
set -u
mkdir -p gpurun_out
python tests/gpu_probes/conv_graph.py 10 256 64 64 8 > gpurun_out/r02h_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 10 -c 1 -o gpurun_out/r02h_conv64_t10 python tests/gpu_probes/conv_graph.py 10 256 64 64 8 > gpurun_out/r02h_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 10 -c 1 -o gpurun_out/r02h_conv128_t1 python tests/gpu_probes/conv_graph.py 1 128 128 128 8 > gpurun_out/r02h_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
