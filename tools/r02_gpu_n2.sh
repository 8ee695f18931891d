#!/bin/bash
# round-2 two-GPU call: NCCL sharding tests + the N=2 bench line (weak replicas + strong x8 32-frame clip)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4 > gpurun_out/r02f_pytest_multi.log; cat gpurun_out/r02f_pytest_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 1 --warmup 3 > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err
cut -c1-400 gpurun_out/r02f_bench_n2.json; tail -3 gpurun_out/r02f_bench_n2.err
