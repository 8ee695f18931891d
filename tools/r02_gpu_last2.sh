# functional check of the 8-GPU strong-scaling workloads (BASELINE configs 3 / 4) on ONE GPU: same code path as
# bench.py --gpus 8 (window plan, per-window noise keys, stitching), world == 1 so no NCCL message
mkdir -p gpurun_out
timeout 110 python bench.py --workload x16 --mode sharded --steps 1 --warmup 0 --no-cpu-baseline --no-sweep \
  > gpurun_out/r02_bench_x16_sharded_n1.json 2> gpurun_out/r02_bench_x16_sharded_n1.err; echo "x16 rc=$?"
cut -c1-250 gpurun_out/r02_bench_x16_sharded_n1.json; tail -3 gpurun_out/r02_bench_x16_sharded_n1.err
timeout 150 python bench.py --workload jpeg --mode sharded --steps 1 --warmup 0 --no-cpu-baseline --no-sweep \
  > gpurun_out/r02_bench_jpeg_sharded_n1.json 2> gpurun_out/r02_bench_jpeg_sharded_n1.err; echo "jpeg rc=$?"
cut -c1-250 gpurun_out/r02_bench_jpeg_sharded_n1.json; tail -3 gpurun_out/r02_bench_jpeg_sharded_n1.err
