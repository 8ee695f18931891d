#!/bin/bash
# round-2 GPU call 19: fused GroupNorm statistics on / off at HEAD (ABBA), Up(q) from staged LR rows (parity + probe),
# bench line with the statistics kernel separate
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_sampler.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r02y_pytest.log; cat gpurun_out/r02y_pytest.log
timeout 300 python tests/gpu_probes/hbm_probe.py 2>&1 | tail -5 > gpurun_out/r02y_hbm.log; cat gpurun_out/r02y_hbm.log
for f in 1 0 0 1; do
  echo -n "FLAIR_FUSED_GN=$f: " >> gpurun_out/r02y_fwd.log
  FLAIR_FUSED_GN=$f timeout 300 python tests/gpu_probes/unet_perf_probe.py 10 video 2>&1 | tail -2 | head -1 >> gpurun_out/r02y_fwd.log
done
cat gpurun_out/r02y_fwd.log
FLAIR_FUSED_GN=0 timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/r02y_bench_nofuse.json 2> gpurun_out/r02y_bench_nofuse.err; cut -c1-200 gpurun_out/r02y_bench_nofuse.json
