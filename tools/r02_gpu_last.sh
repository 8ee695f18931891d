set -x
mkdir -p gpurun_out
timeout 170 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r02_last_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_last_smoke.log
tail -3 gpurun_out/r02_last_smoke.log
timeout 150 python -m pytest tests/test_gpu_ops.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/r02_last_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_last_pytest.log
tail -3 gpurun_out/r02_last_pytest.log
