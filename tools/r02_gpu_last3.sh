mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_ragged.py -m gpu -q -s > gpurun_out/r02_ragged.log 2>&1; echo "ragged rc=$?" | tee -a gpurun_out/r02_ragged.log
grep -h "rel L2\|passed\|failed\|Error" gpurun_out/r02_ragged.log | tail -15
timeout 40 python tests/gpu_probes/t1_probe.py > gpurun_out/r02_t1_probe.log 2>&1; cat gpurun_out/r02_t1_probe.log | tail -4
timeout 150 python -m pytest tests -m gpu -q --deselect tests/test_gpu_ragged.py > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "full rc=$?" | tee -a gpurun_out/r02_pytest_gpu_final.log
tail -4 gpurun_out/r02_pytest_gpu_final.log
