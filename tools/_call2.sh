mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_ops.py -x -q -k "label_pv or stream or resample_ or spatial_target" ) > gpurun_out/c2_pytest.log 2>&1
tail -15 gpurun_out/c2_pytest.log
timeout 600 python tools/e2e_sweep.py > gpurun_out/c2_e2e_sweep.log 2>&1
cat gpurun_out/c2_e2e_sweep.log | tail -8
