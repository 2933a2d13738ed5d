mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_ops.py -x -q -k "label_pv or stream" ) > gpurun_out/c3_pytest.log 2>&1
tail -5 gpurun_out/c3_pytest.log
TIO_B200_K1_REUSE=0 timeout 300 python tools/k1_dev.py time > gpurun_out/c3_k1_reuse0.log 2>&1
TIO_B200_K1_REUSE=3 timeout 300 python tools/k1_dev.py check time > gpurun_out/c3_k1_reuse3.log 2>&1
grep -E "TIME|CHECK" gpurun_out/c3_k1_reuse0.log gpurun_out/c3_k1_reuse3.log
timeout 300 python tools/e2e_profile.py > gpurun_out/c3_e2e_profile.log 2>&1
head -5 gpurun_out/c3_e2e_profile.log
