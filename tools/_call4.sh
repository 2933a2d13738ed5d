mkdir -p gpurun_out
TIO_B200_K1_REUSE=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:resample_fast -c 2 -o gpurun_out/r2_k1_reuse python tools/k1_dev.py ncu > gpurun_out/c4_ncu.log 2>&1
tail -3 gpurun_out/c4_ncu.log
