mkdir -p gpurun_out
for pf in 0 222 444 888 1776; do
  echo "== PREFETCH=$pf" >> gpurun_out/c5_k1.log
  TIO_B200_K1_PREFETCH=$pf timeout 300 python tools/k1_dev.py time 2>&1 | grep -E "TIME (affine|elastic) +box=24" >> gpurun_out/c5_k1.log
done
TIO_B200_K1_PREFETCH=444 timeout 300 python tools/k1_dev.py check 2>&1 | grep CHECK >> gpurun_out/c5_k1.log
cat gpurun_out/c5_k1.log
