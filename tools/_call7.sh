mkdir -p gpurun_out
TIO_B200_NOISE=exact TIO_B200_NOISE_OVERLAP=1 TRACE_STEPS=2 timeout 300 python tools/gpu_trace.py 2>&1 | grep -v "^$" | tail -45 > gpurun_out/c7_trace_overlap1.log
TIO_B200_NOISE=exact TIO_B200_NOISE_OVERLAP=0 TRACE_STEPS=2 TRACE_SUMMARY=1 timeout 300 python tools/gpu_trace.py 2>&1 | tail -2 > gpurun_out/c7_trace_overlap0.log
cat gpurun_out/c7_trace_overlap1.log gpurun_out/c7_trace_overlap0.log
timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 10 > gpurun_out/c7_bench.json 2> gpurun_out/c7_b.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c7_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["e2e"]["plain_call_ms_per_step"])
PY
