mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c6_pytest.log 2>&1
tail -5 gpurun_out/c6_pytest.log
TIO_B200_NOISE_OVERLAP=0 timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/c6_bench_overlap0.json 2> gpurun_out/c6_b0.err
TIO_B200_NOISE_OVERLAP=1 timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/c6_bench_overlap1.json 2> gpurun_out/c6_b1.err
python - <<'PY'
import json
for f in ("gpurun_out/c6_bench_overlap0.json","gpurun_out/c6_bench_overlap1.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["e2e"]["ms_per_step"], d["e2e"]["plain_call_ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/c6_b1.err
