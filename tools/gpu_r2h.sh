#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_r2h.log
DB1_WGRAD_STREAM=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_r2h_s0.json 2> gpurun_out/bench_r2h_s0.err
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_r2h_s1.json 2> gpurun_out/bench_r2h_s1.err
python - <<'PY'
import json
for f in ("bench_r2h_s0","bench_r2h_s1"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
    except Exception as e:
        print(f,"ERR",e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
