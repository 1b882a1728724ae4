#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_relattn_gpu.py tests/test_elementwise_gpu.py tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_r2i.log
timeout 300 python tools/bench_attn_sweep.py 1024 4096 2>&1 | tee gpurun_out/attn_sweep_r2i.log
timeout 300 python tools/bench_ln.py 2>&1 | tail -12 | tee gpurun_out/bench_ln_r2i.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2i.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
t=open("gpurun_out/bench_r2i.err").read()
i=t.index("per-kernel breakdown (ms/step): ")+len("per-kernel breakdown (ms/step): ")
b=json.loads(t[i:].splitlines()[0])
print({k:round(v["ms_per_step"],3) for k,v in b.items()})
PY
