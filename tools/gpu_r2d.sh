#!/bin/bash
# GPU visit r2d: 16-softmax-warp attention build vs the 8-warp build, decode graph, full GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_r2d.log
DB1_ATTN_SPLIT=2 timeout 300 python tools/bench_attn_sweep.py 2>&1 | tee gpurun_out/attn_sweep_split2_r2d.log
timeout 300 python tools/bench_attn_sweep.py 2>&1 | tee gpurun_out/attn_sweep_split4_r2d.log
timeout 400 python tools/bench_decode.py 1 32 2>&1 | tee gpurun_out/bench_decode_r2d.json
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -c 600 gpurun_out/bench_r2d.err
