#!/bin/bash
# GPU visit r2f: few-row GEMM + tiled decode attention (tests, decode bench A/B), full GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_r2f.log
timeout 400 python tools/bench_decode.py 1 32 2>&1 | tee gpurun_out/bench_decode_r2f.json
DB1_NO_SKINNY=1 timeout 400 python tools/bench_decode.py 1 32 2>&1 | tee gpurun_out/bench_decode_r2f_noskinny.json
DB1_DECODE_PER_KEY=1 timeout 400 python tools/bench_decode.py 1 32 2>&1 | tee gpurun_out/bench_decode_r2f_perkey.json
