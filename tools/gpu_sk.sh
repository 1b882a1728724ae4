#!/bin/bash
# GPU visit: full parity suite, then the stream-K GEMM tests (own process: a hang there must not take the suite down),
# then A/B of the GEMM micro-benchmark and the bench line with / without the stream-K tail.
TAG=${1:-r1k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gemm_gpu.py::test_gemm_stream_k_tail --deselect tests/test_gemm_gpu.py::test_gemm_stream_k_fused_epilogues 2>&1 | tail -8 | tee gpurun_out/pytest_$TAG.log
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k stream_k 2>&1 | tail -15 | tee gpurun_out/pytest_sk_$TAG.log
if grep -q "passed" gpurun_out/pytest_sk_$TAG.log && ! grep -q "failed" gpurun_out/pytest_sk_$TAG.log; then
  timeout 200 python tools/bench_gemm.py 2>&1 | tee gpurun_out/bench_gemm_$TAG.log
  DB1_GEMM_NO_SK=1 timeout 200 python tools/bench_gemm.py 2>&1 | tee gpurun_out/bench_gemm_nosk_$TAG.log
  timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 600 gpurun_out/bench_$TAG.json
  cp gpurun_out/bench_breakdown_n1.json gpurun_out/breakdown_$TAG.json 2>/dev/null
  DB1_GEMM_NO_SK=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nosk_$TAG.json 2> gpurun_out/bench_nosk_$TAG.err; tail -c 600 gpurun_out/bench_nosk_$TAG.json
else
  DB1_GEMM_NO_SK=1 timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_nosk_$TAG.json 2> gpurun_out/bench_nosk_$TAG.err; tail -c 600 gpurun_out/bench_nosk_$TAG.json
fi
