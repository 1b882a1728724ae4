#!/bin/bash
# One GPU visit: parity tests, GEMM micro-benchmark, bench line, ncu launch list + full capture of the top kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [tests|bench|ncu|all ...]
TAG=${1:-rX}; shift
WHAT=${*:-all}
mkdir -p gpurun_out
has() { [[ " $WHAT " == *" $1 "* || " $WHAT " == *" all "* ]]; }
if has tests; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_$TAG.log; fi
if has gemm; then timeout 300 python tools/bench_gemm.py 2>&1 | tee gpurun_out/bench_gemm_$TAG.log; fi
if has bench; then timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 1500 gpurun_out/bench_$TAG.json; cp gpurun_out/bench_breakdown_n1.json gpurun_out/breakdown_$TAG.json; fi
if has ncu; then
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 1 --warmup 2 --profile-only > gpurun_out/ncu_list_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_list_$TAG.log
fi
if has ncufull; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 1 -o gpurun_out/prof_gemm_$TAG -f \
      python tools/bench_gemm.py "ff2" > gpurun_out/ncu_gemm_$TAG.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:relattn -s 3 -c 1 -o gpurun_out/prof_attn_$TAG -f \
      python tools/bench_attn.py 1024 > gpurun_out/ncu_attn_$TAG.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:relattn -s 4 -c 1 -o gpurun_out/prof_attnbwd_$TAG -f \
      python tools/bench_attn_bwd.py 1024 > gpurun_out/ncu_attnbwd_$TAG.log 2>&1
  timeout 600 ncu --set full --clock-control none -k regex:ln_bwd_fused -s 2 -c 1 -o gpurun_out/prof_lnbwd_$TAG -f \
      python bench.py --steps 1 --warmup 1 --profile-only > gpurun_out/ncu_lnbwd_$TAG.log 2>&1
  ls -la gpurun_out/*_$TAG.ncu-rep
fi
