#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_mems_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_r2g.log
timeout 400 python tools/bench_decode.py 1 32 2>&1 | tee gpurun_out/bench_decode_r2g.json
DB1_DECODE_PER_KEY=1 timeout 400 python tools/bench_decode.py 1 32 2>&1 | tee gpurun_out/bench_decode_r2g_perkey.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none --profile-from-start off --graph-profiling node --csv --log-file gpurun_out/decode_kernels_r2g.csv python tools/decode_kernel_times.py > gpurun_out/decode_kernels_r2g.log 2>&1
tail -3 gpurun_out/decode_kernels_r2g.log; wc -l gpurun_out/decode_kernels_r2g.csv
