#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_r2k.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_r2k.log
timeout 400 python tools/bench_decode.py 1 32 2>&1 | tee gpurun_out/bench_decode_r2k.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none --profile-from-start off --graph-profiling node --csv --log-file gpurun_out/decode_kernels_r2k.csv python tools/decode_kernel_times.py > gpurun_out/decode_kernels_r2k.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off --graph-profiling node --kernel-name-base demangled -k regex:"skinny_gemm_kernel<.int.1, .int.2" -s 5 -c 1 -o gpurun_out/full_skinny_geglu_r2 -f python tools/decode_kernel_times.py > gpurun_out/full_skinny_geglu_r2.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err
tail -c 1500 gpurun_out/bench_r2k.json
