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

# ---- round 2 additions
KM="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,launch__registers_per_thread,launch__grid_size"
if has kmetrics; then
  # one metrics pass over every launch of one step of each workload (+ the fused optimizer step, + the decode loop)
  timeout 900 ncu --metrics $KM --clock-control none --profile-from-start off -c 1500 --csv --log-file gpurun_out/kmetrics_rl_$TAG.csv \
      python bench.py --steps 1 --warmup 2 --profile-only --optimizer > gpurun_out/kmetrics_rl_$TAG.log 2>&1
  timeout 900 ncu --metrics $KM --clock-control none --profile-from-start off -c 1500 --csv --log-file gpurun_out/kmetrics_atari_$TAG.csv \
      python bench.py --steps 1 --warmup 2 --profile-only --workload atari > gpurun_out/kmetrics_atari_$TAG.log 2>&1
  timeout 600 ncu --metrics $KM --clock-control none -k regex:"decode|ring_append|masked_argmax" -s 200 -c 200 --csv --log-file gpurun_out/kmetrics_decode_$TAG.csv \
      python tools/bench_decode.py 1 8 > gpurun_out/kmetrics_decode_$TAG.log 2>&1
  ls -la gpurun_out/kmetrics_*_$TAG.csv
fi
if has decode; then timeout 600 python tools/bench_decode.py 1 32 2>&1 | tee gpurun_out/bench_decode_$TAG.json; fi
if has ncufull2; then
  for k in "gemm_kernel" "relattn_fwd_kernel<128, 0>" "relattn_fwd_kernel<128, 2>" "relattn_bwd_dkdv" "relattn_bwd_band_kernel<128, 0>" "relattn_bwd_band_kernel<128, 1>" "ln_bwd" "ln_fwd"; do
    f=$(echo "$k" | tr -c 'a-zA-Z0-9' '_')
    timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$k" -s 30 -c 1 -o gpurun_out/full_${f}_$TAG -f \
        python bench.py --steps 1 --warmup 2 --profile-only > gpurun_out/full_${f}_$TAG.log 2>&1
  done
  ls -la gpurun_out/full_*_$TAG.ncu-rep
fi
