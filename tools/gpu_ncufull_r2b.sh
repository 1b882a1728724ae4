#!/bin/bash
# One `ncu --set full` capture per kernel of the step (and of the decode step), round 2. Reports land in gpurun_out/;
# tools/ncu_summary.py turns them into profiles/ncu_full_r2.md.
mkdir -p gpurun_out
cap() { # file-tag, kernel regex, launches to skip, command...
  f=$1; k=$2; s=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off --graph-profiling node --kernel-name-base demangled -k regex:"$k" -s $s -c 1 \
      -o gpurun_out/full_${f}_r2 -f "$@" > gpurun_out/full_${f}_r2.log 2>&1
  tail -1 gpurun_out/full_${f}_r2.log
}
B="python bench.py --steps 1 --warmup 2 --profile-only"
cap gemm_plain      "gemm_kernel<.int.256, .int.0, .int.2>"            40 $B
cap gemm_geglu      "gemm_kernel<.int.256, .int.2, .int.2>"            5  $B
cap gemm_dgeglu     "gemm_kernel<.int.256, .int.3, .int.2>"            5  $B
cap gemm_qkv        "gemm_kernel<.int.256, .int.1, .int.2>"            5  $B
cap attn_fwd        "relattn_fwd_kernel<.int.128, .int.0>"        5  $B
cap attn_bwd_ds     "relattn_fwd_kernel<.int.128, .int.2>"        5  $B
cap attn_bwd_dq     "relattn_bwd_band_kernel<.int.128, .int.0>"   5  $B
cap attn_bwd_dr     "relattn_bwd_band_kernel<.int.128, .int.1>"   5  $B
cap skinny_geglu    "skinny_gemm_kernel<.int.1, .int.2>"          5  python tools/decode_kernel_times.py
ls -la gpurun_out/full_*_r2.ncu-rep | awk '{print $5, $9}'
