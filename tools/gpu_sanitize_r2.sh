#!/bin/bash
# compute-sanitizer memcheck over the kernels added / rewritten in round 2 (attention backward consumers, tiled P / dS,
# decode attention + ring append + masked arg-max, few-row GEMM, CUDA-graph decode, packed-math LayerNorm backward)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest -x -q -m gpu tests/test_relattn_gpu.py tests/test_mems_gpu.py tests/test_elementwise_gpu.py \
  "tests/test_gemm_gpu.py::test_few_row_gemm_matches_fp32_and_the_tensor_core_path" \
  > gpurun_out/sanitizer_r2.log 2>&1
echo "exit code $?" >> gpurun_out/sanitizer_r2.log
tail -12 gpurun_out/sanitizer_r2.log
