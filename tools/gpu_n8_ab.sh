#!/bin/bash
# 8-GPU A/B of the engine's bucket layout on ONE box: (A) round-1 style tail (no embedding split, one bucket per layer),
# (B) embedding split + one bucket per layer, (C) embedding split + ff / attn buckets per layer
mkdir -p gpurun_out
run() { # tag, env...
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-extra > gpurun_out/bench_n8_$tag.json 2> gpurun_out/bench_n8_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n8_$tag.json").read().strip().splitlines()[-1])
    print("$tag", d["value"], d["ms_per_step"], d.get("exposed_comm_ms"), d["dp_check"]["ok"], d["dp_check"]["buckets"])
except Exception as e:
    print("$tag ERR", e); print(open("gpurun_out/bench_n8_$tag.err").read()[-800:])
PY
}
run A DB1_EMB_SPLIT=0 DB1_BUCKET_SPLIT_FF=0
run B DB1_EMB_SPLIT=1 DB1_BUCKET_SPLIT_FF=0
run C DB1_EMB_SPLIT=1 DB1_BUCKET_SPLIT_FF=1
