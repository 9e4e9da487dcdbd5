#!/bin/bash
# usage: tools/scale_probe.sh N tag [bench args...]  -- one N-GPU bench line into gpurun_out/r2_scale_<tag>.json (env passes through)
N=$1; tag=$2; shift 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 8 --warmup 3 "$@" > gpurun_out/r2_scale_$tag.json 2> gpurun_out/r2_scale_$tag.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_scale_$tag.json").read().splitlines()[-1]); print("$tag", d["n_gpus"], "gpus", round(d["value"],1), "samples/s", round(d["ms_per_step"],2), "ms")
except Exception as e: print("$tag failed", e)
PY
