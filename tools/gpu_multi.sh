#!/bin/bash
# Multi-GPU session: sharded fit == unsharded fit (tools/dist_check.py) and the bench at N ranks.
#   gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_multi.sh <tag> <N> [bench|nobench]'
TAG=${1:-multi}
N=${2:-2}
WHAT=${3:-bench}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/dist_check.py > $OUT/dist$N.log 2>&1
echo "dist_check rc=$?"
grep -v "^W\|^\[W\|Warning" $OUT/dist$N.log | tail -9
if [ "$WHAT" = "bench" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench$N.json 2> $OUT/bench$N.err
  echo "bench rc=$?"
  tail -c 300 $OUT/bench$N.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench$N.json").read().strip().splitlines()[-1])
    c = d["roofline"]["composite"] if d.get("roofline") else {}
    print("N=$N value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"], "composite", c.get("frac"),
          "iso", d["s_iso"] and d["s_iso"]["value"])
    print("passes", c.get("passes_per_step"), "R", c.get("rounds_executed_per_step"))
except Exception as e:
    print("parse failed", e)
PY
fi
