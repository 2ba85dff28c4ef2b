#!/bin/bash
# BASELINE configs 3 and 4 at N ranks: 10 M x 512 rows sharded over N GPUs (one global fit), encode sweep to 100 M.
#   gpurun --gpus N --timeout 900 -- 'bash tools/gpu_ns3_multi.sh <tag> <N> [encode]'
TAG=${1:-ns3}
N=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
ROWS=$((10000000 / N))
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --rows $ROWS --steps 12 --warmup 3 --no-extras > $OUT/bench_10m_${N}gpu.json 2> $OUT/bench_10m_${N}gpu.err
echo "bench rc=$?"; tail -c 200 $OUT/bench_10m_${N}gpu.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_10m_${N}gpu.json").read().strip().splitlines()[-1])
    c = d["roofline"]["composite"]
    print("10M over $N GPUs: value", d["value"], "ms/step", d["ms_per_step"], "composite", c["frac"], "R", c["rounds_executed_per_step"])
except Exception as e:
    print("parse failed", e)
PY
if [ "$3" = "encode" ]; then
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
      tools/ns3_measure.py encode > $OUT/encode_${N}gpu.log 2>&1
  echo "encode rc=$?"; grep "^encode" $OUT/encode_${N}gpu.log
fi
