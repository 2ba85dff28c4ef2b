#!/bin/bash
# One GPU session: tests, bench, launch list.  Usage (from the repo root, through gpurun):
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tag] [pytest -k expression]'
TAG=${1:-run}
KEXPR=${2:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
if [ -n "$KEXPR" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 -k "$KEXPR" -s > $OUT/pytest.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 -rA > $OUT/pytest.log 2>&1
fi
echo "pytest rc=$?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?"
tail -c 600 $OUT/bench.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"] and (d["e2e"]["value"], d["e2e"]["seconds"], d["e2e"]["pinned"]),
          "roofline", d["roofline"] and (d["roofline"]["frac"], d["roofline"]["ms_per_launch"], d["roofline"]["composite"]["frac"]),
          "iso", d["s_iso"] and (d["s_iso"]["value"], d["s_iso"]["composite"]["frac"]),
          "10m", d["fit_10m"] and [(l["k"], l["ms_per_iteration"], l["composite"]["frac"], l["rounds_executed"]) for l in d["fit_10m"].get("levels", [])] or d["fit_10m"],
          "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "encode", d["encode"])
except Exception as e:
    print("bench parse failed", e)
PY
