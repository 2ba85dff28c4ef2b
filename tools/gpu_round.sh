#!/bin/bash
# One GPU session: tests, bench, launch list, one ncu capture.  Usage (from the repo root, through gpurun):
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh <tag> [pytest -k expression] [ncu kernel regex]'
TAG=${1:-run}
KEXPR=${2:-}
NCUK=${3:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
if [ "$KEXPR" != "none" ]; then
if [ -n "$KEXPR" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 -k "$KEXPR" -rA > $OUT/pytest.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 -rA > $OUT/pytest.log 2>&1
fi
echo "pytest rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?"
tail -c 600 $OUT/bench.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"] and (d["e2e"]["value"], d["e2e"]["seconds"], d["e2e"]["pinned"]),
          "roofline", d["roofline"] and (d["roofline"]["frac"], d["roofline"]["ms_per_launch"], d["roofline"]["composite"]["frac"], d["roofline"].get("other_kernel", {}).get("ms_per_launch")),
          "iso", d["s_iso"] and (d["s_iso"]["value"], d["s_iso"]["composite"]["frac"]),
          "10m", d["fit_10m"] and [(l["k"], l["ms_per_iteration"], l["composite"]["frac"], l["rounds_executed"], l.get("passes")) for l in d["fit_10m"].get("levels", [])] or d["fit_10m"],
          "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "encode", d["encode"] and d["encode"].get("value"))
    c = d["roofline"]["composite"]
    print("mix passes", c["passes_per_step"], "R", c["rounds_executed_per_step"], "miss", c["window_misses_per_step"])
except Exception as e:
    print("bench parse failed", e)
PY
for data in iso mix; do
  PROF_DATA=$data PROF_RESIDUAL=1 PROF_K=${PROF_K:-128} python tools/profile_iter.py > $OUT/profile_iter_$data.log 2>&1
  tail -4 $OUT/profile_iter_$data.log
done
if [ -n "$NCUK" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 3 --warmup 3 --no-extras > $OUT/ncu_bench.log 2>&1
  echo "launch list rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:$NCUK -s ${NCU_SKIP:-30} -c 1 -f -o $OUT/$NCUK \
      python tools/profile_iter.py > $OUT/ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
