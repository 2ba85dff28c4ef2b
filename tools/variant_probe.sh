#!/bin/bash
# Builds kernel variants (RQK_NVCC_EXTRA switches) on the GPU box and times each with `bench.py --no-extras`:
# HIST launch alone (roofline.ms_per_launch), the bid-list launch, and the bench step.
#   gpurun --timeout 1200 -- 'bash tools/variant_probe.sh "" "-DRQK_HIST_PINGPONG" ...'
mkdir -p gpurun_out/variants
for v in "$@"; do
  RQK_NVCC_EXTRA="$v" python -c "from generative_ranking_recommender_b200 import build; build.build(force=True)" > /dev/null 2>&1
  tag=$(echo "base$v" | tr -c 'A-Za-z0-9_\n' '_')
  python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/variants/$tag.json 2> gpurun_out/variants/$tag.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/variants/$tag.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("variant [$v]: step %.3f ms, HIST %.1f us (%.3f), bidlist %.1f us, composite %.3f" % (
    d["ms_per_step"], r["ms_per_launch"] * 1e3, r["frac"], r["other_kernel"]["ms_per_launch"] * 1e3, r["composite"]["frac"]))
PY
done
python -c "from generative_ranking_recommender_b200 import build; build.build(force=True)" > /dev/null 2>&1
