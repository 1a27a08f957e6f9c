#!/bin/bash
# A/B of environment settings (TH_CHAIN_*, TH_CHUNK_PTS, ...) on one GPU box: one short bench run per
# setting, prints rays/s and the per-category split.
# usage: tools/env_ab.sh <tag> "<env assignments>" ...
tag=$1; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  env $envs timeout 300 python bench.py --no-cpu-baseline --no-culled --steps 3 --warmup 2 > gpurun_out/${tag}_ab$i.json 2> gpurun_out/${tag}_ab$i.err
  python - "$envs" gpurun_out/${tag}_ab$i.json <<'PY'
import json, sys
lines = [l for l in open(sys.argv[2]) if l.startswith("{")]
if not lines:
    print(sys.argv[1], "NO OUTPUT"); sys.exit(0)
d = json.loads(lines[-1])
print(f"{sys.argv[1]:40s} {d['value']:10.0f} rays/s {d['ms_per_step']:7.2f} ms", d["ms_per_step_by_category"], "launches", d.get("gpu_launches"))
PY
  tail -2 gpurun_out/${tag}_ab$i.err
  i=$((i+1))
done
