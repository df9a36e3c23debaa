#!/bin/bash
# 1 vs 2 co-resident CTAs per SM for the conv kernels: step time and the per-op table for each setting
for cfg in "YX_CTAS_PER_SM=1" "YX_CTAS_MAXPIX=400" "YX_CTAS_MAXPIX=1600" "YX_CTAS_MAXPIX=6400" "YX_CTAS_PER_SM=2"; do
  tag=$(echo $cfg | tr '=' '_')
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras --profile-ops > gpurun_out/ctas_$tag.json 2> gpurun_out/ctas_$tag.txt
  python - <<PY
import json
d = json.loads(open("gpurun_out/ctas_$tag.json").read().strip().splitlines()[-1])
print("$cfg", "value %.0f img/s  %.3f ms/step  e2e %.0f  parity %s  clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity_checked"], d["clocks"]["sm_mhz"]))
PY
done
