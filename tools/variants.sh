#!/bin/bash
# run bench.py under several tuning-env settings; one line per variant
# usage: tools/variants.sh "VAR=val VAR2=val" "..." ...
mkdir -p gpurun_out
for v in "$@"; do
  env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err
  python - "$v" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/v.json"))
    print(sys.argv[1] or "default", "| step", round(d["ms_per_step"],4), "fwd", round(d["fwd_ms"],4), "bwd", round(d["bwd_ms"],4), {k:round(v["ms"],3) for k,v in d["stages"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
