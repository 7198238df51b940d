#!/bin/bash
# multi-GPU session.  usage: bash tools/gpu_scale.sh TAG N [N2 ...]   (run under gpurun --gpus max(N)): one default bench line per N.
# Sourced (`source tools/gpu_scale.sh TAG`) it only defines `run NAME N [ENV=..]... -- bench-args`, e.g. the sessions of round 2:
#   run c3_4gpu_pull 4 GSR_PEER_REDUCE=3 -- --no-train-step --no-other-configs --no-cpu-baseline
#   run c5_8gpu 8 GSR_DUMMY=0 -- --config C5 --no-train-step --no-other-configs --no-cpu-baseline --no-e2e --steps 10
TAG=$1; shift
mkdir -p gpurun_out
run() {  # run NAME N [env...] -- bench args
  local name=$1 n=$2; shift 2
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
      bench.py --gpus $n "$@" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${TAG}_${name}.json") if l.startswith("{")][-1])
    print("  ${name}: step", round(d["ms_per_step"], 4), "e2e", {k: round(v, 3) for k, v in (d.get("e2e") or {}).items() if k.startswith("ms")}, "check", (d.get("check") or {}).get("passed"))
    print("   ", {k: round(v["ms"], 3) for k, v in d.get("stages", {}).items()})
except Exception as e:
    print("  ${name} FAILED", e)
PY
}
if [ "${BASH_SOURCE[0]}" = "$0" ]; then for n in "$@"; do
  run c3_${n}gpu $n GSR_DUMMY=0 -- --no-train-step --no-other-configs --no-cpu-baseline
done; fi
