#!/bin/bash
# quick GPU check: parity tests + one bench line per arm without the slow extras.  usage: bash tools/gpu_quick.sh TAG [extra bench args]
TAG=${1:-q}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --no-train-step --no-other-configs --no-cpu-baseline "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 400 python bench.py --impl reference --no-other-configs "$@" > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/${TAG}_bench_ref.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.json", "gpurun_out/${TAG}_bench_ref.json"):
    try:
        d = json.load(open(f)); print(f, "step", round(d["ms_per_step"], 4), "e2e", d.get("e2e"))
        print("  ", {k: round(v["ms"], 3) for k, v in d.get("stages", {}).items()})
    except Exception as e:
        print(f, "FAILED", e)
PY
