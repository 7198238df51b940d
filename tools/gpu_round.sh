#!/bin/bash
# One GPU-box session: tests, both bench arms, the ncu launch list and one --set full capture of the top kernels.
# usage (from the repo root on the box): bash tools/gpu_round.sh TAG
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference > gpurun_out/${TAG}_bench_c3_reference.json 2> gpurun_out/${TAG}_bench_c3_reference.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-train-step --no-other-configs > gpurun_out/${TAG}_launches_bench.log 2>&1; echo "launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'render_|preprocess_' -s 12 -c 4 -o gpurun_out/${TAG}_ncu_render -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-train-step --no-other-configs > gpurun_out/${TAG}_ncu_render.log 2>&1; echo "ncu render rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:appearance_ -s 2 -c 2 -o gpurun_out/${TAG}_ncu_appearance -f \
  python tools/profile_appearance.py > gpurun_out/${TAG}_ncu_appearance.log 2>&1; echo "ncu appearance rc=$?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_c3.json", "gpurun_out/${TAG}_bench_c3_reference.json"):
    try:
        d = json.load(open(f)); print(f, d["ms_per_step"], d.get("e2e", {}).get("ms_per_step"), {k: round(v["ms"], 3) for k, v in d.get("stages", {}).items()})
        print(" train_step", d.get("train_step")); print(" colour_op", {k: v for k, v in (d.get("colour_op") or {}).items() if k.endswith("_ms")}); print(" other", d.get("other_configs"))
    except Exception as e:
        print(f, "FAILED", e)
PY
