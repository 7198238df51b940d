#!/bin/bash
# Final evidence session of a round (ONE GPU, bounded): smoke, parity tests, both bench arms, aux timings, launch list,
# one ncu --set full pass over the kernels that changed.  usage: bash tools/gpu_final.sh TAG
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
timeout 240 python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench rc=$?"
timeout 150 python bench.py --impl reference > gpurun_out/${TAG}_bench_c3_reference.json 2> gpurun_out/${TAG}_bench_c3_reference.err; echo "ref rc=$?"
timeout 90 python tools/bench_aux.py > gpurun_out/${TAG}_aux.json 2> gpurun_out/${TAG}_aux.err; echo "aux rc=$?"; cat gpurun_out/${TAG}_aux.json
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-train-step --no-other-configs > gpurun_out/${TAG}_launches_bench.log 2>&1; echo "launches rc=$?"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:'preprocess_bwd|tile_offsets|rowscan_wide|cell_scatter' -s 8 -c 4 -o gpurun_out/${TAG}_ncu_misc -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-train-step --no-other-configs > gpurun_out/${TAG}_ncu_misc.log 2>&1; echo "ncu rc=$?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_c3.json", "gpurun_out/${TAG}_bench_c3_reference.json"):
    try:
        d = json.load(open(f)); print(f, d["ms_per_step"], {k: v for k, v in d.get("e2e", {}).items() if k.startswith("ms")})
        print("  ", {k: round(v["ms"], 3) for k, v in d.get("stages", {}).items()})
        print(" train_step", d.get("train_step")); print(" colour_op", {k: v for k, v in (d.get("colour_op") or {}).items() if k.endswith("_ms")})
    except Exception as e:
        print(f, "FAILED", e)
PY
