#!/bin/bash
# GPU session: full parity tests, golden generation for knn, aux timings, full bench of both arms
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 300 python tests/golden/make_golden_knn.py > gpurun_out/${TAG}_make_golden_knn.log 2>&1; echo "golden knn rc=$?"; cp tests/golden/knn_*.npz gpurun_out/ 2>/dev/null
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/bench_aux.py > gpurun_out/${TAG}_aux.json 2> gpurun_out/${TAG}_aux.err; echo "aux rc=$?"; cat gpurun_out/${TAG}_aux.json; tail -3 gpurun_out/${TAG}_aux.err
timeout 600 python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference > gpurun_out/${TAG}_bench_c3_reference.json 2> gpurun_out/${TAG}_bench_c3_reference.err; echo "ref rc=$?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_c3.json", "gpurun_out/${TAG}_bench_c3_reference.json"):
    try:
        d = json.load(open(f)); print(f, d["ms_per_step"], {k: v for k, v in d.get("e2e", {}).items() if k.startswith("ms")})
        print(" train_step", d.get("train_step")); print(" colour_op", {k: v for k, v in (d.get("colour_op") or {}).items() if k.endswith("_ms")})
    except Exception as e:
        print(f, "FAILED", e)
PY
