timeout 60 python -m pytest tests/test_adam.py tests/test_host_logic.py -m gpu -q 2>&1 | tail -2
timeout 100 python bench.py --no-train-step --no-other-configs --no-cpu-baseline --steps 5 > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/z_bench.json')); print(d['ms_per_step'], d['roofline_path'], {k:v for k,v in d['e2e'].items() if k.startswith('ms')})"
