source tools/gpu_scale.sh r2s >/dev/null 2>&1 || true
TAG=r2s
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/${TAG}_pytest_multi4.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/${TAG}_pytest_multi4.log
run c3_4gpu 4 GSR_DUMMY=0 -- --no-train-step --no-other-configs --no-cpu-baseline
run c3_2gpu 2 GSR_DUMMY=0 -- --no-train-step --no-other-configs --no-cpu-baseline
run c3_4gpu_multimem 4 GSR_PEER_REDUCE=2 -- --no-train-step --no-other-configs --no-cpu-baseline --no-e2e
