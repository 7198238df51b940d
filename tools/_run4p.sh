source tools/gpu_scale.sh r2p >/dev/null 2>&1 || true
TAG=r2p
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "pull" > gpurun_out/${TAG}_pytest_multi4.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/${TAG}_pytest_multi4.log
run c3_4gpu_pull 4 GSR_PEER_REDUCE=3 -- --no-train-step --no-other-configs --no-cpu-baseline
run c3_2gpu_pull 2 GSR_PEER_REDUCE=3 -- --no-train-step --no-other-configs --no-cpu-baseline --no-e2e
