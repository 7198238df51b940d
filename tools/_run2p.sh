source tools/gpu_scale.sh r2p >/dev/null 2>&1 || true
TAG=r2p
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "pull or peer_red" > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/${TAG}_pytest_multi.log
timeout 200 python tools/debug_multi.py 2 3 2>&1 | grep "^mode" | sort
run c3_2gpu_pull 2 GSR_PEER_REDUCE=3 -- --no-train-step --no-other-configs --no-cpu-baseline --no-e2e
