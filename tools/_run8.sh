# 8-GPU session (gpurun --gpus 8): multi-GPU parity tests on every visible GPU, C3 at 8 and 4 GPUs, the multicast variant, C5 at 8
source tools/gpu_scale.sh r2s >/dev/null 2>&1 || true
TAG=r2s
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/${TAG}_pytest_multi.log
run c3_8gpu 8 GSR_DUMMY=0 -- --no-train-step --no-other-configs --no-cpu-baseline
run c3_8gpu_multimem 8 GSR_PEER_REDUCE=2 -- --no-train-step --no-other-configs --no-cpu-baseline --no-e2e
run c5_8gpu 8 GSR_DUMMY=0 -- --config C5 --no-train-step --no-other-configs --no-cpu-baseline --no-e2e --steps 10
run c3_4gpu 4 GSR_DUMMY=0 -- --no-train-step --no-other-configs --no-cpu-baseline
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
