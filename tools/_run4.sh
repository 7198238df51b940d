timeout 600 python -m pytest tests/test_knn.py tests/test_filter3d.py tests/test_multi_gpu.py -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c_pytest.log
bash tools/gpu_scale.sh r2c 2
timeout 300 python tools/bench_aux.py 3000000 1000 8 > gpurun_out/r2c_aux.json 2>gpurun_out/r2c_aux.err; cat gpurun_out/r2c_aux.json
