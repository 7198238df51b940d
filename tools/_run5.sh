timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2d_pytest.log
bash tools/variants.sh "GSR_ACCUM_CLEAR=1" "GSR_ACCUM_CLEAR=0" 2>&1 | tee gpurun_out/r2d_variants.log
python tools/pcie_probe.py | tee gpurun_out/r2d_pcie.json
