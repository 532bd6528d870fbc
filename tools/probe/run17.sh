cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_exact.py -m gpu -x -q -k "calibration_state" > gpurun_out/r02_pytest9.log 2>&1; echo "pytest rc=$?"
tail -n 15 gpurun_out/r02_pytest9.log
